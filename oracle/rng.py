"""Random draws of the RJ-MCMC step, restated so that they can be REPLAYED.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference interleaves CPython's Mersenne Twister (``random.choices``, ``randint``,
``normalvariate``, ``gauss``, ``uniform``, ``random``) with legacy ``numpy.random``
(``uniform``, ``normal``, ``choice``) -- SURVEY.md quirk Q8.  Neither can be matched by a
counter-based device generator, so parity is defined on a *stream of uniforms*
``u in [0, 1)`` with fixed rules turning uniforms into each kind of draw.  The same rules
are implemented three times and tested against each other:

* here (oracle),
* in ``tests/golden/make_golden.py`` as monkey-patches of ``random.*`` / ``np.random.*``
  inside the unmodified reference (so the reference consumes the same tape), and
* in ``bayesbay_b200/csrc/rng.cuh`` (device).

Rules (``u()`` = next uniform of the stream):

=====================================  =========================================================
reference call                          restatement
=====================================  =========================================================
``random.random()``                     ``1 - u()``  (in (0, 1]; avoids log(0), quirk Q11)
``random.uniform(a, b)``                ``a + (b - a) * u()``
``np.random.uniform(a, b, n)``          ``a[i] + (b[i] - a[i]) * u()`` for i in order
``random.randint(a, b)``                ``a`` if a == b (no draw) else ``a + min(int(u()*n), n-1)``
``random.choices(pop, w)[0]``           ``0`` if len(pop) == 1 (no draw) else
                                        ``bisect_right(cumsum(w), u() * sum(w))`` clipped to n-1
                                        (CPython's own rule, Lib/random.py ``choices``)
``normalvariate/gauss(mu, s)``          ``mu + s * z`` with Box-Muller
``np.random.normal(mu, s, n)``          ``z = sqrt(-2 log(1 - u())) * cos(2 pi u())`` (2 draws per z)
``np.random.laplace(mu, b)``            numpy's inverse-CDF rule on one ``u()``
``np.random.choice(chains, 2, False)``  ordered pair ``a = int(u()*n)``, ``b = int(u()*(n-1))``, b += (b >= a)
=====================================  =========================================================

Streams: :class:`TapeStream` reads a recorded array; :class:`PhiloxStream` is
Philox4x32-10 (Salmon et al. 2011, Random123) with ``key = (seed_lo, seed_hi)`` and
``counter = (block, chain_id, iter_lo, iter_hi)``; draw ``n`` of a (chain, iteration) is
word pair ``n % 2`` of block ``n // 2``; a double is ``((w0 >> 5) * 2**26 + (w1 >> 6)) / 2**53``.
"""
from __future__ import annotations

import math
from bisect import bisect_right

import numpy as np

_M0 = 0xD2511F53
_M1 = 0xCD9E8D57
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = 0xFFFFFFFF
TWO_PI = 2.0 * math.pi

# domain-separation tags carried in the high bits of the iteration counter
STREAM_STEP = 0
STREAM_INIT = 1
STREAM_SWAP = 2


def philox4x32(counter, key, rounds=10):
    """Philox4x32 on python ints. counter: 4 words, key: 2 words -> 4 words."""
    c0, c1, c2, c3 = (int(x) & _MASK for x in counter)
    k0, k1 = (int(x) & _MASK for x in key)
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def words_to_double(w0, w1):
    return ((w0 >> 5) * 67108864.0 + (w1 >> 6)) / 9007199254740992.0


class PhiloxStream:
    """Uniform stream of one (seed, chain, iteration[, stream tag])."""

    def __init__(self, seed, chain, iteration, tag=STREAM_STEP):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.chain = chain & _MASK
        it = int(iteration)
        self.it_lo = it & _MASK
        self.it_hi = ((it >> 32) & 0x0FFFFFFF) | (tag << 28)
        self.n = 0
        self._block = None
        self._words = None

    def u(self):
        block = self.n >> 1
        if block != self._block:
            self._words = philox4x32((block, self.chain, self.it_lo, self.it_hi), self.key)
            self._block = block
        w = self._words
        half = self.n & 1
        self.n += 1
        return words_to_double(w[2 * half], w[2 * half + 1])


class TapeStream:
    """Uniform stream read from a recorded array (step-replay parity)."""

    def __init__(self, tape, cursor=0):
        self.tape = tape
        self.n = cursor

    def u(self):
        v = float(self.tape[self.n])
        self.n += 1
        return v


class NumpyStream:
    """Uniform stream from a numpy Generator (used by the CPU baseline; fast)."""

    def __init__(self, gen):
        self.gen = gen
        self.n = 0

    def u(self):
        self.n += 1
        return self.gen.random()


# ---- derived draws ---------------------------------------------------------
def draw_unit(s):
    """``random.random()`` restated: in (0, 1]."""
    return 1.0 - s.u()


def draw_uniform(s, a, b):
    return a + (b - a) * s.u()


def draw_randint(s, a, b):
    if a == b:
        return a
    n = b - a + 1
    return a + min(int(s.u() * n), n - 1)


def draw_choice(s, weights):
    n = len(weights)
    if n == 1:
        return 0
    cum = []
    tot = 0.0
    for w in weights:
        tot += w
        cum.append(tot)
    return min(bisect_right(cum, s.u() * tot), n - 1)


def draw_std_normal(s):
    u1 = s.u()
    u2 = s.u()
    return math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(TWO_PI * u2)


def draw_normal(s, mu, sigma):
    return mu + sigma * draw_std_normal(s)


def draw_laplace(s, mu, scale):
    u = s.u()
    if u >= 0.5:
        return mu - scale * math.log(2.0 - u - u)
    if u > 0.0:
        return mu + scale * math.log(u + u)
    return mu  # measure-zero; numpy rejects and redraws


def draw_pair(s, n):
    a = min(int(s.u() * n), n - 1)
    b = min(int(s.u() * (n - 1)), n - 2)
    if b >= a:
        b += 1
    return a, b
