"""The UNMODIFIED reference, importable on the build container and on the GPU box.

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``): used by ``tests/`` (plug-in and parity tests) and by
``bench.py --impl reference`` / the ``cpu_baseline`` leg.  The product never imports this module.

``__graft_entry__.build()`` installs the reference with pip into the git-ignored ``baseline/_ref`` (the driver's
recipe: ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref``
from a writable copy of /root/reference; ``--no-deps`` because matplotlib and shapely are not in the image) -- the
directory travels to the GPU box with the snapshot.  matplotlib / shapely are hard imports of
``discretization/_voronoi.py:7-11`` used only for plotting and polygon domains: empty ON-DISK stub packages stand in
for them (in-process ``sys.modules`` stubs do not reach joblib's loky workers, which start fresh interpreters).
"""
from __future__ import annotations

import importlib
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "baseline", "_ref")
STUB_DIR = os.path.join(REPO, "baseline", "_ref", "_stubs")

_STUBS = (("matplotlib/__init__.py", ""), ("matplotlib/pyplot.py", ""), ("shapely/__init__.py", ""),
          ("shapely/geometry.py", "class Polygon:\n    pass\n\n\nclass Point:\n    pass\n"))


def write_stubs():
    for rel, body in _STUBS:
        path = os.path.join(STUB_DIR, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if not os.path.exists(path):
            with open(path, "w") as f:
                f.write(body)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "bayesbay"))


def import_reference():
    """``import bayesbay`` from ``baseline/_ref`` (raises ImportError when it is not installed).  Also exports the
    paths through PYTHONPATH so that loky worker processes find the package, the stubs and this repository."""
    if "bayesbay" in sys.modules:
        # (already imported in this process -- e.g. by the fixture generator's scratch build of the same sources)
        bayesbay = sys.modules["bayesbay"]
        root = os.path.dirname(os.path.dirname(os.path.abspath(bayesbay.__file__)))
        env = os.environ.get("PYTHONPATH", "")
        # (stub packages that stand in for matplotlib / shapely in this process must reach the workers too)
        stubbed = any((getattr(sys.modules.get(m), "__file__", None) or "").startswith(STUB_DIR) for m in ("matplotlib", "shapely"))
        want = [root] + ([STUB_DIR] if stubbed else []) + [REPO]
        have = env.split(os.pathsep)
        if any(p not in have for p in want):
            os.environ["PYTHONPATH"] = os.pathsep.join(want + ([env] if env else []))
        return bayesbay
    if not available():
        raise ImportError("the reference is not installed under baseline/_ref (run __graft_entry__.build() where "
                          "/root/reference exists)")
    write_stubs()
    need_stub = []
    for mod in ("matplotlib", "shapely"):
        try:
            importlib.import_module(mod)
        except ImportError:
            need_stub.append(mod)
    paths = [REF_DIR] + ([STUB_DIR] if need_stub else []) + [REPO]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)
    env = os.environ.get("PYTHONPATH", "")
    os.environ["PYTHONPATH"] = os.pathsep.join(paths + ([env] if env else []))
    import bayesbay

    if not os.path.abspath(bayesbay.__file__).startswith(REF_DIR):
        raise ImportError(f"`bayesbay` resolved to {bayesbay.__file__}, not to baseline/_ref")
    return bayesbay


def build_problem(bayesbay, ing):
    """``(parameterization, log_likelihood)`` of a synthetic problem built with the reference's own classes; the
    forward is the device operator descriptor, which is a valid reference forward (bayesbay_b200/forward.py)."""
    from bayesbay_b200 import synthetic

    return synthetic.build_problem(ing, bayesbay, synthetic.forward_operator)


def run_pool(ing, n_chains: int, n_iterations: int, n_jobs: int, save_every=None):
    """The reference's own chain pool: ``BayesianInversion.run(..., parallel_config={"n_jobs": n_jobs})`` =
    ``joblib.Parallel(backend="loky")`` over pickled ``MarkovChain`` objects (samplers/_samplers.py:191-239).
    Returns ``(chain-steps/s, wall seconds, inversion)``; the construction of the chains (serial, in the parent:
    one forward per chain) is not timed, the pool start-up of the timed call is."""
    bayesbay = import_reference()
    par, ll = build_problem(bayesbay, ing)
    inv = bayesbay.BayesianInversion(par, ll, n_chains=n_chains)
    t0 = time.perf_counter()
    inv.run(n_iterations=n_iterations, burnin_iterations=0, save_every=save_every or max(n_iterations, 1), verbose=False,
            parallel_config={"n_jobs": n_jobs})
    wall = time.perf_counter() - t0
    return n_chains * n_iterations / wall, wall, inv


def timed_pool(ing, cores: int, seconds: float, repeats: int = 1, warm_iterations: int = 30):
    """Bounded timing of the reference's chain pool on `cores` host cores: one chain per core (the reference asks for
    one worker per chain, samplers/_samplers.py:232-235), a warm-up call that starts the loky workers, a calibration
    call, then `repeats` timed calls of about `seconds` each on the same chains.
    Returns ``dict(rates=[chain-steps/s ...], walls=[s ...], n_iterations=N, n_chains=cores)``."""
    bayesbay = import_reference()
    par, ll = build_problem(bayesbay, ing)
    inv = bayesbay.BayesianInversion(par, ll, n_chains=cores)
    cfg = {"n_jobs": cores}

    def call(n):
        t0 = time.perf_counter()
        inv.run(n_iterations=n, burnin_iterations=0, save_every=max(n, 1), verbose=False, parallel_config=cfg)
        return time.perf_counter() - t0

    call(warm_iterations)                     # pool start-up, imports in the workers
    t_cal = call(warm_iterations)             # per-call overhead (pickling the chains out and back) + iterations
    t_cal2 = call(3 * warm_iterations)
    per_it = max((t_cal2 - t_cal) / (2 * warm_iterations), 1e-6)
    overhead = max(t_cal - warm_iterations * per_it, 0.0)
    n = int(max(warm_iterations, (seconds - overhead) / per_it))
    rates, walls = [], []
    for _ in range(repeats):
        w = call(n)
        walls.append(w)
        rates.append(cores * n / w)
    return dict(rates=rates, walls=walls, n_iterations=n, n_chains=cores)
