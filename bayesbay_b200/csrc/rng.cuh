// Per-chain random draws: Philox4x32-10 keyed by (seed), counter (block, chain id, iteration),
// or a recorded tape of uniforms (step-replay parity).  The rules that turn uniforms into the
// reference's draws (random.choices / randint / normalvariate / np.random.*) are the ones
// documented in oracle/rng.py; SURVEY.md quirk Q8 explains why parity is defined on uniforms.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bbb {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

enum { STREAM_STEP = 0, STREAM_INIT = 1, STREAM_SWAP = 2 };

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline double words_to_double(uint32_t w0, uint32_t w1) {
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}

struct Rng {
    // warp-cooperative mode: the 32 lanes of a warp work on ONE chain; lane 0 owns the generator and every
    // draw is broadcast, so all lanes follow the same control flow with the same values
    bool coop = false;
    // philox state
    uint32_t k0, k1, chain, it_lo, it_hi;
    uint32_t n;        // draws consumed in this (chain, iteration)
    uint32_t blk;      // block currently cached in w
    uint32_t w[4];
    // tape state (tape != nullptr selects it)
    const double *tape;
    long long cursor;
    long long tape_len = 0;  // entries of the chain's tape (bounds the look-ahead below)
    // cooperative look-ahead: lane l holds the uniform at stream position l of this (chain, iteration) and the
    // standard normal made of positions (l, l + 1), all computed in parallel up front; draws at positions < 32 are
    // then one shuffle each instead of a serial Philox round / log / sqrt / cos on lane 0
    bool ahead = false;
    double ahead_u = 0.0, ahead_z = 0.0;

    __device__ void init_philox(uint64_t seed, uint64_t chain_id, uint64_t iteration, uint32_t tag) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        chain = (uint32_t)chain_id;
        it_lo = (uint32_t)iteration;
        it_hi = ((uint32_t)(iteration >> 32) & 0x0FFFFFFFu) | (tag << 28);
        n = 0;
        blk = 0xFFFFFFFFu;
        tape = nullptr;
        cursor = 0;
    }
    __device__ void init_tape(const double *t, long long cur) {
        tape = t;
        cursor = cur;
        n = 0;
        blk = 0xFFFFFFFFu;
    }
    __device__ double raw_u() {
        if (tape != nullptr) {
            ++n;
            return tape[cursor++];
        }
        const uint32_t b = n >> 1;
        if (b != blk) {
            philox4x32_10(b, chain, it_lo, it_hi, k0, k1, w);
            blk = b;
        }
        const uint32_t h = (n & 1u) * 2u;
        ++n;
        return words_to_double(w[h], w[h + 1]);
    }
    // call once, by all 32 lanes, right after init_* of a cooperative generator that starts at position 0
    __device__ void look_ahead() {
        const uint32_t lane = threadIdx.x & 31u;
        if (tape != nullptr) {
            ahead_u = (cursor + lane < tape_len) ? tape[cursor + lane] : 0.0;
        } else {
            uint32_t o[4];
            philox4x32_10(lane >> 1, chain, it_lo, it_hi, k0, k1, o);
            const uint32_t h = (lane & 1u) * 2u;
            ahead_u = words_to_double(o[h], o[h + 1]);
        }
        const double u2 = __shfl_down_sync(0xffffffffu, ahead_u, 1);
        ahead_z = sqrt(-2.0 * log(1.0 - ahead_u)) * cos(6.283185307179586 * u2);  // same expression as std_normal()
        ahead = true;
    }
    __device__ double u() {
        if (!coop) return raw_u();
        if (ahead && n < 32u) {
            const double v = __shfl_sync(0xffffffffu, ahead_u, (int)n);
            ++n;
            ++cursor;
            return v;
        }
        double v = 0.0;
        if ((threadIdx.x & 31) == 0) v = raw_u();
        else { ++n; ++cursor; }
        return __shfl_sync(0xffffffffu, v, 0);
    }
    // random.random() restated: (0, 1]
    __device__ double unit() { return 1.0 - u(); }
    __device__ double uniform(double a, double b) { return a + (b - a) * u(); }
    __device__ int randint(int a, int b) {
        if (a == b) return a;
        const int m = b - a + 1;
        int i = (int)(u() * (double)m);
        if (i > m - 1) i = m - 1;
        return a + i;
    }
    // random.choices(range(n), weights)[0]: bisect_right(cumsum(w), u * total), clipped
    __device__ int choice(const double *wts, int m) {
        if (m == 1) return 0;
        double tot = 0.0;
        for (int i = 0; i < m; ++i) tot += wts[i];
        const double x = u() * tot;
        double cum = 0.0;
        int i = 0;
        for (; i < m - 1; ++i) {
            cum += wts[i];
            if (x < cum) break;
        }
        return i;
    }
    __device__ double std_normal() {
        if (coop && ahead && n + 1u < 32u) {
            const double z = __shfl_sync(0xffffffffu, ahead_z, (int)n);
            n += 2u;
            cursor += 2;
            return z;
        }
        const double u1 = u();
        const double u2 = u();
        return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
    }
    __device__ double normal(double mu, double sigma) { return mu + sigma * std_normal(); }
    __device__ double laplace(double mu, double scale) {
        const double v = u();
        if (v >= 0.5) return mu - scale * log(2.0 - v - v);
        if (v > 0.0) return mu + scale * log(v + v);
        return mu;
    }
};

}  // namespace bbb
