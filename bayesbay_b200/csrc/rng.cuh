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

// Out-of-line copies of the long arithmetic sequences: every draw site of the step kernels used to inline its own
// Philox rounds / logarithm / square root / cosine, which made those kernels several hundred KB of code (an
// instruction-cache problem, not an arithmetic one: a proposal is a latency-bound chain on one warp or thread).
__device__ __noinline__ inline uint4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    uint32_t o[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __noinline__ inline double box_muller(double u1, double u2) {
    return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
}
__device__ __noinline__ inline double dlog(double x) { return log(x); }

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
            const uint4 o = philox_block(b, chain, it_lo, it_hi, k0, k1);
            w[0] = o.x; w[1] = o.y; w[2] = o.z; w[3] = o.w;
            blk = b;
        }
        const bool hi = (n & 1u) != 0;
        ++n;
        return hi ? words_to_double(w[2], w[3]) : words_to_double(w[0], w[1]);
    }
    // call once, by all 32 lanes, right after init_* of a cooperative generator that starts at position 0
    __device__ void look_ahead() {
        const uint32_t lane = threadIdx.x & 31u;
        if (tape != nullptr) {
            ahead_u = (cursor + lane < tape_len) ? tape[cursor + lane] : 0.0;
        } else {
            const uint4 o = philox_block(lane >> 1, chain, it_lo, it_hi, k0, k1);
            ahead_u = (lane & 1u) ? words_to_double(o.z, o.w) : words_to_double(o.x, o.y);
        }
        const double u2 = __shfl_down_sync(0xffffffffu, ahead_u, 1);
        ahead_z = box_muller(ahead_u, u2);  // same expression as std_normal()
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
    // (values that enter a chain's state are formed with explicitly unfused operations, like the reference's Python
    // floats: every kernel that draws them -- whatever its inlining context -- then produces the same bits)
    __device__ double uniform(double a, double b) { return __dadd_rn(a, __dmul_rn(b - a, u())); }
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
        return box_muller(u1, u2);
    }
    __device__ double normal(double mu, double sigma) { return __dadd_rn(mu, __dmul_rn(sigma, std_normal())); }
    __device__ double laplace(double mu, double scale) {
        const double v = u();
        if (v >= 0.5) return __dsub_rn(mu, __dmul_rn(scale, dlog(2.0 - v - v)));
        if (v > 0.0) return __dadd_rn(mu, __dmul_rn(scale, dlog(v + v)));
        return mu;
    }
    // tape position (tape-driven) / draws consumed in this (chain, iteration)
    __device__ long long after() const { return tape != nullptr ? cursor : (long long)n; }
};

// The same stream for a warp that draws ONE chain's proposal with all its lanes (proposal kernel): the first 32 uniforms
// of the (chain, iteration) -- and the standard normals made of positions (l, l + 1) -- are computed by the 32 lanes at
// once into two shared-memory tables; a draw is then one shared-memory read that every lane makes alike (no shuffles,
// no lane-0 ownership), which keeps the draw sites of the proposal code a few instructions long.  Positions past the
// tables (long resampling loops) are computed on demand by every lane redundantly.
__device__ __noinline__ inline double stream_u(const double *tape, long long cursor0, long long tape_len, uint32_t k0, uint32_t k1,
                                               uint32_t chain, uint32_t it_lo, uint32_t it_hi, int pos) {
    if (tape != nullptr) return cursor0 + pos < tape_len ? tape[cursor0 + pos] : 0.0;
    const uint4 o = philox_block((uint32_t)pos >> 1, chain, it_lo, it_hi, k0, k1);
    return (pos & 1) ? words_to_double(o.z, o.w) : words_to_double(o.x, o.y);
}
struct FastRng {
    static constexpr bool coop = true;
    const double *su, *sz;
    const double *tape;
    long long cursor0, tape_len;
    uint32_t k0, k1, chain, it_lo, it_hi;
    int n;
    // all 32 lanes; su / sz: 32 doubles each, private to the calling warp
    __device__ void begin(double *su_w, double *sz_w, const double *tape_c, long long cursor, long long tape_length, uint64_t seed,
                          uint64_t chain_id, uint64_t iteration, uint32_t tag) {
        su = su_w; sz = sz_w;
        tape = tape_c; cursor0 = cursor; tape_len = tape_length;
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        chain = (uint32_t)chain_id;
        it_lo = (uint32_t)iteration;
        it_hi = ((uint32_t)(iteration >> 32) & 0x0FFFFFFFu) | (tag << 28);
        n = 0;
        const int lane = threadIdx.x & 31;
        const double ul = stream_u(tape, cursor0, tape_len, k0, k1, chain, it_lo, it_hi, lane);
        su_w[lane] = ul;
        __syncwarp();
        sz_w[lane] = lane < 31 ? box_muller(ul, su_w[lane + 1]) : 0.0;
        __syncwarp();
    }
    __device__ __forceinline__ double u() {
        const int i = n++;
        return i < 32 ? su[i] : stream_u(tape, cursor0, tape_len, k0, k1, chain, it_lo, it_hi, i);
    }
    __device__ __forceinline__ long long after() const { return tape != nullptr ? cursor0 + n : (long long)n; }
    __device__ __forceinline__ double unit() { return 1.0 - u(); }
    __device__ __forceinline__ double uniform(double a, double b) { return __dadd_rn(a, __dmul_rn(b - a, u())); }
    __device__ __forceinline__ int randint(int a, int b) {
        if (a == b) return a;
        const int m = b - a + 1;
        int i = (int)(u() * (double)m);
        if (i > m - 1) i = m - 1;
        return a + i;
    }
    __device__ __forceinline__ int choice(const double *wts, int m) {
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
    __device__ __forceinline__ double std_normal() {
        const int i = n;
        if (i + 1 < 32) {
            n += 2;
            return sz[i];
        }
        const double u1 = u();
        const double u2 = u();
        return box_muller(u1, u2);
    }
    __device__ __forceinline__ double normal(double mu, double sigma) { return __dadd_rn(mu, __dmul_rn(sigma, std_normal())); }
    __device__ __forceinline__ double laplace(double mu, double scale) {
        const double v = u();
        if (v >= 0.5) return __dsub_rn(mu, __dmul_rn(scale, dlog(2.0 - v - v)));
        if (v > 0.0) return __dadd_rn(mu, __dmul_rn(scale, dlog(v + v)));
        return mu;
    }
};

}  // namespace bbb
