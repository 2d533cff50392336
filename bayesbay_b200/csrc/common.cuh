// Shared device-side helpers: SoA views over the caller-owned buffers, prior formulas.
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/bbb200.h"
#include "rng.cuh"

namespace bbb {

constexpr double LOG_TWO_PI_HALF = 0.91893853320467274178;  // 0.5*log(2*pi)
constexpr double SQRT_TWO_PI = 2.50662827463100050242;
constexpr int RAY_TILE = 64;       // rays per SpMM block (partial-misfit granularity)
constexpr int CHAIN_TILE = 32;     // chains per warp-lane tile

// SimulatedAnnealing.on_begin_iteration (samplers/_samplers.py:402-414): temperature of iteration `it`
struct Annealing {
    int enabled;
    double t0, rate;
    long long cooling_iterations, burnin_iterations;
};

struct EngineParams {
    bbb_problem_spec spec;
    bbb_buffers buf;
    Annealing anneal;
    int n_partial[BBB_MAX_TARGETS];  // rows of partial[t] the ray kernel writes (its ray groups)
    int chain_local;                 // 1: the chain-local kernel (chain_kernels.cu) evaluates the ray target
    int ray_t;                       // chain-local engines: the one ray target (else -1)
};

// Longest-first block order of the chain kernel.  The proposal kernel files every chain of a launch range
// [c_first, c_end) under a cost class (atomic counter per class), the chain kernel's block b takes the b-th chain in
// class order.  Class counts are double-buffered by launch parity: a chain kernel zeroes the set the NEXT proposal
// kernel of its part will count into.  Whatever the order, results are the same (chains never interact in a step).
__host__ __device__ inline int32_t *order_counts(int32_t *block_order, long long C, int parity, int part) {
    return block_order + (long long)BBB_ORDER_BUCKETS * C + (parity * 4 + part) * 16;
}

__device__ __forceinline__ double annealing_temperature(const Annealing &a, long long it) {
    if (it < a.burnin_iterations && it < a.cooling_iterations) {
        const double t = a.t0 * exp(-a.rate * (double)it);
        return t > 1.0 ? t : 1.0;  // guard against rounding below 1
    }
    return 1.0;
}

// ---- slot-addressed views (arrays are [2][...][K][C], chain fastest) -------------------------
__device__ __forceinline__ double &site_ref(const bbb_buffers &B, const bbb_space_spec &S, int s,
                                            int slot, int d, int j, long long c) {
    return B.sites[s][(((long long)slot * S.ndim + d) * S.kmax + j) * B.n_chains + c];
}
__device__ __forceinline__ double &value_ref(const bbb_buffers &B, const bbb_space_spec &S, int s,
                                             int slot, int p, int j, long long c) {
    return B.values[s][(((long long)slot * S.n_params + p) * S.kmax + j) * B.n_chains + c];
}
__device__ __forceinline__ int32_t &k_ref(const bbb_buffers &B, int s, int slot, long long c) {
    return B.k[s][(long long)slot * B.n_chains + c];
}

// ---- priors (prior/_prior.py: Uniform :400-481, Gaussian :599-682, Laplace :800-881) ----------
// (out of line, like the generator's arithmetic in rng.cuh: plain values in, one value out)
__device__ __noinline__ inline double prior_log_pdf(int kind, double a, double b, double v) {
    if (kind == BBB_PRIOR_UNIFORM) return (a <= v && v <= b) ? -log(b - a) : -INFINITY;
    if (kind == BBB_PRIOR_GAUSSIAN) {
        const double z = (v - a) / b;
        return -LOG_TWO_PI_HALF - log(b) - 0.5 * (z * z);
    }
    return -0.69314718055994530942 - log(b) - fabs(v - a) / b;
}
__device__ __forceinline__ double prior_value_ratio(int kind, double a, double b, double old_v,
                                                    double new_v) {
    if (kind == BBB_PRIOR_UNIFORM) return (a <= new_v && new_v <= b) ? 0.0 : -INFINITY;
    if (kind == BBB_PRIOR_GAUSSIAN) return (old_v - new_v) * (old_v + new_v - 2.0 * a) / (2.0 * (b * b));
    return (fabs(old_v - a) - fabs(new_v - a)) / b;
}
template <class R>
__device__ __forceinline__ double prior_sample(int kind, double a, double b, R &rng) {
    if (kind == BBB_PRIOR_UNIFORM) return rng.uniform(a, b);
    if (kind == BBB_PRIOR_GAUSSIAN) return rng.normal(a, b);
    return rng.laplace(a, b);
}

// Hyper-parameters of parameter p at site position x: Prior.get_hyper_param (prior/_prior.py:257-297) with
// interpolate_linear_1d (_utils_1d.pyx:68-84) for position-dependent ones, the scalars otherwise.
struct PriorAt {
    int kind;
    double a, b, pstd, pstd_birth;
};
// position-dependent part (the table lives in global memory)
__device__ __noinline__ inline void prior_at_table(const double *X, int n, double x, double &a, double &b, double &pstd, double &pstd_birth) {
    int i0, i1;
    if (x <= X[0]) { i0 = i1 = 0; }
    else if (x >= X[n - 1]) { i0 = i1 = n - 1; }
    else {
        i0 = 0;
        while (!(X[i0] <= x && x <= X[i0 + 1])) ++i0;
        i1 = i0 + 1;
    }
    auto at = [&](int row) {
        const double y0 = X[(long long)row * n + i0];
        if (i1 == i0) return y0;
        return y0 + (x - X[i0]) * (X[(long long)row * n + i1] - y0) / (X[i1] - X[i0]);
    };
    a = at(1); b = at(2); pstd = at(3); pstd_birth = at(4);
}
__device__ __forceinline__ PriorAt prior_at(const bbb_space_spec &S, int p, double x) {
    PriorAt out{S.prior_kind[p], S.prior_a[p], S.prior_b[p], S.perturb_std[p], S.perturb_std_birth[p]};
    const int n = S.n_pos[p];
    if (n > 0) prior_at_table(S.pos_table[p], n, x, out.a, out.b, out.pstd, out.pstd_birth);
    return out;
}
__device__ __forceinline__ double prior_log_pdf(const PriorAt &q, double v) { return prior_log_pdf(q.kind, q.a, q.b, v); }
__device__ __forceinline__ double prior_value_ratio(const PriorAt &q, double o, double n) { return prior_value_ratio(q.kind, q.a, q.b, o, n); }
template <class R>
__device__ __forceinline__ double prior_sample(const PriorAt &q, R &rng) { return prior_sample(q.kind, q.a, q.b, rng); }

// Targets whose misfit is not a plain weighted sum of squares (AR(1) noise, full inverse covariance matrix): their
// residual vector is materialised per chain in dpred_scratch[t] ([n_data][C]) and reduced from there.
__host__ __device__ __forceinline__ bool reduces_dpred(const bbb_target_spec &T) {
    return T.correlated != 0 || T.cov_kind == BBB_COV_MATRIX;
}

__device__ __forceinline__ int n_move_types(const bbb_problem_spec &P) { return 4 * P.n_spaces + 1; }

}  // namespace bbb
