// F3/F4 per-chain forwards, L1/L2 likelihood ratio and the Metropolis-Hastings decision (P0).
#pragma once
#include "common.cuh"

namespace bbb {

__device__ __forceinline__ double data_weight(const bbb_target_spec &T, int r) {
    return (T.cov_kind == BBB_COV_VECTOR) ? T.cov_vec[r] : 1.0;
}

// 1-D nearest-nuclei cell of xp: interpolate_nearest_1d + nearest_neighbour_1d
// (_utils_1d.pyx:87-114): xp <= x[0] -> 0, xp >= x[k-1] -> k-1, else the first i with
// x[i] <= xp <= x[i+1], tie -> upper.  `hint` is a monotone sweep pointer (data sorted ascending);
// `at(j)` returns site j of the chain.
template <typename At>
__device__ __forceinline__ int cell_1d_at(At at, int k, double xp, int &hint) {
    if (xp <= at(0)) return 0;
    if (xp >= at(k - 1)) return k - 1;
    int i = hint;
    if (i > k - 2 || at(i) > xp) i = 0;  // data not ascending: restart the sweep
    while (i < k - 2 && at(i + 1) < xp) ++i;
    hint = i;
    const double x0 = at(i);
    const double x1 = at(i + 1);
    return (fabs(x0 - xp) < fabs(x1 - xp)) ? i : i + 1;
}
__device__ __forceinline__ int cell_1d(const bbb_buffers &B, const bbb_space_spec &S, int s, int slot,
                                       long long c, int k, double xp, int &hint) {
    return cell_1d_at([&](int j) { return site_ref(B, S, s, slot, 0, j, c); }, k, xp, hint);
}

// Residual sums of one target with the noise parameters factored out (so noise moves need no forward):
//   s0 = sum_r w_r res_r^2,   se = res_0^2 + res_{n-1}^2,   s1 = sum_r res_r res_{r+1}
// For correlated noise  res^T C^-1 res = [(1 + rho^2) s0 - rho^2 se - 2 rho s1] / (sigma^2 (1 - rho^2))  with the
// tridiagonal C^-1 of inverse_covariance (_utils_1d.pyx:198-211).
struct Phi3 {
    double s0, se, s1;
};
__device__ __forceinline__ Phi3 load_phi(const bbb_buffers &B, int t, int slot, long long c) {
    const long long C = B.n_chains;
    const double *p = B.phi[t] + (long long)slot * 3 * C + c;
    return Phi3{p[0], p[C], p[2 * C]};
}
__device__ __forceinline__ void store_phi(const bbb_buffers &B, int t, int slot, long long c, const Phi3 &v) {
    const long long C = B.n_chains;
    double *p = B.phi[t] + (long long)slot * 3 * C + c;
    p[0] = v.s0; p[C] = v.se; p[2 * C] = v.s1;
}
// Full inverse covariance matrix (BBB_COV_MATRIX): s0 = res^T M res = sum_i res_i (M res)_i, the order of
// `residual @ (covariance_mat_inv @ residual)` (likelihood/_log_likelihood.py:293-301, _target.py:150-151); `res` is
// the chain's column of a [n][C] scratch, M (row-major) is read by every chain alike (broadcast loads).
__device__ __noinline__ inline double matrix_quadratic_form(const double *M, const double *res, long long stride, int n) {
    double tot = 0.0;
    for (int i = 0; i < n; ++i) {
        double mv = 0.0;
        const double *row = M + (long long)i * n;
        for (int j = 0; j < n; ++j) mv += row[j] * res[(long long)j * stride];
        tot += res[(long long)i * stride] * mv;
    }
    return tot;
}
struct ResidualSums {  // streaming accumulation of Phi3 over r = 0 .. n-1
    Phi3 v{0.0, 0.0, 0.0};
    double prev = 0.0;
    double *keep = nullptr;  // BBB_COV_MATRIX: the residuals are kept (stride `stride`) for matrix_quadratic_form
    long long stride = 0;
    __device__ __forceinline__ void add(int r, int n, double res, double w) {
        if (keep) keep[(long long)r * stride] = res;
        v.s0 += w * (res * res);
        if (r == 0 || r == n - 1) v.se += res * res;
        if (r > 0) v.s1 += prev * res;
        prev = res;
    }
};

// Residual sums of a thread-evaluated target; optionally stores d_pred (stride C) and the cell index
// (lookup1d, stride C).
__device__ __noinline__ inline Phi3 small_target_phi(const bbb_problem_spec &P, const bbb_buffers &B, int t, int slot,
                                                     long long c, double *dpred_out, int32_t *idx_out) {
    const bbb_target_spec &T = P.targets[t];
    const int s = T.space;
    const bbb_space_spec &S = P.spaces[s];
    const long long C = B.n_chains;
    ResidualSums acc;
    const bool full_matrix = T.cov_kind == BBB_COV_MATRIX;
    if (full_matrix) { acc.keep = B.dpred_scratch[t] + c; acc.stride = C; }
    if (T.fwd_kind == BBB_FWD_LOOKUP1D) {
        // Voronoi1D.interpolate_tessellation(..., 'nuclei') (discretization/_voronoi.py:705-732)
        const int k = k_ref(B, s, slot, c);
        int hint = 0;
        for (int r = 0; r < T.n_data; ++r) {
            const int i = cell_1d(B, S, s, slot, c, k, T.x[r], hint);
            const double d = value_ref(B, S, s, slot, T.param, i, c);
            if (dpred_out) dpred_out[(long long)r * C + c] = d;
            if (idx_out) idx_out[(long long)r * C + c] = i;
            acc.add(r, T.n_data, d - T.dobs[r], data_weight(T, r));
        }
    } else if (T.fwd_kind == BBB_FWD_POLYNOMIAL) {
        // np.vander(x, k, increasing=True) @ m (03_transd_polyfit.ipynb cell 5): powers by repeated multiplication
        const int k = k_ref(B, s, slot, c);
        for (int r = 0; r < T.n_data; ++r) {
            const double xr = T.x[r];
            double pw = 1.0, d = 0.0;
            for (int j = 0; j < k; ++j) {
                d += value_ref(B, S, s, slot, T.param, j, c) * pw;
                pw *= xr;
            }
            if (dpred_out) dpred_out[(long long)r * C + c] = d;
            acc.add(r, T.n_data, d - T.dobs[r], data_weight(T, r));
        }
    } else if (T.fwd_kind == BBB_FWD_LINEAR) {
        // fwd_operator @ m (02_hierarchical_polyfit.ipynb cell 12)
        const int K = S.kmax;
        const int M = T.n_lin_params * K;
        for (int r = 0; r < T.n_data; ++r) {
            double d = 0.0;
            for (int q = 0; q < T.n_lin_params; ++q)
                for (int j = 0; j < K; ++j)
                    d += T.G[(long long)r * M + q * K + j] * value_ref(B, S, s, slot, T.lin_params[q], j, c);
            if (dpred_out) dpred_out[(long long)r * C + c] = d;
            acc.add(r, T.n_data, d - T.dobs[r], data_weight(T, r));
        }
    }
    if (full_matrix) acc.v.s0 = matrix_quadratic_form(T.cov_vec, acc.keep, C, T.n_data);
    return acc.v;
}

__device__ __forceinline__ double ray_partial_sum(const EngineParams &ep, int t, long long c) {
    const bbb_buffers &B = ep.buf;
    const int n_tiles = ep.n_partial[t];
    double phi = 0.0;
    for (int i = 0; i < n_tiles; ++i) phi += B.partial[t][(long long)i * B.n_chains + c];
    return phi;
}

// Target.inverse_covariance_times_vector / log_determinant_covariance (likelihood/_target.py:136-208)
// (out of line, plain values in and out: see rng.cuh)
struct MisfitLogdet {
    double misfit, logdet;
};
__device__ __noinline__ inline MisfitLogdet misfit_logdet_values(int cov_kind, int correlated, int n_data, double cov_scalar, double s0,
                                                                  double se, double s1, double sigma, double rho) {
    MisfitLogdet out;
    if (cov_kind == BBB_COV_HIERARCHICAL) {
        if (correlated) {
            const double factor = 1.0 / ((sigma * sigma) * (1.0 - rho * rho));
            out.misfit = factor * ((1.0 + rho * rho) * s0 - (rho * rho) * se - 2.0 * rho * s1);
            out.logdet = (2.0 * n_data) * log(sigma) + (double)(n_data - 1) * log(1.0 - rho * rho);
        } else {
            out.misfit = (1.0 / (sigma * sigma)) * s0;
            out.logdet = (2.0 * n_data) * log(sigma);
        }
    } else if (cov_kind == BBB_COV_SCALAR) {
        out.misfit = cov_scalar * s0;
        out.logdet = 0.0;
    } else {
        out.misfit = s0;
        out.logdet = 0.0;
    }
    return out;
}
__device__ __forceinline__ void misfit_logdet_terms(const bbb_target_spec &T, const Phi3 &phi, double sigma, double rho,
                                                    double &misfit, double &logdet) {
    const MisfitLogdet v = misfit_logdet_values(T.cov_kind, T.correlated, T.n_data, T.cov_scalar, phi.s0, phi.se, phi.s1, sigma, rho);
    misfit = v.misfit;
    logdet = v.logdet;
}

// True when the current proposal of chain c requires target t's forward to be re-evaluated.
__device__ __forceinline__ bool target_touched(const bbb_problem_spec &P, int move, double lpr, int t) {
    if (isinf(lpr)) return false;                 // forward skipped, _markov_chain.py:222-223
    if (move == 4 * P.n_spaces) return false;     // noise move: d_pred unchanged (quirk Q1)
    return P.targets[t].space == (move >> 2);
}
__device__ __forceinline__ bool move_is_geometry(int move) {
    const int inner = move & 3;
    return inner != BBB_MOVE_VALUES;
}

// Generator of chain c for the current iteration (Philox counters or the recorded tape); `continue_step`
// resumes after the draws the proposal kernel consumed.
__device__ __forceinline__ void rng_begin(Rng &rng, const EngineParams &ep, long long c, uint32_t tag,
                                          bool continue_step) {
    const bbb_buffers &B = ep.buf;
    if (B.tape != nullptr) {
        rng.init_tape(B.tape + c * B.tape_stride, B.tape_cursor[c]);
        rng.tape_len = B.tape_stride;
    } else {
        rng.init_philox(ep.spec.seed, (uint64_t)(B.chain_id0 + c), (uint64_t)B.iteration[c], tag);
        if (continue_step) rng.n = (uint32_t)B.tape_cursor[c];
    }
}
__device__ __forceinline__ void rng_end(const Rng &rng, const EngineParams &ep, long long c) {
    const bbb_buffers &B = ep.buf;
    B.tape_cursor[c] = (B.tape != nullptr) ? rng.cursor : (long long)rng.n;
}

// Everything the decision reads from global memory, gathered in one go (the chain-local kernel does this at its
// start from an otherwise idle thread, so the decision at its end is arithmetic and stores only).
struct FinishIn {
    int move, sel_s;
    double lpr, temperature;
    long long iteration, cursor, n_prop, n_acc;
    Phi3 phi_old[BBB_MAX_TARGETS];
    double sig[BBB_MAX_TARGETS][2], rho[BBB_MAX_TARGETS][2];  // [current, proposed]
};

// (the caller that drew the proposal itself passes what it already holds -- move, ratio, generator position -- so
// that nothing here depends on a load)
__device__ inline void finish_gather_rest(const EngineParams &ep, long long c, FinishIn &in);
__device__ inline void finish_gather(const EngineParams &ep, long long c, FinishIn &in) {
    const bbb_buffers &B = ep.buf;
    in.move = B.move[c];
    in.lpr = B.log_prob_ratio[c];
    in.cursor = B.tape_cursor[c];
    finish_gather_rest(ep, c, in);
}
__device__ inline void finish_gather_rest(const EngineParams &ep, long long c, FinishIn &in) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    in.iteration = B.iteration[c];
    in.temperature = ep.anneal.enabled ? annealing_temperature(ep.anneal, in.iteration) : B.temperature[c];
    const bool noise = (in.move == 4 * P.n_spaces);
    in.sel_s = noise ? 0 : (int)B.sel[in.move >> 2][c];
    in.n_prop = B.n_proposed[(long long)in.move * C + c];
    in.n_acc = B.n_accepted[(long long)in.move * C + c];
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        const bool hier = (T.cov_kind == BBB_COV_HIERARCHICAL);
        in.phi_old[t] = T.correlated ? load_phi(B, t, 0, c) : Phi3{B.phi[t][c], 0.0, 0.0};  // (only AR(1) noise reads se, s1)
        in.sig[t][0] = hier ? B.sigma[t][c] : 1.0;
        in.sig[t][1] = (hier && noise) ? B.sigma[t][C + c] : in.sig[t][0];
        in.rho[t][0] = T.correlated ? B.corr[t][c] : 0.0;
        in.rho[t][1] = (T.correlated && noise) ? B.corr[t][C + c] : in.rho[t][0];
    }
}

// Generator continuing the chain's step from gathered counters (same streams as rng_begin(..., true)).
__device__ __forceinline__ void rng_resume(Rng &rng, const EngineParams &ep, long long c, const FinishIn &in) {
    const bbb_buffers &B = ep.buf;
    if (B.tape != nullptr) {
        rng.init_tape(B.tape + c * B.tape_stride, in.cursor);
    } else {
        rng.init_philox(ep.spec.seed, (uint64_t)(B.chain_id0 + c), (uint64_t)in.iteration, STREAM_STEP);
        rng.n = (uint32_t)in.cursor;
    }
}

// The decision in two parts, so that the chain-local kernel can run the long part (logarithms, the generator,
// the other targets) on an idle thread WHILE the forward of its ray target is being evaluated:
//   finish_pre   everything that does not depend on the new misfit of the deferred ray target `ray_t`
//                (ray_t = -1: nothing is deferred),
//   finish_post  that target's misfit term, the sums in target order, the Metropolis-Hastings test, commit,
//                statistics.
// Together they evaluate exactly the expressions of _next_iteration (_markov_chain.py:222-243) in the same order.
struct FinishPre {
    double log_u, factor_ray;
    double m0[BBB_MAX_TARGETS], m1[BBB_MAX_TARGETS], l0[BBB_MAX_TARGETS], l1[BBB_MAX_TARGETS];
    Phi3 phi_new[BBB_MAX_TARGETS];
};

__device__ inline void finish_pre(const EngineParams &ep, long long c, const FinishIn &in, Rng &rng, int ray_t,
                                  FinishPre &pre) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const bool noise = (in.move == 4 * P.n_spaces);
    const int s = in.move >> 2;
    pre.factor_ray = 1.0;
    if (!isinf(in.lpr)) {
        for (int t = 0; t < P.n_targets; ++t) {
            const bbb_target_spec &T = P.targets[t];
            const Phi3 phi_old = in.phi_old[t];
            pre.phi_new[t] = phi_old;
            if (!noise && T.space == s && t != ray_t) {
                if (T.fwd_kind != BBB_FWD_RAY2D) pre.phi_new[t] = small_target_phi(P, B, t, in.sel_s ^ 1, c, nullptr, nullptr);
                else if (reduces_dpred(T)) pre.phi_new[t] = load_phi(B, t, 1, c);  // k_corr_sums
                else pre.phi_new[t] = Phi3{ray_partial_sum(ep, t, c), 0.0, 0.0};
                store_phi(B, t, 1, c, pre.phi_new[t]);
            }
            misfit_logdet_terms(T, phi_old, in.sig[t][0], in.rho[t][0], pre.m0[t], pre.l0[t]);
            misfit_logdet_terms(T, pre.phi_new[t], in.sig[t][1], in.rho[t][1], pre.m1[t], pre.l1[t]);
            if (t == ray_t)  // the factor misfit_logdet_terms applies to phi.s0 (the deferred target is uncorrelated)
                pre.factor_ray = T.cov_kind == BBB_COV_HIERARCHICAL ? 1.0 / (in.sig[t][1] * in.sig[t][1])
                                                                    : (T.cov_kind == BBB_COV_SCALAR ? T.cov_scalar : 1.0);
        }
    }
    pre.log_u = dlog(rng.unit());
}

__device__ inline bool finish_post(const EngineParams &ep, long long c, const FinishIn &in, FinishPre &pre, int ray_t,
                                   double ray_phi_new) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    if (ep.anneal.enabled) B.temperature[c] = in.temperature;
    const int move = in.move;
    const double lpr = in.lpr;
    const bool noise = (move == 4 * P.n_spaces);
    const int s = move >> 2;
    double llr = 0.0;
    if (!isinf(lpr)) {
        if (ray_t >= 0 && !noise && P.targets[ray_t].space == s) {
            pre.phi_new[ray_t] = Phi3{ray_phi_new, 0.0, 0.0};
            store_phi(B, ray_t, 1, c, pre.phi_new[ray_t]);
            pre.m1[ray_t] = P.targets[ray_t].cov_kind == BBB_COV_VECTOR ? ray_phi_new : pre.factor_ray * ray_phi_new;
        }
        double old_m = 0.0, new_m = 0.0, old_ld = 0.0, new_ld = 0.0;
        for (int t = 0; t < P.n_targets; ++t) {
            old_m += pre.m0[t]; new_m += pre.m1[t];
            if (P.targets[t].cov_kind == BBB_COV_HIERARCHICAL) { old_ld += pre.l0[t]; new_ld += pre.l1[t]; }
        }
        // _log_likelihood_ratio_from_targets (likelihood/_log_likelihood.py:245-251), then / T (:243)
        llr = (old_ld - new_ld + old_m - new_m) / 2.0 / in.temperature;
    }
    const bool accepted = (lpr + llr) > pre.log_u;  // NaN compares false -> rejected
    // what the reference would surface as a ForwardException / a NaN misfit: counted, the proposal is rejected
    if (!(fabs(llr) <= 1.79769313486231570815e308) && B.n_exceptions != nullptr) B.n_exceptions[c] += 1;
    if (accepted) {  // (an accepted proposal has a finite lpr, so phi_new is set)
        if (noise) {
            for (int t = 0; t < P.n_targets; ++t) {
                if (P.targets[t].cov_kind == BBB_COV_HIERARCHICAL) B.sigma[t][c] = in.sig[t][1];
                if (P.targets[t].correlated) B.corr[t][c] = in.rho[t][1];
            }
        } else {
            B.sel[s][c] = (uint8_t)(in.sel_s ^ 1);
            for (int t = 0; t < P.n_targets; ++t)
                if (P.targets[t].space == s) store_phi(B, t, 0, c, pre.phi_new[t]);
        }
    }
    // nearest-site index: the proposed index (slot 1) becomes current when a geometry move is accepted; the
    // rasteriser commits it to slot 0 at the start of the next step (tomo_kernels.cu)
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        if (T.fwd_kind != BBB_FWD_RAY2D || ep.chain_local) continue;
        const bool geom_accepted = accepted && !noise && T.space == s && move_is_geometry(move) && !isinf(lpr);
        B.idx_sel[t][c] = geom_accepted ? 1 : 0;  // any older pending commit was done by this step's rasteriser
    }
    B.log_like_ratio[c] = llr;
    B.accepted[c] = accepted ? 1 : 0;
    B.n_proposed[(long long)move * C + c] = in.n_prop + 1;
    B.n_accepted[(long long)move * C + c] = in.n_acc + (accepted ? 1 : 0);
    B.iteration[c] = in.iteration + 1;
    return accepted;
}

__device__ inline bool finish_chain(const EngineParams &ep, long long c, Rng &rng) {
    FinishIn in;
    FinishPre pre;
    finish_gather(ep, c, in);
    finish_pre(ep, c, in, rng, -1, pre);
    return finish_post(ep, c, in, pre, -1, 0.0);
}

}  // namespace bbb
