// P0-P7: move choice and proposal for one chain (one thread per chain, lanes = chains).
//
// Every reference move (perturbations/*.py, discretization/_voronoi.py, parameterization/
// _parameter_space.py) is reduced to one "edit script" on the padded cell arrays:
//     remove cell `rm` (or none), insert a cell (site, values) at output position `ins` (or none)
// so the O(K) part -- writing the proposed slot -- is one loop that every lane of a warp runs in
// lock step, whatever move each chain drew:
//     value move      rm = ins = idx, same site, new values
//     site move 2-D   rm = ins = idx, new site, same values
//     site move 1-D   rm = idx, ins = sorted position of the new site (Voronoi1D.perturb_value
//                     re-sorts with argsort, discretization/_voronoi.py:573-591)
//     birth 2-D       ins = k (np.vstack/np.append, _voronoi.py:330-337)          quirk Q10
//     birth 1-D       ins = bisect_left(sites, new) (_voronoi.py:601-609)
//     birth plain     ins = randint(0, k) (parameterization/_parameter_space.py:197-275)
//     death           rm = randint(0, k-1) (np.delete, _voronoi.py:376-385)
#pragma once
#include "common.cuh"

namespace bbb {

constexpr int MAX_RESAMPLE = 100000;  // the reference loops forever; see DESIGN.md

struct Edit {
    int rm, ins;
    double site[2];
    double vals[BBB_MAX_PARAMS];
    double lpr;
    bool write;
    bool gave_up;  // a rejection-sampling loop hit MAX_RESAMPLE (the old value was kept)
    // filled by propose_chain for callers that go on with the proposal in the same kernel: the move id, the slot of the
    // current state and the cell counts of the current / proposed state (space moves only)
    int move, cur, k_old, k_new;
};

// Where a proposal reads the cells of the CURRENT state from: the [..][K][C] arrays in global memory.  (An accessor
// type so that the same move code can draw from other representations of the cells.)
struct GlobalCells {
    const bbb_buffers &B;
    const bbb_space_spec &S;
    int s, slot;
    long long c;
    __device__ __forceinline__ double site(int d, int j) const { return site_ref(B, S, s, slot, d, j, c); }
    __device__ __forceinline__ double value(int p, int j) const { return value_ref(B, S, s, slot, p, j, c); }
};
// nearest_neighbour_1d (_utils_1d.pyx:102-114) over the cells of the state, optionally skipping `skip`
template <class Cells>
__device__ inline int nearest_1d(const Cells &cells, int k, int skip, double xp) {
    auto at = [&](int j) { return cells.site(0, j + (skip >= 0 && j >= skip ? 1 : 0)); };
    const int n = k - (skip >= 0 ? 1 : 0);
    if (n == 1 || xp < at(0)) return 0;
    for (int i = 0; i < n - 1; ++i) {
        const double x0 = at(i), x1 = at(i + 1);
        if (x0 <= xp && xp <= x1) return (fabs(x0 - xp) < fabs(x1 - xp)) ? i : i + 1;
    }
    return n - 1;
}

// Voronoi.nearest_neighbour (discretization/_voronoi.py:462-465): argmin of the Euclidean distance,
// lowest index on ties; squared distance with separately rounded products (no FMA).
template <class Cells>
__device__ inline int nearest_2d(const Cells &cells, int k, int skip, double px, double py, bool coop) {
    double best = INFINITY;
    int bi = 0x7fffffff;
    if (!coop) {
        int out = 0;
        bi = 0;
        for (int j = 0; j < k; ++j) {
            if (j == skip) continue;
            const double dx = cells.site(0, j) - px;
            const double dy = cells.site(1, j) - py;
            const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d < best) { best = d; bi = out; }
            ++out;
        }
        return bi;  // index among the remaining cells
    }
    // warp-cooperative: lanes = sites, then a lexicographic (distance, index) shuffle reduction
    for (int j = threadIdx.x & 31; j < k; j += 32) {
        if (j == skip) continue;
        const double dx = cells.site(0, j) - px;
        const double dy = cells.site(1, j) - py;
        const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d < best) { best = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < best || (od == best && oj < bi)) { best = od; bi = oj; }
    }
    return bi - ((skip >= 0 && bi > skip) ? 1 : 0);  // index among the remaining cells
}

// ---- outer choice (_markov_chain.py:208-210), then the inner choice of ParamSpacePerturbation
// (.perturb_param_space_state, perturbations/_param_space.py:79-92).  Returns the move id: 4 * space + kind, or
// 4 * n_spaces for the noise perturbation.
// (the weights are walked twice -- total, then cumulative sums in the same order -- instead of being copied into a
// local array first: dynamically indexed local arrays live in local memory, an L2 round trip per access when the L1 is
// as small as the chain kernels leave it.  Same draws and same sums as R::choice: one uniform when there are at least
// two candidates, bisect_right(cumsum(w), u * total) clipped to the last.)
template <class R>
__device__ inline int choose_move(const bbb_problem_spec &P, R &rng) {
    auto outer_w = [&](int s) {
        const bbb_space_spec &S = P.spaces[s];
        return S.w_outer > 0.0 ? S.w_outer : S.w_birth + S.w_death + S.w_values + S.w_sites;
    };
    bool any_hier = false;
    for (int t = 0; t < P.n_targets; ++t) any_hier |= (P.targets[t].cov_kind == BBB_COV_HIERARCHICAL);
    const bool with_noise = any_hier && P.w_noise != 0.0;
    const double w_noise = P.w_noise > 0.0 ? P.w_noise : 1.0;
    const int n_outer = P.n_spaces + (with_noise ? 1 : 0);
    int i_outer = 0;
    if (n_outer > 1) {
        double tot = 0.0;
        for (int s = 0; s < P.n_spaces; ++s) tot += outer_w(s);
        if (with_noise) tot += w_noise;
        const double x = rng.u() * tot;
        double cum = 0.0;
        for (; i_outer < n_outer - 1; ++i_outer) {
            cum += i_outer < P.n_spaces ? outer_w(i_outer) : w_noise;
            if (x < cum) break;
        }
    }
    if (i_outer >= P.n_spaces) return 4 * P.n_spaces;
    const bbb_space_spec &S = P.spaces[i_outer];
    // candidates in the order birth, death, values, sites; those with a zero weight are not candidates
    const double wk[4] = {S.w_birth > 0 ? S.w_birth : 0.0, S.w_death > 0 ? S.w_death : 0.0, S.w_values > 0 ? S.w_values : 0.0,
                          S.w_sites > 0 ? S.w_sites : 0.0};
    const int n_inner = (wk[0] > 0 ? 1 : 0) + (wk[1] > 0 ? 1 : 0) + (wk[2] > 0 ? 1 : 0) + (wk[3] > 0 ? 1 : 0);
    int last = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (wk[q] > 0) last = q;
    if (n_inner == 1) return 4 * i_outer + last;
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (wk[q] > 0) tot += wk[q];
    const double x = rng.u() * tot;
    double cum = 0.0;
    int kind = last;
    bool found = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (wk[q] > 0 && q != last && !found) {
            cum += wk[q];
            if (x < cum) { kind = q; found = true; }
        }
    }
    return 4 * i_outer + kind;
}

// ---- P7 NoisePerturbation.perturb (perturbations/_data_noise.py:21-77) of one hierarchical target: new std (and
// correlation: to_be_perturbed = ["std", "correlation"], :61-76), each resampled until inside its range
// (returns true when a loop gave up)
template <class R>
__device__ inline bool propose_noise_target(const bbb_target_spec &T, R &rng, double old_v, double old_r, double &new_v,
                                            double &new_r) {
    bool gave_up = true;
    new_v = old_v;
    for (int it = 0; it < MAX_RESAMPLE; ++it) {
        const double cand = old_v + rng.normal(0.0, T.std_perturb_std);
        if (cand < T.std_min || cand > T.std_max) continue;
        new_v = cand;
        gave_up = false;
        break;
    }
    new_r = old_r;
    if (T.correlated) {
        bool found = false;
        for (int it = 0; it < MAX_RESAMPLE; ++it) {
            const double cand = old_r + rng.normal(0.0, T.corr_perturb_std);
            if (cand < T.corr_min || cand > T.corr_max) continue;
            new_r = cand;
            found = true;
            break;
        }
        gave_up |= !found;
    }
    return gave_up;
}

// ---- P2-P6: one move of kind `inner` on a space whose current state (k cells) is read through `cells`; the
// result is the edit script (header of this file) with the prior/proposal ratio
template <class Cells, class R>
__device__ inline void propose_space_move(const bbb_space_spec &S, const Cells &cells, int k, int inner, R &rng, Edit &e) {
    e.rm = -1;
    e.ins = -1;
    e.lpr = 0.0;
    e.write = true;
    e.gave_up = false;
    e.site[0] = e.site[1] = 0.0;

    if (inner == BBB_MOVE_VALUES) {
        // P2 ParamPerturbation.perturb_param_space_state (perturbations/_param_values.py:60-84)
        const int idx = rng.randint(0, k - 1);
        e.rm = e.ins = idx;
        for (int d = 0; d < S.ndim; ++d) e.site[d] = cells.site(d, idx);
        for (int p = 0; p < S.n_params; ++p) {
            const PriorAt q = prior_at(S, p, e.site[0]);
            const double old_v = cells.value(p, idx);
            const double new_v = old_v + rng.normal(0.0, q.pstd);
            e.lpr += prior_value_ratio(q, old_v, new_v);
            e.vals[p] = new_v;
        }
    } else if (inner == BBB_MOVE_SITES) {
        // P3 Voronoi.perturb_value + _perturb_site (_voronoi.py:171-232; 1-D :566-591)
        const int idx = rng.randint(0, k - 1);
        e.rm = idx;
        for (int p = 0; p < S.n_params; ++p) e.vals[p] = cells.value(p, idx);
        // (constant trip counts: the two coordinates stay in registers)
        double old_site[2] = {0.0, 0.0};
#pragma unroll
        for (int d = 0; d < 2; ++d)
            if (d < S.ndim) old_site[d] = e.site[d] = cells.site(d, idx);
        e.gave_up = true;
        for (int it = 0; it < MAX_RESAMPLE; ++it) {
            double cand[2] = {0.0, 0.0};
            bool ok = true;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                if (d < S.ndim) {
                    cand[d] = old_site[d] + rng.normal(0.0, S.site_std[d]);
                    ok &= (cand[d] >= S.vmin[d]) && (cand[d] <= S.vmax[d]);
                }
            }
            if (ok) {
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (d < S.ndim) e.site[d] = cand[d];
                e.gave_up = false;
                break;
            }
        }
        for (int p = 0; p < S.n_params; ++p) {
            if (S.n_pos[p] <= 0) continue;  // log p(v; new site) - log p(v; old site), _voronoi.py:220-226
            e.lpr += prior_log_pdf(prior_at(S, p, e.site[0]), e.vals[p]) - prior_log_pdf(prior_at(S, p, old_site[0]), e.vals[p]);
        }
        if (S.ndim == 1) {
            int pos = 0;
            for (int j = 0; j < k; ++j)
                if (j != idx && cells.site(0, j) < e.site[0]) ++pos;
            e.ins = pos;
        } else {
            e.ins = idx;
        }
    } else if (inner == BBB_MOVE_BIRTH) {
        // P4/P6 Voronoi.birth (_voronoi.py:234-339, 1-D :593-611) / ParameterSpace.birth
        if (k == S.kmax) {
            e.lpr = -INFINITY;
            e.write = false;
        } else if (S.kind == BBB_SPACE_PLAIN) {
            e.ins = rng.randint(0, k);
            for (int p = 0; p < S.n_params; ++p)
                e.vals[p] = prior_sample(S.prior_kind[p], S.prior_a[p], S.prior_b[p], rng);
        } else {
            for (int d = 0; d < S.ndim; ++d) e.site[d] = rng.uniform(S.vmin[d], S.vmax[d]);
            if (S.birth_from == BBB_BIRTH_FROM_PRIOR) {
                // _initialize_params_from_prior (discretization/_discretization.py:299-321)
                for (int p = 0; p < S.n_params; ++p) e.vals[p] = prior_sample(prior_at(S, p, e.site[0]), rng);
            } else {
                // _initialize_params_from_neighbour (discretization/_discretization.py:323-375)
                const int i_near = (S.ndim == 1) ? nearest_1d(cells, k, -1, e.site[0])
                                                 : nearest_2d(cells, k, -1, e.site[0], e.site[1], rng.coop);
                for (int p = 0; p < S.n_params; ++p) {
                    const PriorAt q = prior_at(S, p, e.site[0]);
                    const double theta = q.pstd_birth;
                    const double old_v = cells.value(p, i_near);
                    const double new_v = old_v + rng.normal(0.0, theta);
                    e.vals[p] = new_v;
                    const double diff = new_v - old_v;
                    e.lpr += prior_log_pdf(q, new_v) + (dlog(theta * SQRT_TWO_PI) + diff * diff / (2.0 * (theta * theta)));
                }
            }
            if (S.ndim == 1) {
                int pos = 0;  // bisect_left
                for (int j = 0; j < k; ++j)
                    if (cells.site(0, j) < e.site[0]) ++pos;
                e.ins = pos;
            } else {
                e.ins = k;
            }
        }
    } else {
        // P5/P6 Voronoi.death (_voronoi.py:341-385, 1-D :613-625) / ParameterSpace.death
        if (k == S.kmin) {
            e.lpr = -INFINITY;
            e.write = false;
        } else {
            e.rm = rng.randint(0, k - 1);
            if (S.kind != BBB_SPACE_PLAIN && S.birth_from == BBB_BIRTH_FROM_NEIGHBOUR) {
                // _log_prob_death_parameters (discretization/_discretization.py:435-465)
                double pos[2] = {0.0, 0.0};
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (d < S.ndim) pos[d] = cells.site(d, e.rm);
                int i_near = (S.ndim == 1) ? nearest_1d(cells, k, e.rm, pos[0])
                                           : nearest_2d(cells, k, e.rm, pos[0], pos[1], rng.coop);
                i_near += (i_near >= e.rm) ? 1 : 0;  // back to indices of the current state
                for (int p = 0; p < S.n_params; ++p) {
                    const PriorAt q = prior_at(S, p, pos[0]);
                    const double theta = q.pstd_birth;
                    const double gone = cells.value(p, e.rm);
                    const double near_v = cells.value(p, i_near);
                    const double diff = gone - near_v;
                    e.lpr -= prior_log_pdf(q, gone);
                    e.lpr -= dlog(theta * SQRT_TWO_PI) + diff * diff / (2.0 * (theta * theta));
                }
            }
        }
    }
}

// Writes the proposed slot (cur ^ 1) of space s from the current slot and the edit, and its cell count; returns the
// proposed cell count.  coop: the 32 lanes of the warp (which all hold the same edit) share the work, lanes = output
// cells; else the calling thread writes everything when `apply_now`.
__device__ inline int write_proposed_slot(const bbb_buffers &B, const bbb_space_spec &S, int s, int cur, int k, long long c,
                                          const Edit &e, bool coop, bool apply_now) {
    // ---- write the proposed slot ------------------------------------------------------------------
    // The edit (rm, ins, new cell) is published for the forward kernels; the proposed slot is written here: by
    // the warp's lanes (lanes = output cells) in the cooperative proposal kernel, by the chain's own thread in the
    // fused small-problem kernel -- one loop for every move kind.
    const int prop = cur ^ 1;
    const bool writer = !coop || (threadIdx.x & 31) == 0;
    const int k_new = e.write ? k - (e.rm >= 0 ? 1 : 0) + (e.ins >= 0 ? 1 : 0) : k;
    if (writer) k_ref(B, s, prop, c) = k_new;
    if (e.write && coop) {
        // warp-cooperative proposal kernel: the 32 lanes (which all hold the same edit) write the proposed slot,
        // lanes = output cells
        const int ndim = S.ndim, n_params = S.n_params;
        const long long C = B.n_chains, KC = (long long)S.kmax * C;
        // the two slots never overlap: __restrict__ lets the loads of a cell (and of the next cells) go out together
        const double *__restrict__ s_cur = B.sites[s] + (long long)cur * ndim * KC + c;
        double *__restrict__ s_new = B.sites[s] + (long long)prop * ndim * KC + c;
        const double *__restrict__ v_cur = B.values[s] + (long long)cur * n_params * KC + c;
        double *__restrict__ v_new = B.values[s] + (long long)prop * n_params * KC + c;
        if (e.rm >= 0 || e.ins >= 0) {
            for (int jo = threadIdx.x & 31; jo < k_new; jo += 32) {
                const int jp = jo - ((e.ins >= 0 && jo > e.ins) ? 1 : 0);
                const int src = (jo == e.ins) ? 0 : jp + ((e.rm >= 0 && jp >= e.rm) ? 1 : 0);
                // constant trip counts: the values stay in registers and all loads of a cell go out together
                double ts[2], tv[BBB_MAX_PARAMS];
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (d < ndim) ts[d] = s_cur[d * KC + (long long)src * C];
#pragma unroll
                for (int p = 0; p < BBB_MAX_PARAMS; ++p)
                    if (p < n_params) tv[p] = v_cur[p * KC + (long long)src * C];
                const bool fresh = jo == e.ins;
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (d < ndim) s_new[d * KC + (long long)jo * C] = fresh ? e.site[d] : ts[d];
#pragma unroll
                for (int p = 0; p < BBB_MAX_PARAMS; ++p)
                    if (p < n_params) v_new[p * KC + (long long)jo * C] = fresh ? e.vals[p] : tv[p];
            }
        }
    } else if (e.write && apply_now) {
        const int ndim = S.ndim, n_params = S.n_params;
        const long long C = B.n_chains, KC = (long long)S.kmax * C;
        const double *__restrict__ s_cur = B.sites[s] + (long long)cur * ndim * KC + c;
        double *__restrict__ s_new = B.sites[s] + (long long)prop * ndim * KC + c;
        const double *__restrict__ v_cur = B.values[s] + (long long)cur * n_params * KC + c;
        double *__restrict__ v_new = B.values[s] + (long long)prop * n_params * KC + c;
        int src = 0;
        for (int jo = 0; jo < k_new; ++jo) {
            if (jo == e.ins) {
                for (int d = 0; d < ndim; ++d) s_new[d * KC + (long long)jo * C] = e.site[d];
                for (int p = 0; p < n_params; ++p) v_new[p * KC + (long long)jo * C] = e.vals[p];
            } else {
                if (src == e.rm) ++src;
                for (int d = 0; d < ndim; ++d) s_new[d * KC + (long long)jo * C] = s_cur[d * KC + (long long)src * C];
                for (int p = 0; p < n_params; ++p) v_new[p * KC + (long long)jo * C] = v_cur[p * KC + (long long)src * C];
                ++src;
            }
        }
    }
    return k_new;
}

// `e`: where the edit script is built.  A thread that proposes for its own chain passes a local; the warp-cooperative
// proposers of the fused chain kernel (all lanes build the same script) pass one in shared memory, which keeps the
// script's arrays out of local memory.
template <class R>
__device__ inline void propose_chain(const bbb_problem_spec &P, const bbb_buffers &B, long long c, R &rng,
                                     bool apply_now, Edit &e) {
    const int move = choose_move(P, rng);
    const bool writer = !rng.coop || (threadIdx.x & 31) == 0;  // cooperative mode: lane 0 publishes the proposal
    e.move = move;

    if (move == 4 * P.n_spaces) {
        for (int t = 0; t < P.n_targets; ++t) {
            const bbb_target_spec &T = P.targets[t];
            if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
            double new_v, new_r;
            const bool gave_up = propose_noise_target(T, rng, B.sigma[t][c], T.correlated ? B.corr[t][c] : 0.0, new_v, new_r);
            if (gave_up && writer && B.n_exceptions != nullptr) B.n_exceptions[B.n_chains + c] += 1;
            if (writer) B.sigma[t][B.n_chains + c] = new_v;
            if (T.correlated && writer) B.corr[t][B.n_chains + c] = new_r;
        }
        if (writer) {
            B.move[c] = move;
            B.edit[c] = -1;
            B.edit[B.n_chains + c] = -1;
            B.log_prob_ratio[c] = 0.0;
        }
        e.lpr = 0.0;
        return;
    }

    const int s = move >> 2, inner = move & 3;
    const bbb_space_spec &S = P.spaces[s];
    const int cur = B.sel[s][c];
    const int k = k_ref(B, s, cur, c);
    propose_space_move(S, GlobalCells{B, S, s, cur, c}, k, inner, rng, e);

    const int k_new = write_proposed_slot(B, S, s, cur, k, c, e, rng.coop, apply_now);
    if (writer) {
        for (int d = 0; d < S.ndim; ++d) B.edit_site[(long long)d * B.n_chains + c] = e.site[d];
        for (int p = 0; p < S.n_params; ++p) B.edit_vals[(long long)p * B.n_chains + c] = e.vals[p];
    }
    e.cur = cur;
    e.k_old = k;
    e.k_new = k_new;
    if (writer && e.gave_up && B.n_exceptions != nullptr) B.n_exceptions[B.n_chains + c] += 1;
    if (writer) {
        B.move[c] = 4 * s + inner;
        B.edit[c] = e.rm;
        B.edit[B.n_chains + c] = e.ins;
        B.log_prob_ratio[c] = e.lpr;
    }
}
template <class R>
__device__ inline void propose_chain(const bbb_problem_spec &P, const bbb_buffers &B, long long c, R &rng,
                                     bool apply_now) {
    Edit e;
    propose_chain(P, B, c, rng, apply_now, e);
}

}  // namespace bbb
