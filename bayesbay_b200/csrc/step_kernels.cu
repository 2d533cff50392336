// Chain-step kernels: proposal (P0-P7), accept/commit, fused many-steps-per-launch loop for
// thread-evaluable forwards, initialisation (I1), result gathering (O1), parallel-tempering swap
// round (T1) and the small stand-alone parity kernels (F3, F4, L1-L3).
#include "finish.cuh"
#include "hosted.cuh"
#include "kernels.h"
#include "propose.cuh"

namespace bbb {

constexpr int STEP_THREADS = 64;  // 2 warps per block: spreads C = 1024 chains over 16 SMs' worth of blocks

// One WARP per chain: lane 0 owns the generator (every draw is broadcast), the O(k) nearest-site searches of
// birth/death-from-neighbour run with lanes = sites, and the lanes write the proposed slot (lanes = cells).
#ifndef PROPOSE_WARPS_N
#define PROPOSE_WARPS_N 4
#endif
constexpr int PROPOSE_WARPS = PROPOSE_WARPS_N;
// Cost class of chain c's block in the chain kernel, 0 = longest.  est = a[move] + b[move] * (points per cell): block
// times measured on the C3 shape with the two-pass rasterisation (tools/timeline.py: birth 20.1 + 2.49 x, death
// 18.8 + 2.72 x, values 14.5 + 2.11 x, sites 21.0 + 0.96 x microseconds, x = G / (100 k)).  Only the ORDER of the
// blocks depends on these numbers.
__device__ __forceinline__ int block_cost_class(const EngineParams &ep, long long c) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const int t = ep.ray_t;
    const int move = B.move[c];
    if (!target_touched(P, move, B.log_prob_ratio[c], t)) return BBB_ORDER_BUCKETS - 1;  // no forward: a few microseconds
    const bbb_target_spec &T = P.targets[t];
    const int s = T.space;
    const int k_new = k_ref(B, s, (int)B.sel[s][c] ^ 1, c);
    const float x = (float)T.n_points / (100.0f * (float)(k_new > 0 ? k_new : 1));
    const int inner = move & 3;
    const float a = inner == BBB_MOVE_BIRTH ? 20.1f : (inner == BBB_MOVE_DEATH ? 18.8f : (inner == BBB_MOVE_VALUES ? 14.5f : 21.0f));
    const float b = inner == BBB_MOVE_BIRTH ? 2.49f : (inner == BBB_MOVE_DEATH ? 2.72f : (inner == BBB_MOVE_VALUES ? 2.11f : 0.96f));
    const float est = a + b * x;
    // classes of 2.5 us between 15.5 and 38 us
    int cls = (int)((38.0f - est) * 0.4f);
    cls = cls < 0 ? 0 : (cls > BBB_ORDER_BUCKETS - 2 ? BBB_ORDER_BUCKETS - 2 : cls);
    return cls;
}

__global__ void __launch_bounds__(PROPOSE_WARPS * 32) k_propose(const __grid_constant__ EngineParams ep, long long c_first, long long c_end,
                                                                int order_slot, const int32_t *guard, const int32_t *wait_done, int wait_seq,
                                                                int32_t *wait_failed, int finish_no_forward) {
    // the chain kernel behind this launch may become resident now (it waits for this grid before it reads a proposal)
    asm volatile("griddepcontrol.launch_dependents;");
    // work counter of the (persistent) chain kernel behind this launch: the 16th word of the launch's class counts
    if (order_slot >= 0 && blockIdx.x == 0 && threadIdx.x == 0)
        order_counts(ep.buf.block_order, ep.buf.n_chains, order_slot >> 2, order_slot & 3)[15] = 0;
    const long long c = c_first + (long long)blockIdx.x * PROPOSE_WARPS + (threadIdx.x >> 5);
    if (c >= c_end) return;
    // hosted step (bbb_engine_step_hosted): the uploaded states of this part differ from the resident ones, so the
    // caches are stale and nothing may be drawn or stepped until the host has re-derived them
    if (guard != nullptr && *guard != 0) return;
    // Launched as the programmatic dependent of a hosted chain kernel (chain_kernels.cu): this warp may run while that
    // grid's last blocks are still stepping, and waits for ITS chain's completion signal (released by the chain's block
    // after the step's last write).  Bounded: a signal that never comes is reported, not waited for forever.
    if (wait_done != nullptr) {
        const int32_t *flag = wait_done + c;
        int v = 0;
        for (long long spin = 0; spin < (1LL << 24); ++spin) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == wait_seq) break;
            __nanosleep(64);
        }
        if (v != wait_seq) {
            if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile int32_t *>(wait_failed) = 1;
            return;
        }
    }
    // the warp's generator: the first 32 draws of the (chain, iteration) computed by the lanes at once into shared memory
    __shared__ double s_draw[PROPOSE_WARPS][2][32];
    const bbb_buffers &B = ep.buf;
    const int w = threadIdx.x >> 5;
    FastRng rng;
    rng.begin(s_draw[w][0], s_draw[w][1], B.tape != nullptr ? B.tape + c * B.tape_stride : nullptr,
              B.tape != nullptr ? B.tape_cursor[c] : 0, B.tape_stride, ep.spec.seed, (uint64_t)(B.chain_id0 + c), (uint64_t)B.iteration[c],
              STREAM_STEP);
    __shared__ Edit s_edit[PROPOSE_WARPS];
    Edit &ed = s_edit[w];
    propose_chain(ep.spec, ep.buf, c, rng, false, ed);
    const long long cursor = rng.after();
    if ((threadIdx.x & 31) == 0) B.tape_cursor[c] = cursor;
    // Chain-local engines, resident stepping: a proposal that needs no forward (noise move, move in another space,
    // rejected by the prior) is DECIDED here, by the warp that drew it -- same expressions, same generator positions as
    // the chain kernel's own path for such chains (finish.cuh) -- and the chain's block of the chain kernel exits at
    // once.  (Not when the proposal is drawn ahead of a hosted step: nothing may be decided before the uploaded state
    // has been compared.)
    if (finish_no_forward) {
        __syncwarp();
        const int move = ed.move;
        const double lpr = ed.lpr;
        if (!target_touched(ep.spec, move, lpr, ep.ray_t) && (threadIdx.x & 31) == 0) {
            __shared__ FinishIn s_fin[PROPOSE_WARPS];
            __shared__ FinishPre s_pre[PROPOSE_WARPS];
            FinishIn &fin = s_fin[w];
            fin.move = move;
            fin.lpr = lpr;
            fin.cursor = cursor;
            finish_gather_rest(ep, c, fin);
            Rng r2;
            rng_resume(r2, ep, c, fin);
            finish_pre(ep, c, fin, r2, -1, s_pre[w]);
            rng_end(r2, ep, c);
            finish_post(ep, c, fin, s_pre[w], -1, 0.0);
        }
    }
    if (order_slot >= 0) {  // file the chain for the longest-first block order of the chain kernel
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            const long long C = ep.buf.n_chains;
            const int cls = block_cost_class(ep, c);
            const int pos = atomicAdd(order_counts(ep.buf.block_order, C, order_slot >> 2, order_slot & 3) + cls, 1);
            if (pos < c_end - c_first) ep.buf.block_order[(long long)cls * C + c_first + pos] = (int32_t)c;
        }
    }
}

__global__ void __launch_bounds__(STEP_THREADS) k_finish(const __grid_constant__ EngineParams ep) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ep.buf.n_chains) return;
    Rng rng;
    rng_begin(rng, ep, c, STREAM_STEP, true);
    finish_chain(ep, c, rng);
    rng_end(rng, ep, c);
}

// All forwards thread-evaluable (lookup1d / linear): the whole step is one thread's work and there
// is no cross-chain dependency, so n_steps iterations run inside ONE launch (the 1-D and polyfit
// configurations are launch-latency bound otherwise).
__global__ void __launch_bounds__(STEP_THREADS) k_fused_steps(const __grid_constant__ EngineParams ep, long long n_steps) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ep.buf.n_chains) return;
    for (long long it = 0; it < n_steps; ++it) {
        Rng rng;
        rng_begin(rng, ep, c, STREAM_STEP, false);
        propose_chain(ep.spec, ep.buf, c, rng, true);
        finish_chain(ep, c, rng);
        rng_end(rng, ep, c);
    }
}

// I1: Voronoi._initialize (_voronoi.py:151-169; 1-D :551-564), ParameterSpace._initialize
// (_parameter_space.py:146-160), Prior.initialize (_prior.py:396-398,:577-597), Target.initialize
// (_target.py:115-134).  Draw order: k, sites, each parameter's k values, then sigma per target.
__global__ void __launch_bounds__(STEP_THREADS) k_init_states(const EngineParams *epp) {
    const EngineParams &ep = *epp;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B.n_chains) return;
    Rng rng;
    rng_begin(rng, ep, c, STREAM_INIT, false);
    for (int s = 0; s < P.n_spaces; ++s) {
        const bbb_space_spec &S = P.spaces[s];
        const int k = S.trans_d ? rng.randint(S.kmin, S.kinit_max) : S.kmin;
        for (int j = 0; j < k; ++j)
            for (int d = 0; d < S.ndim; ++d) site_ref(B, S, s, 0, d, j, c) = rng.uniform(S.vmin[d], S.vmax[d]);
        if (S.ndim == 1) {  // np.sort
            for (int j = 1; j < k; ++j) {
                const double v = site_ref(B, S, s, 0, 0, j, c);
                int i = j - 1;
                while (i >= 0 && site_ref(B, S, s, 0, 0, i, c) > v) {
                    site_ref(B, S, s, 0, 0, i + 1, c) = site_ref(B, S, s, 0, 0, i, c);
                    --i;
                }
                site_ref(B, S, s, 0, 0, i + 1, c) = v;
            }
        }
        for (int p = 0; p < S.n_params; ++p)
            for (int j = 0; j < k; ++j)
                value_ref(B, S, s, 0, p, j, c) =
                    prior_sample(prior_at(S, p, S.ndim == 1 ? site_ref(B, S, s, 0, 0, j, c) : 0.0), rng);
        k_ref(B, s, 0, c) = k;
        k_ref(B, s, 1, c) = k;
        B.sel[s][c] = 0;
    }
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        if (T.cov_kind == BBB_COV_HIERARCHICAL) {
            const double sg = rng.uniform(T.std_min, T.std_max);
            B.sigma[t][c] = sg;
            B.sigma[t][B.n_chains + c] = sg;
            if (T.correlated) {  // Target.initialize: std, then correlation (likelihood/_target.py:124-131)
                const double r0 = rng.uniform(T.corr_min, T.corr_max);
                B.corr[t][c] = r0;
                B.corr[t][B.n_chains + c] = r0;
            }
        }
        if (T.fwd_kind == BBB_FWD_RAY2D) B.idx_sel[t][c] = 0;
    }
    rng_end(rng, ep, c);
}

// phi of the CURRENT state for every target (ray targets: sum of the partials a CURRENT-mode SpMM
// just wrote).
__global__ void __launch_bounds__(STEP_THREADS) k_refresh_finish(const EngineParams *epp) {
    const EngineParams &ep = *epp;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B.n_chains) return;
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        Phi3 phi;
        if (T.fwd_kind != BBB_FWD_RAY2D) phi = small_target_phi(P, B, t, B.sel[T.space][c], c, nullptr, nullptr);
        else if (reduces_dpred(T)) phi = load_phi(B, t, 0, c);  // written by k_corr_sums from d_pred
        else phi = Phi3{ray_partial_sum(ep, t, c), 0.0, 0.0};
        store_phi(B, t, 0, c, phi);
        store_phi(B, t, 1, c, phi);
    }
}

// Correlated noise on a ray target: the three residual sums need neighbouring rays, which different warps of the
// ray kernel own, so that kernel writes d_pred and this one reduces it (one thread per chain, coalesced rows).
// proposal != 0: only chains whose proposal touched the target, result into the proposed slot.
__global__ void k_corr_sums(const EngineParams *epp, int t, int proposal) {
    const EngineParams &ep = *epp;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const bbb_target_spec &T = P.targets[t];
    const long long C = B.n_chains;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (proposal && !target_touched(P, B.move[c], B.log_prob_ratio[c], t)) return;
    ResidualSums acc;
    if (T.cov_kind == BBB_COV_MATRIX) {
        // d_pred -> residual in place (every consumer of the scratch re-derives it with the SpMM first), then res^T M res
        double *col = B.dpred_scratch[t] + c;
        for (int r = 0; r < T.n_data; ++r) col[(long long)r * C] -= T.dobs[r];
        acc.v.s0 = matrix_quadratic_form(T.cov_vec, col, C, T.n_data);
    } else {
        for (int r = 0; r < T.n_data; ++r) acc.add(r, T.n_data, B.dpred_scratch[t][(long long)r * C + c] - T.dobs[r], 1.0);
    }
    store_phi(B, t, proposal ? 1 : 0, c, acc.v);
}

// O1: compact the current slot of one space, [..][K][C] (blockIdx.y = cell j)
__global__ void k_gather_space(const EngineParams *epp, int s, int32_t *k_out, double *sites_out,
                               double *values_out) {
    const EngineParams &ep = *epp;
    const bbb_space_spec &S = ep.spec.spaces[s];
    const bbb_buffers &B = ep.buf;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B.n_chains) return;
    const int j = blockIdx.y;
    const int slot = B.sel[s][c];
    if (j == 0) k_out[c] = k_ref(B, s, slot, c);
    for (int d = 0; d < S.ndim; ++d)
        sites_out[((long long)d * S.kmax + j) * B.n_chains + c] = site_ref(B, S, s, slot, d, j, c);
    for (int p = 0; p < S.n_params; ++p)
        values_out[((long long)p * S.kmax + j) * B.n_chains + c] = value_ref(B, S, s, slot, p, j, c);
}

// ------------------------------------------------------------------------------------------------
// Chain-state image: the CURRENT state of every chain as one packed buffer of 8-byte words, so that a host <->
// device exchange of the chains (what the reference's pool does with pickles on every advance_chain,
// samplers/_samplers.py:225-237) is ONE copy each way.  Rows of C words:
//   per space: k (int64), then sites [ndim][kb] and values [n_params][kb] (cells j >= k hold 0),
//   per hierarchical target: sigma (and correlation), then temperature, iteration (int64).
// MODE 0 packs the engine's state into the image, 1 compares (sets *differ when a chain's state is not bit-identical
// to the image), 2 writes the image into slot 0 of the engine (the caller then re-derives the caches).
// ------------------------------------------------------------------------------------------------
struct ImageRows {
    int kb[BBB_MAX_SPACES];
    int first_row[BBB_MAX_SPACES + 1];  // first row of each space; [n_spaces] = first noise row
    int n_rows;
};

template <int MODE>
__global__ void k_state_image(const __grid_constant__ EngineParams ep, const __grid_constant__ ImageRows ir,
                              unsigned long long *image, int32_t *differ, long long c_first, long long n_c) {
    // the image holds the chains [c_first, c_first + n_c): rows of n_c words
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const long long lc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lc >= n_c) return;
    const long long c = c_first + lc;
    const int row = blockIdx.y;
    unsigned long long *slot = image + (long long)row * n_c + lc;
    // where does this row live in the engine?
    unsigned long long cur = 0;
    unsigned long long *dst = nullptr, *dst2 = nullptr;  // MODE 2 targets
    bool live = true;                                    // false: padding cell (j >= k), not compared
    if (row < ir.first_row[P.n_spaces]) {
        int s = 0;
        while (row >= ir.first_row[s + 1]) ++s;
        const bbb_space_spec &S = P.spaces[s];
        const int rel = row - ir.first_row[s];
        const int sl = (MODE == 2) ? 0 : (int)B.sel[s][c];
        if (rel == 0) {
            cur = (unsigned long long)(long long)k_ref(B, s, sl, c);
            if (MODE == 2) {
                const int kv = (int)(long long)*slot;
                k_ref(B, s, 0, c) = kv;
                k_ref(B, s, 1, c) = kv;
                B.sel[s][c] = 0;
                return;
            }
        } else {
            const int q = (rel - 1) / ir.kb[s], j = (rel - 1) % ir.kb[s];
            const int kc = (MODE == 2) ? (int)(long long)image[(long long)ir.first_row[s] * n_c + lc] : k_ref(B, s, sl, c);
            live = j < kc;
            double *p = q < S.ndim ? &site_ref(B, S, s, sl, q, j, c) : &value_ref(B, S, s, sl, q - S.ndim, j, c);
            cur = live ? (unsigned long long)__double_as_longlong(*p) : 0ull;
            dst = reinterpret_cast<unsigned long long *>(p);
        }
    } else {
        int rel = row - ir.first_row[P.n_spaces];
        bool found = false;
        for (int t = 0; t < P.n_targets && !found; ++t) {
            const bbb_target_spec &T = P.targets[t];
            if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
            const int n = T.correlated ? 2 : 1;
            if (rel < n) {
                double *p = rel == 0 ? B.sigma[t] : B.corr[t];
                dst = reinterpret_cast<unsigned long long *>(p + c);
                dst2 = reinterpret_cast<unsigned long long *>(p + C + c);
                found = true;
            } else {
                rel -= n;
            }
        }
        if (!found) dst = rel == 0 ? reinterpret_cast<unsigned long long *>(B.temperature + c)
                                   : reinterpret_cast<unsigned long long *>(B.iteration + c);
        cur = *dst;
    }
    if (MODE == 0) {
        *slot = cur;
    } else if (MODE == 1) {
        if (live && *slot != cur) *differ = 1;
    } else if (live) {
        *dst = *slot;
        if (dst2 != nullptr) *dst2 = *slot;
    }
}

// ------------------------------------------------------------------------------------------------
// RAGGED chain-state image (the hosted step's wire format): only the cells IN USE travel -- what a pickle of the
// reference's chain holds (samplers/_samplers.py:225-237 moves arrays of k cells, not n_dimensions_max).  Image of
// the chains [c_first, c_first + n_c), 8-byte words:
//   header, HR rows of n_c words: k of every space (int64) | sigma (and correlation) per hierarchical target |
//           temperature | iteration (int64);
//   then per space s, chain after chain: the chain's k_s cells, w_s = ndim + n_params words each (site coordinates,
//           then parameter values); chain lc of space s starts at word
//           HR n_c + sum_{s' < s} w_s' sum_lc' k_s'[lc'] + w_s sum_{lc' < lc} k_s[lc'].
// One warp per (space, chain), lanes = cells; the offsets come from block-wide sums over the k row (of the engine when
// packing, of the image when comparing / unpacking).  blockIdx.y == n_spaces: the header rows.
// MODE as k_state_image: 0 pack, 1 compare (sets *differ), 2 unpack into slot 0.
// ------------------------------------------------------------------------------------------------
constexpr int RAGGED_THREADS = 256, RAGGED_WARPS = RAGGED_THREADS / 32;

template <int MODE>
__global__ void __launch_bounds__(RAGGED_THREADS) k_state_image_ragged(const __grid_constant__ EngineParams ep, unsigned long long *image,
                                                                       int32_t *differ, long long c_first, long long n_c) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n_noise = 0;
    for (int t = 0; t < P.n_targets; ++t)
        if (P.targets[t].cov_kind == BBB_COV_HIERARCHICAL) n_noise += P.targets[t].correlated ? 2 : 1;
    const int HR = P.n_spaces + n_noise + 2;
    if ((int)blockIdx.y == P.n_spaces) {
        // ---- header rows: one thread per (row, chain)
        for (long long i = (long long)blockIdx.x * RAGGED_THREADS + tid; i < (long long)HR * n_c; i += (long long)gridDim.x * RAGGED_THREADS) {
            const int row = (int)(i / n_c);
            const long long lc = i % n_c, c = c_first + lc;
            unsigned long long *slot = image + i;
            if (row < P.n_spaces) {
                const int s = row;
                if (MODE == 2) {
                    const int kv = (int)(long long)*slot;
                    k_ref(B, s, 0, c) = kv;
                    k_ref(B, s, 1, c) = kv;
                    B.sel[s][c] = 0;
                } else {
                    const unsigned long long cur = (unsigned long long)(long long)k_ref(B, s, (int)B.sel[s][c], c);
                    if (MODE == 0) *slot = cur;
                    else if (*slot != cur) *differ = 1;
                }
                continue;
            }
            int rel = row - P.n_spaces;
            unsigned long long *dst = nullptr, *dst2 = nullptr;
            for (int t = 0; t < P.n_targets && dst == nullptr; ++t) {
                const bbb_target_spec &T = P.targets[t];
                if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
                const int n = T.correlated ? 2 : 1;
                if (rel < n) {
                    double *p = rel == 0 ? B.sigma[t] : B.corr[t];
                    dst = reinterpret_cast<unsigned long long *>(p + c);
                    dst2 = reinterpret_cast<unsigned long long *>(p + C + c);
                } else {
                    rel -= n;
                }
            }
            if (dst == nullptr) dst = rel == 0 ? reinterpret_cast<unsigned long long *>(B.temperature + c)
                                               : reinterpret_cast<unsigned long long *>(B.iteration + c);
            if (MODE == 0) *slot = *dst;
            else if (MODE == 1) { if (*slot != *dst) *differ = 1; }
            else { *dst = *slot; if (dst2 != nullptr) *dst2 = *slot; }
        }
        return;
    }
    // ---- cells of space s: warp = chain
    const int s = blockIdx.y;
    const bbb_space_spec &S = P.spaces[s];
    __shared__ long long s_part[RAGGED_WARPS];
    __shared__ int s_k[RAGGED_WARPS];
    // number of cells a chain of space q occupies in the image
    auto k_of = [&](int q, long long lc) -> int {
        if (MODE == 0) return k_ref(B, q, (int)B.sel[q][c_first + lc], c_first + lc);
        const long long kv = (long long)image[(long long)q * n_c + lc];
        return kv < 0 ? 0 : (kv > P.spaces[q].kmax ? P.spaces[q].kmax : (int)kv);
    };
    const long long lc0 = (long long)blockIdx.x * RAGGED_WARPS;  // first chain of this block
    // words in front of this block's first chain: the earlier spaces' cells, and the earlier chains' cells of this space
    long long mine = 0;
    for (int q = 0; q < s; ++q) {
        const int w = P.spaces[q].ndim + P.spaces[q].n_params;
        for (long long i = tid; i < n_c; i += RAGGED_THREADS) mine += (long long)w * k_of(q, i);
    }
    const int w = S.ndim + S.n_params;
    for (long long i = tid; i < lc0 && i < n_c; i += RAGGED_THREADS) mine += (long long)w * k_of(s, i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    const long long lc = lc0 + warp;
    const int k_img = lc < n_c ? k_of(s, lc) : 0;
    if (lane == 0) { s_part[warp] = mine; s_k[warp] = k_img; }
    __syncthreads();
    if (lc >= n_c) return;
    long long base = (long long)HR * n_c;
    for (int i = 0; i < RAGGED_WARPS; ++i) base += s_part[i];
    for (int i = 0; i < warp; ++i) base += (long long)w * s_k[i];
    const long long c = c_first + lc;
    const int sl = MODE == 2 ? 0 : (int)B.sel[s][c];
    if (MODE == 1 && k_ref(B, s, sl, c) != k_img) {  // (the header block flags it too)
        if (lane == 0) *differ = 1;
        return;
    }
    unsigned long long *cells = image + base;
    for (int j = lane; j < k_img; j += 32) {
        for (int q = 0; q < w; ++q) {
            double *p = q < S.ndim ? &site_ref(B, S, s, sl, q, j, c) : &value_ref(B, S, s, sl, q - S.ndim, j, c);
            unsigned long long *slot = cells + (long long)j * w + q;
            if (MODE == 0) *slot = (unsigned long long)__double_as_longlong(*p);
            else if (MODE == 1) { if (*slot != (unsigned long long)__double_as_longlong(*p)) *differ = 1; }
            else *p = __longlong_as_double((long long)*slot);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Save point (`_save_state`, src/bayesbay/_markov_chain.py:115-134): the CURRENT state of every chain as ONE
// chain-major image of 8-byte words, so that a save point is one asynchronous device -> host copy and a chain's
// sample is a contiguous slice on the host.  Sections (include/bbb200.h):
//   per space: k [C] (int64) | sites [C][kb][ndim] | values [C][n_params][kb]   (cells j >= k hold 0)
//   per hierarchical target: sigma [C] | correlation [C] (correlated targets);  temperature [C];
//   d_pred [C][n_data] per target (written by k_save_dpred_local / a forward + transpose, abi.cu).
// One thread per (chain, cell): reads are coalesced over chains, the (small) writes are strided.
// ------------------------------------------------------------------------------------------------
struct SaveLayout {
    long long space_off[BBB_MAX_SPACES];  // word offset of the space's k row
    int kb[BBB_MAX_SPACES];
    long long noise_off;                  // sigma / correlation rows in target order, then temperature
};

__global__ void k_save_states(const __grid_constant__ EngineParams ep, const __grid_constant__ SaveLayout lay, unsigned long long *image) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int s = blockIdx.z, j = blockIdx.y;
    if (s < P.n_spaces) {
        const bbb_space_spec &S = P.spaces[s];
        const int kb = lay.kb[s];
        if (j >= kb) return;
        const int slot = B.sel[s][c];
        const int k = k_ref(B, s, slot, c);
        unsigned long long *base = image + lay.space_off[s];
        if (j == 0) base[c] = (unsigned long long)(long long)k;
        unsigned long long *sites = base + C, *vals = sites + C * (long long)kb * S.ndim;
        for (int d = 0; d < S.ndim; ++d)
            sites[(c * kb + j) * S.ndim + d] = j < k ? (unsigned long long)__double_as_longlong(site_ref(B, S, s, slot, d, j, c)) : 0ull;
        for (int p = 0; p < S.n_params; ++p)
            vals[(c * S.n_params + p) * kb + j] = j < k ? (unsigned long long)__double_as_longlong(value_ref(B, S, s, slot, p, j, c)) : 0ull;
    } else if (j == 0) {  // the noise rows and the temperature
        unsigned long long *row = image + lay.noise_off;
        for (int t = 0; t < P.n_targets; ++t) {
            const bbb_target_spec &T = P.targets[t];
            if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
            row[c] = (unsigned long long)__double_as_longlong(B.sigma[t][c]);
            row += C;
            if (T.correlated) {
                row[c] = (unsigned long long)__double_as_longlong(B.corr[t][c]);
                row += C;
            }
        }
        row[c] = (unsigned long long)__double_as_longlong(B.temperature[c]);
    }
}

// d_pred of a chain-local ray target from the resident residuals: out[c][r] = res[c][r] + dobs[r]
__global__ void k_save_dpred_local(const double *res, const double *dobs, long long n, int R, double *out) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const double2 v = *reinterpret_cast<const double2 *>(res + i);
        const int r = (int)(i % R);
        double2 o;
        o.x = v.x + dobs[r];
        o.y = v.y + dobs[r + 1 < R ? r + 1 : 0];
        *reinterpret_cast<double2 *>(out + i) = o;
    } else if (i < n) {
        out[i] = res[i] + dobs[i % R];
    }
}

__global__ void __launch_bounds__(STEP_THREADS) k_small_dpred(const EngineParams *epp, int t, double *dpred_out) {
    const EngineParams &ep = *epp;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ep.buf.n_chains) return;
    small_target_phi(ep.spec, ep.buf, t, ep.buf.sel[ep.spec.targets[t].space][c], c, dpred_out, nullptr);
}

__global__ void __launch_bounds__(STEP_THREADS) k_misfit_logdet(const EngineParams *epp, double *misfit_out,
                                                                double *logdet_out) {
    const EngineParams &ep = *epp;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B.n_chains) return;
    double m = 0.0, ld = 0.0;
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        const bool hier = (T.cov_kind == BBB_COV_HIERARCHICAL);
        double mt, lt;
        misfit_logdet_terms(T, load_phi(B, t, 0, c), hier ? B.sigma[t][c] : 1.0, T.correlated ? B.corr[t][c] : 0.0, mt, lt);
        m += mt;
        if (hier) ld += lt;
    }
    if (logdet_out != nullptr) {
        misfit_out[c] = m;
        logdet_out[c] = ld;
    } else {  // packed rows (misfit, logdet, T): what a rank contributes to a parallel-tempering swap round
        misfit_out[3 * c + 0] = m;
        misfit_out[3 * c + 1] = ld;
        misfit_out[3 * c + 2] = B.temperature[c];
    }
}

// ------------------------------------------------------------------------------------------------
// T1: ParallelTempering.on_end_advance_chain (samplers/_samplers.py:334-342).  n_total sequential
// random-pair attempts (temperatures live in shared memory when they fit).  Quirk Q2 (ratio already
// divided by T1) is reproduced.
// ------------------------------------------------------------------------------------------------
constexpr int SWAP_THREADS = 256;
constexpr int SWAP_SMEM_MAX = 8192;   // temperatures, reciprocals and conflict stamps in shared memory up to this many chains
constexpr int SWAP_CHUNK = 1024;      // attempts drawn by the whole block at a time

// Round structure: the whole block draws SWAP_CHUNK attempts (pair, log u, temperature-free part of the ratio) in
// parallel into shared memory; warp 0 then applies them in order, 32 at a time.  The attempts are sequential only
// through the temperatures of the chains they touch: pairwise disjoint pairs commute, so every round of a batch
// applies the longest pending prefix of mutually disjoint pairs at once (the whole batch 3 times out of 4 at 8192
// chains, a few rounds otherwise).
// 1 / T is kept next to T (same rounding as dividing afresh), so a decision costs one division.
__global__ void __launch_bounds__(SWAP_THREADS) k_pt_swap(double *gathered, long long n_total, uint64_t seed,
                                                          long long round, long long id0, long long n_local,
                                                          double *temperature_out, int32_t *n_accepted_out,
                                                          const double *tape, int use_smem) {
    extern __shared__ double swap_smem[];
    __shared__ double s_core[SWAP_CHUNK], s_logu[SWAP_CHUNK];
    __shared__ int s_a[SWAP_CHUNK], s_b[SWAP_CHUNK];
    double *temps = use_smem ? swap_smem : nullptr;
    double *itemps = swap_smem + n_total;                               // use_smem: 1 / temps
    int *stamp = reinterpret_cast<int *>(swap_smem + 2 * n_total);      // use_smem: which attempt touches a chain
    if (use_smem) {
        for (long long i = threadIdx.x; i < n_total; i += blockDim.x) {
            const double tv = gathered[3 * i + 2];
            temps[i] = tv;
            itemps[i] = 1.0 / tv;
            stamp[i] = 0x7fffffff;
        }
    }
    int n_acc = 0;  // warp 0
    for (long long chunk0 = 0; chunk0 < n_total; chunk0 += SWAP_CHUNK) {
        __syncthreads();
        for (int j = threadIdx.x; j < SWAP_CHUNK; j += blockDim.x) {
            const long long i = chunk0 + j;
            if (i >= n_total) break;
            Rng rng;
            if (tape != nullptr) {
                rng.init_tape(tape, 3 * i);
            } else {
                rng.init_philox(seed, 0, (uint64_t)round, STREAM_SWAP);
                rng.n = (uint32_t)(3 * i);
            }
            const int n = (int)n_total;
            const int a = min((int)(rng.u() * (double)n), n - 1);
            int b = min((int)(rng.u() * (double)(n - 1)), n - 2);
            if (b >= a) ++b;
            s_a[j] = a;
            s_b[j] = b;
            s_logu[j] = log(rng.unit());
            s_core[j] = (gathered[3 * (long long)a + 1] - gathered[3 * (long long)b + 1] +
                         gathered[3 * (long long)a + 0] - gathered[3 * (long long)b + 0]) / 2.0;
        }
        __syncthreads();
        if (threadIdx.x >= 32) continue;
        const int lane = threadIdx.x;
        const int n_chunk = (int)min((long long)SWAP_CHUNK, n_total - chunk0);
        for (int base = 0; base < n_chunk; base += 32) {
            const int j = base + lane;
            const bool mine = j < n_chunk;
            const int a = mine ? s_a[j] : 0, b = mine ? s_b[j] : 0;
            const double log_u = mine ? s_logu[j] : 0.0, core = mine ? s_core[j] : 0.0;
            if (use_smem) {
                // lanes still to be applied, in order: the longest prefix whose pairs are pairwise disjoint commutes
                // and is decided by its own lanes at once; a lane whose pair meets an EARLIER pending lane's waits
                unsigned todo = __ballot_sync(0xffffffffu, mine);
                while (todo) {
                    const bool active = (todo >> lane) & 1u;
                    if (active) { atomicMin(&stamp[a], lane); atomicMin(&stamp[b], lane); }
                    __syncwarp();
                    const bool early = active && (stamp[a] < lane || stamp[b] < lane);
                    const unsigned em = __ballot_sync(0xffffffffu, early);
                    const unsigned go = em ? (todo & ((1u << (__ffs(em) - 1)) - 1u)) : todo;
                    if (active) { stamp[a] = 0x7fffffff; stamp[b] = 0x7fffffff; }
                    __syncwarp();
                    bool swap = false;
                    if ((go >> lane) & 1u) {
                        const double t1 = temps[a], t2 = temps[b], i1 = itemps[a], i2 = itemps[b];
                        const double llr = core / t1;
                        const double prob = (i1 - i2) * llr;
                        swap = prob > log_u;
                        if (swap) { temps[a] = t2; temps[b] = t1; itemps[a] = i2; itemps[b] = i1; }
                    }
                    n_acc += __popc(__ballot_sync(0xffffffffu, swap));
                    todo &= ~go;
                    __syncwarp();
                }
                continue;
            }
            const int cnt = min(32, n_chunk - base);
            for (int q = 0; q < cnt; ++q) {
                const int qa = __shfl_sync(0xffffffffu, a, q);
                const int qb = __shfl_sync(0xffffffffu, b, q);
                const double qlu = __shfl_sync(0xffffffffu, log_u, q);
                const double qcore = __shfl_sync(0xffffffffu, core, q);
                const double t1 = use_smem ? temps[qa] : gathered[3 * (long long)qa + 2];
                const double t2 = use_smem ? temps[qb] : gathered[3 * (long long)qb + 2];
                const double i1 = use_smem ? itemps[qa] : 1.0 / t1, i2 = use_smem ? itemps[qb] : 1.0 / t2;
                const double llr = qcore / t1;
                const double prob = (i1 - i2) * llr;
                const bool swap = prob > qlu;
                __syncwarp();
                if (swap && lane == 0) {
                    if (use_smem) { temps[qa] = t2; temps[qb] = t1; itemps[qa] = i2; itemps[qb] = i1; }
                    else { gathered[3 * (long long)qa + 2] = t2; gathered[3 * (long long)qb + 2] = t1; }
                }
                if (!use_smem) __threadfence_block();
                __syncwarp();
                n_acc += swap ? 1 : 0;
            }
        }
    }
    if (threadIdx.x == 0 && n_accepted_out != nullptr) *n_accepted_out = n_acc;
    __syncthreads();
    if (use_smem) {
        for (long long i = threadIdx.x; i < n_total; i += blockDim.x) gathered[3 * i + 2] = temps[i];
    }
    for (long long i = threadIdx.x; i < n_local; i += blockDim.x)
        temperature_out[i] = use_smem ? temps[id0 + i] : gathered[3 * (id0 + i) + 2];
}

constexpr int SWAP_WAVE_THREADS = 1024;  // attempts in flight in the wave-parallel kernel (one per thread)

// The attempts are sequential only through the temperatures of the chains they touch: two attempts commute unless
// they share a chain.  Every thread holds one attempt of the current window of 1024; in every wave the attempts stamp
// their two chains with atomicMin(index), and an attempt whose two stamps are its own index has no EARLIER pending
// attempt on either chain -- everything before it on those chains has been applied -- so it decides and swaps now.
// All such attempts of a wave are pairwise disjoint.  A window drains in as many waves as its longest chain of
// attempts sharing chains (3-4 at 8192 chains, ~10 at 1024); the result is bit-identical to the sequential loop.
// 1 / T is kept next to T (same rounding as dividing afresh), so a decision costs one division.
__global__ void __launch_bounds__(SWAP_WAVE_THREADS) k_pt_swap_waves(double *gathered, long long n_total, uint64_t seed,
                                                                     long long round, long long id0, long long n_local,
                                                                     double *temperature_out, int32_t *n_accepted_out,
                                                                     const double *tape) {
    extern __shared__ double swap_smem[];
    double *temps = swap_smem, *itemps = swap_smem + n_total;
    int *stamp = reinterpret_cast<int *>(swap_smem + 2 * n_total);
    __shared__ int s_acc;
    const int n = (int)n_total, tid = threadIdx.x;
    for (int i = tid; i < n; i += SWAP_WAVE_THREADS) {
        const double tv = gathered[3 * (long long)i + 2];
        temps[i] = tv;
        itemps[i] = 1.0 / tv;
        stamp[i] = 0x7fffffff;
    }
    if (tid == 0) s_acc = 0;
    __syncthreads();
    int n_acc = 0;
    // Stamps carry the wave they were written in, counted DOWN, above the attempt's index: a stamp of a later wave is
    // smaller than any stamp of an earlier one, so atomicMin overrides stale stamps and nothing has to be reset between
    // waves (two barriers per wave instead of three).
    int wave_key = 0x7fffffff >> 10;
    for (long long w0 = 0; w0 < n_total; w0 += SWAP_WAVE_THREADS) {
        const long long i = w0 + tid;
        bool pending = i < n_total;
        int a = 0, b = 0;
        double log_u = 0.0, core = 0.0;
        if (pending) {
            Rng rng;
            if (tape != nullptr) {
                rng.init_tape(tape, 3 * i);
            } else {
                rng.init_philox(seed, 0, (uint64_t)round, STREAM_SWAP);
                rng.n = (uint32_t)(3 * i);
            }
            a = min((int)(rng.u() * (double)n), n - 1);
            b = min((int)(rng.u() * (double)(n - 1)), n - 2);
            if (b >= a) ++b;
            log_u = dlog(rng.unit());
            core = (gathered[3 * (long long)a + 1] - gathered[3 * (long long)b + 1] + gathered[3 * (long long)a + 0] -
                    gathered[3 * (long long)b + 0]) / 2.0;
        }
        for (;;) {
            --wave_key;
            const int key = (wave_key << 10) | tid;
            if (pending) {
                atomicMin(&stamp[a], key);
                atomicMin(&stamp[b], key);
            }
            __syncthreads();
            const bool ready = pending && stamp[a] == key && stamp[b] == key;
            if (ready) {
                const double t1 = temps[a], t2 = temps[b], i1 = itemps[a], i2 = itemps[b];
                const double llr = core / t1;
                const double prob = (i1 - i2) * llr;
                if (prob > log_u) {
                    temps[a] = t2; temps[b] = t1; itemps[a] = i2; itemps[b] = i1;
                    ++n_acc;
                }
                pending = false;
            }
            if (!__syncthreads_or(pending ? 1 : 0)) break;
        }
    }
    if (n_acc) atomicAdd(&s_acc, n_acc);
    __syncthreads();
    if (tid == 0 && n_accepted_out != nullptr) *n_accepted_out = s_acc;
    for (int i = tid; i < n; i += SWAP_WAVE_THREADS) gathered[3 * (long long)i + 2] = temps[i];
    for (long long i = tid; i < n_local; i += SWAP_WAVE_THREADS) temperature_out[i] = temps[id0 + i];
}

// ------------------------------------------------------------------------------------------------
// stand-alone parity kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_raster1d(const double *sites, const double *values, const int32_t *k, int K, long long C,
                           const double *x, int R, int32_t *idx_out, double *field_out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int kc = k[c];
    int hint = 0;
    auto at = [&](int j) { return sites[(long long)j * C + c]; };
    for (int r = 0; r < R; ++r) {
        const int i = cell_1d_at(at, kc, x[r], hint);
        if (idx_out) idx_out[(long long)r * C + c] = i;
        if (field_out) field_out[(long long)r * C + c] = values[(long long)i * C + c];
    }
}

__global__ void k_dense_forward(const double *G, const double *m, int R, int M, long long C, double *dpred) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= C) return;
    double d = 0.0;
    for (int q = 0; q < M; ++q) d += G[(long long)r * M + q] * m[(long long)q * C + c];
    dpred[(long long)r * C + c] = d;
}

__global__ void k_loglike(const double *dpred, const double *dobs, const double *sigma, int R, long long C,
                          double *misfit, double *logdet) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double phi = 0.0;
    for (int r = 0; r < R; ++r) {
        const double res = dpred[(long long)r * C + c] - dobs[r];
        phi += res * res;
    }
    const double sg = sigma[c];
    misfit[c] = (1.0 / (sg * sg)) * phi;
    logdet[c] = (2.0 * R) * log(sg);
}

__global__ void k_logprior(bbb_space_spec S, const double *values, const int32_t *k, long long C, double *out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double tot = 0.0;
    const int kc = k[c];
    for (int p = 0; p < S.n_params; ++p) {
        double sub = 0.0;
        for (int j = 0; j < kc; ++j)
            sub += prior_log_pdf(S.prior_kind[p], S.prior_a[p], S.prior_b[p],
                                 values[((long long)p * S.kmax + j) * C + c]);
        tot += sub;
    }
    out[c] = tot;
}

__global__ void k_phi_from_partial(const double *partial, int n_tiles, long long C, double *phi) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int i = 0; i < n_tiles; ++i) s += partial[(long long)i * C + c];
    phi[c] = s;
}

// ------------------------------------------------------------------------------------------------
static inline int blocks_for(long long C, int threads) { return (int)((C + threads - 1) / threads); }

int launch_propose(const EngineParams &p, long long c_first, long long c_end, int order_slot, const int32_t *guard, cudaStream_t st,
                   const int32_t *wait_done, int wait_seq, int32_t *wait_failed, bool finish_no_forward) {
    if (!p.chain_local || p.buf.block_order == nullptr) order_slot = -1;
    if (!p.chain_local) finish_no_forward = false;  // (the full-SpMM step decides every chain in k_finish)
    if (wait_done == nullptr) {
        k_propose<<<blocks_for(c_end - c_first, PROPOSE_WARPS), PROPOSE_WARPS * 32, 0, st>>>(p, c_first, c_end, order_slot, guard, nullptr, 0, nullptr,
                                                                                             finish_no_forward ? 1 : 0);
        return (int)cudaGetLastError();
    }
    // programmatic dependent of the hosted chain kernel in front of it in `st` (per-chain completion signals)
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks_for(c_end - c_first, PROPOSE_WARPS));
    cfg.blockDim = dim3(PROPOSE_WARPS * 32);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int rc = (int)cudaLaunchKernelEx(&cfg, k_propose, p, c_first, c_end, order_slot, guard, wait_done, wait_seq, wait_failed, 0);
    return rc != 0 ? rc : (int)cudaGetLastError();
}
int launch_corr_sums(const EngineParams *p, long long C, int target, int proposal, cudaStream_t st) {
    k_corr_sums<<<blocks_for(C, 128), 128, 0, st>>>(p, target, proposal);
    return (int)cudaGetLastError();
}
int launch_finish(const EngineParams &p, long long C, cudaStream_t st) {
    k_finish<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_fused_steps(const EngineParams &p, long long C, long long n_steps, cudaStream_t st) {
    k_fused_steps<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p, n_steps);
    return (int)cudaGetLastError();
}
int launch_init_states(const EngineParams *p, long long C, cudaStream_t st) {
    k_init_states<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_refresh_finish(const EngineParams *p, long long C, cudaStream_t st) {
    k_refresh_finish<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p);
    return (int)cudaGetLastError();
}
int launch_gather_space(const EngineParams *p, long long C, int space, int kmax, int32_t *k_out,
                          double *sites_out, double *values_out, cudaStream_t st) {
    dim3 grid(blocks_for(C, 128), kmax);
    k_gather_space<<<grid, 128, 0, st>>>(p, space, k_out, sites_out, values_out);
    return (int)cudaGetLastError();
}
static ImageRows image_rows(const EngineParams &p, const int32_t *kb) {
    ImageRows ir;
    int row = 0;
    for (int s = 0; s < p.spec.n_spaces; ++s) {
        const bbb_space_spec &S = p.spec.spaces[s];
        ir.kb[s] = kb[s];
        ir.first_row[s] = row;
        row += 1 + (S.ndim + S.n_params) * kb[s];
    }
    ir.first_row[p.spec.n_spaces] = row;
    for (int t = 0; t < p.spec.n_targets; ++t)
        if (p.spec.targets[t].cov_kind == BBB_COV_HIERARCHICAL) row += p.spec.targets[t].correlated ? 2 : 1;
    ir.n_rows = row + 2;  // temperature, iteration
    return ir;
}
long long state_image_rows(const EngineParams &p, const int32_t *kb) { return image_rows(p, kb).n_rows; }
long long state_image_words(const EngineParams &p, const int32_t *kb) { return (long long)image_rows(p, kb).n_rows * p.buf.n_chains; }
int launch_state_image(const EngineParams &p, int mode, const int32_t *kb, void *image, int32_t *differ, long long c_first,
                       long long c_end, cudaStream_t st) {
    const ImageRows ir = image_rows(p, kb);
    if (ir.n_rows > 65535) return (int)cudaErrorInvalidConfiguration;
    if (c_end <= c_first) return 0;
    const long long n_c = c_end - c_first;
    dim3 grid(blocks_for(n_c, 128), ir.n_rows);
    unsigned long long *img = reinterpret_cast<unsigned long long *>(image);
    if (mode == 0) k_state_image<0><<<grid, 128, 0, st>>>(p, ir, img, differ, c_first, n_c);
    else if (mode == 1) k_state_image<1><<<grid, 128, 0, st>>>(p, ir, img, differ, c_first, n_c);
    else k_state_image<2><<<grid, 128, 0, st>>>(p, ir, img, differ, c_first, n_c);
    return (int)cudaGetLastError();
}
// Records of the mapped hosted step (hosted.cuh): one warp per chain.  k_hosted_pack writes the CURRENT state of every
// chain into its record of io.out; k_hosted_unpack writes the records of io.in whose chain is flagged in io.only into
// slot 0 of the engine (the caller then re-derives the caches).
__global__ void __launch_bounds__(256) k_hosted_pack(const __grid_constant__ EngineParams ep, const __grid_constant__ HostedIO io) {
    const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= ep.buf.n_chains) return;
    hosted_write(ep, io, c, threadIdx.x & 31, 32, 0);
}
__global__ void __launch_bounds__(256) k_hosted_unpack(const __grid_constant__ EngineParams ep, const __grid_constant__ HostedIO io) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C || (io.only != nullptr && io.only[c] == 0)) return;
    const unsigned long long *rec = io.in + c * io.stride;
    for (int s = 0; s < P.n_spaces; ++s) {
        const bbb_space_spec &S = P.spaces[s];
        long long kv = (long long)rec[s];
        kv = kv < S.kmin ? S.kmin : (kv > S.kmax ? S.kmax : kv);  // (a record the host filled with nonsense cannot index outside the arrays)
        const int w = S.ndim + S.n_params;
        for (int i = lane; i < (int)kv * w; i += 32) {
            const int j = i / w, q = i - j * w;
            const double v = __longlong_as_double((long long)rec[io.base[s] + i]);
            if (q < S.ndim) site_ref(B, S, s, 0, q, j, c) = v;
            else value_ref(B, S, s, 0, q - S.ndim, j, c) = v;
        }
        if (lane == 0) {
            k_ref(B, s, 0, c) = (int)kv;
            k_ref(B, s, 1, c) = (int)kv;
            B.sel[s][c] = 0;
        }
    }
    if (lane == 0) {
        int i = P.n_spaces;
        for (int t = 0; t < P.n_targets; ++t) {
            const bbb_target_spec &T = P.targets[t];
            if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
            const double sg = __longlong_as_double((long long)rec[i++]);
            B.sigma[t][c] = sg;
            B.sigma[t][C + c] = sg;
            if (T.correlated) {
                const double rho = __longlong_as_double((long long)rec[i++]);
                B.corr[t][c] = rho;
                B.corr[t][C + c] = rho;
            }
        }
        B.temperature[c] = __longlong_as_double((long long)rec[i++]);
        B.iteration[c] = (int64_t)rec[i];
    }
}
int launch_hosted_pack(const EngineParams &p, const HostedIO &io, cudaStream_t st) {
    k_hosted_pack<<<blocks_for(p.buf.n_chains, 8), 256, 0, st>>>(p, io);
    return (int)cudaGetLastError();
}
int launch_hosted_unpack(const EngineParams &p, const HostedIO &io, cudaStream_t st) {
    k_hosted_unpack<<<blocks_for(p.buf.n_chains, 8), 256, 0, st>>>(p, io);
    return (int)cudaGetLastError();
}
// largest current k of every space over the resident chains (out[s], zeroed by the caller)
__global__ void k_max_k(const EngineParams *epp, int32_t *out) {
    const EngineParams &ep = *epp;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < ep.spec.n_spaces; ++s) {
        int k = c < ep.buf.n_chains ? k_ref(ep.buf, s, (int)ep.buf.sel[s][c], c) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
        if ((threadIdx.x & 31) == 0 && k > 0) atomicMax(out + s, k);
    }
}
int launch_max_k(const EngineParams *p, long long C, int32_t *out, cudaStream_t st) {
    k_max_k<<<blocks_for(C, 256), 256, 0, st>>>(p, out);
    return (int)cudaGetLastError();
}
int ragged_header_rows(const EngineParams &p) {
    int rows = p.spec.n_spaces + 2;  // k per space, temperature, iteration
    for (int t = 0; t < p.spec.n_targets; ++t)
        if (p.spec.targets[t].cov_kind == BBB_COV_HIERARCHICAL) rows += p.spec.targets[t].correlated ? 2 : 1;
    return rows;
}
int launch_state_image_ragged(const EngineParams &p, int mode, void *image, int32_t *differ, long long c_first, long long c_end,
                              cudaStream_t st) {
    if (c_end <= c_first) return 0;
    const long long n_c = c_end - c_first;
    dim3 grid((unsigned)((n_c + RAGGED_WARPS - 1) / RAGGED_WARPS), p.spec.n_spaces + 1);
    unsigned long long *img = reinterpret_cast<unsigned long long *>(image);
    if (mode == 0) k_state_image_ragged<0><<<grid, RAGGED_THREADS, 0, st>>>(p, img, differ, c_first, n_c);
    else if (mode == 1) k_state_image_ragged<1><<<grid, RAGGED_THREADS, 0, st>>>(p, img, differ, c_first, n_c);
    else k_state_image_ragged<2><<<grid, RAGGED_THREADS, 0, st>>>(p, img, differ, c_first, n_c);
    return (int)cudaGetLastError();
}
static SaveLayout save_layout(const EngineParams &p, const int32_t *kb, long long *n_words_states) {
    SaveLayout lay;
    const long long C = p.buf.n_chains;
    long long off = 0;
    for (int s = 0; s < p.spec.n_spaces; ++s) {
        const bbb_space_spec &S = p.spec.spaces[s];
        lay.space_off[s] = off;
        lay.kb[s] = kb[s];
        off += C * (1 + (long long)kb[s] * (S.ndim + S.n_params));
    }
    lay.noise_off = off;
    for (int t = 0; t < p.spec.n_targets; ++t)
        if (p.spec.targets[t].cov_kind == BBB_COV_HIERARCHICAL) off += C * (p.spec.targets[t].correlated ? 2 : 1);
    off += C;  // temperature
    *n_words_states = off;
    return lay;
}
long long save_point_state_words(const EngineParams &p, const int32_t *kb) {
    long long n = 0;
    save_layout(p, kb, &n);
    return n;
}
int launch_save_states(const EngineParams &p, const int32_t *kb, void *image, cudaStream_t st) {
    long long n = 0;
    const SaveLayout lay = save_layout(p, kb, &n);
    int kb_max = 1;
    for (int s = 0; s < p.spec.n_spaces; ++s) kb_max = kb[s] > kb_max ? kb[s] : kb_max;
    dim3 grid(blocks_for(p.buf.n_chains, 128), kb_max, p.spec.n_spaces + 1);
    k_save_states<<<grid, 128, 0, st>>>(p, lay, reinterpret_cast<unsigned long long *>(image));
    return (int)cudaGetLastError();
}
int launch_save_dpred_local(const double *res, const double *dobs, long long C, int R, double *out, cudaStream_t st) {
    const long long n = C * R;
    k_save_dpred_local<<<blocks_for((n + 1) / 2, 256), 256, 0, st>>>(res, dobs, n, R, out);
    return (int)cudaGetLastError();
}
int launch_small_dpred(const EngineParams *p, long long C, int target, double *dpred_out, cudaStream_t st) {
    k_small_dpred<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p, target, dpred_out);
    return (int)cudaGetLastError();
}
int launch_misfit_logdet(const EngineParams *p, long long C, double *misfit_out, double *logdet_out,
                         cudaStream_t st) {
    k_misfit_logdet<<<blocks_for(C, STEP_THREADS), STEP_THREADS, 0, st>>>(p, misfit_out, logdet_out);
    return (int)cudaGetLastError();
}
int launch_pt_swap(const double *gathered, long long n_total, uint64_t seed, long long round, long long id0,
                   long long n_local, double *temperature_out, int32_t *n_accepted_out, const double *tape,
                   cudaStream_t st) {
    if (n_total <= SWAP_SMEM_MAX) {  // temperatures fit in shared memory: wave-parallel replay
        const size_t smem_w = (size_t)n_total * (2 * sizeof(double) + sizeof(int));
        if (smem_w > 32 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(k_pt_swap_waves, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
            if (err != cudaSuccess) return (int)err;
        }
        k_pt_swap_waves<<<1, SWAP_WAVE_THREADS, smem_w, st>>>(const_cast<double *>(gathered), n_total, seed, round, id0, n_local,
                                                              temperature_out, n_accepted_out, tape);
        return (int)cudaGetLastError();
    }
    const int use_smem = 0;
    const size_t smem = 0;
    if (smem > 16 * 1024) {  // the kernel also has 24 KB of static shared memory
        cudaError_t err = cudaFuncSetAttribute(k_pt_swap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
    }
    k_pt_swap<<<1, SWAP_THREADS, smem, st>>>(const_cast<double *>(gathered), n_total, seed, round, id0, n_local,
                                             temperature_out, n_accepted_out, tape, use_smem);
    return (int)cudaGetLastError();
}
int launch_raster1d(const double *sites, const double *values, const int32_t *k, int K, long long C,
                    const double *x, int R, int32_t *idx_out, double *field_out, cudaStream_t st) {
    k_raster1d<<<blocks_for(C, 64), 64, 0, st>>>(sites, values, k, K, C, x, R, idx_out, field_out);
    return (int)cudaGetLastError();
}
int launch_dense_forward(const double *G, const double *m, int R, int M, long long C, double *dpred_out,
                         cudaStream_t st) {
    dim3 grid(blocks_for(C, 128), R);
    k_dense_forward<<<grid, 128, 0, st>>>(G, m, R, M, C, dpred_out);
    return (int)cudaGetLastError();
}
int launch_loglike(const double *dpred, const double *dobs, const double *sigma, int R, long long C,
                   double *misfit_out, double *logdet_out, cudaStream_t st) {
    k_loglike<<<blocks_for(C, 128), 128, 0, st>>>(dpred, dobs, sigma, R, C, misfit_out, logdet_out);
    return (int)cudaGetLastError();
}
int launch_logprior(const bbb_space_spec &space, const double *values, const int32_t *k, long long C, double *out,
                    cudaStream_t st) {
    k_logprior<<<blocks_for(C, 128), 128, 0, st>>>(space, values, k, C, out);
    return (int)cudaGetLastError();
}
int launch_phi_from_partial(const double *partial, int n_tiles, long long C, double *phi_out, cudaStream_t st) {
    k_phi_from_partial<<<blocks_for(C, 128), 128, 0, st>>>(partial, n_tiles, C, phi_out);
    return (int)cudaGetLastError();
}

}  // namespace bbb
