// Chain-local step of a 2-D ray-tomography problem: ONE thread block per chain does the change-local
// rasterisation (F1), the change-local forward (F2), the misfit (L1), the Metropolis-Hastings decision (L2, P0)
// and the commit, with the chain's per-ray accumulators in shared memory.
//
// A proposal edits one Voronoi cell, so the slowness field changes only at the grid points D whose nearest site
// changes (birth: points the new cell captures; death: the removed cell's points; site move: both; value move:
// the cell's own points).  `jacobian @ (1 / vel[nearest])` (31_sw_tomography.ipynb cell 12) is linear in the
// field, hence
//     d_pred' = d_pred + sum_{g in D} G[:, g] * (f(v'(g)) - f(v(g)))
// which needs the COLUMNS of the ray matrix: |D| * nnz/G entries (C3: ~220 * 83 = 18 k) instead of all nnz
// (828 k).  The columns come from a ray-tile-blocked CSC copy of the matrix shared by all chains (L2 resident).
//
// Reproducibility: the per-ray changes are accumulated in shared memory by many warps at once.  Floating-point
// atomics would make the result depend on the arrival order, so the terms are accumulated in 64-bit FIXED POINT
// (integer addition is associative): each term w * delta is rounded once to a multiple of 2^-s, with s chosen
// per chain-step from a bound of the sum so that the accumulator cannot overflow 2^62.  The quantum is at least
// 2^9 times finer than the fp64 rounding of a term of maximal size, and the result is bit-identical run to run
// and independent of how chains are sharded.
//
// Merged columns: the changed points come in spatially compact runs, and neighbouring columns hold mostly the
// same rays.  Besides the columns themselves (level 0) the CSC copy therefore holds the sums of every aligned group
// of 4 (level 1) and 16 (level 2) consecutive columns; when all points of a group change between the same two
// cells, ONE merged column replaces 4 or 16 (the host orders the grid points along a Morton curve, so a group is a
// 2x2 / 4x4 block of the grid).  Fewer entries = fewer shared-memory atomics, the kernel's bound.
//
// Resident per-chain caches (chain-major, unlike the [..][C] arrays of the other kernels, because here one block
// walks one chain): idx_cm [C][G'] nearest-site index, res [C][R] residual d_pred - d_obs, phi = sum w res^2.
// Rounding cannot accumulate without bound: the engine re-evaluates res and phi from scratch with the full SpMM
// every `resync_every` steps (abi.cu).
#include <cstdlib>

#include "finish.cuh"
#include "hosted.cuh"
#include "kernels.h"
#include "propose.cuh"

namespace bbb {

#ifndef CS_MISFIT_U
#define CS_MISFIT_U 4
#endif
#ifndef CS_AB_LIST
#define CS_AB_LIST 1
#endif

constexpr int CS_THREADS = 512;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_CHUNK = CS_THREADS;       // change-list entries prepared per round
constexpr unsigned CS_NONE = 0xFFFFFFFFu;  // list entry: point re-derived, nearest site unchanged
constexpr int CS_M_BITS = 12;              // list entry = back << 31 | level << 29 | group index << 12 | cell
constexpr int CS_G_BITS = 17;              // bits of the group index: n_points <= 2^17
constexpr unsigned CS_BACK = 1u << 31;     // entry of a point that LOST its owner: cell = its new owner (else: old owner)
constexpr unsigned CS_G_MASK = (1u << CS_G_BITS) - 1u, CS_M_MASK = (1u << CS_M_BITS) - 1u;
constexpr int CS_PTS = 16;                 // grid points per thread and pass = the largest merged group
constexpr int CS_SMEM_BLOCK = 112 * 1024;  // two resident blocks per SM
constexpr int CS_LIST_SMEM = 1024;         // change-list entries kept in shared memory at most (the rest spill to the global list;
                                           // the launch may choose fewer, see fit_list_cap)
// The misfit phases (touched-ray scan, residual arithmetic, block sum) are mapped onto the first CS_MAIN_WARPS warps in
// EVERY variant of the kernel: the order of the floating-point misfit sum then does not depend on whether the block's
// last warp is a worker (k_chain_step) or the proposer of the fused kernel (k_chain_run), and the variants stay
// bit-identical to each other.
#ifndef CS_N_PROPOSERS
#define CS_N_PROPOSERS 1
#endif
constexpr int CS_PROPOSERS = CS_N_PROPOSERS;  // fused kernel: warps of a block that draw proposals
constexpr int CS_MAIN_WARPS = CS_WARPS - CS_PROPOSERS;
constexpr int CS_MAIN = CS_MAIN_WARPS * 32;

// One drawn proposal on its way from the proposer warp to the workers of the fused kernel (shared memory): the scalars
// k_chain_step reads from the proposal kernel's global outputs, and everything of the decision that does not need the
// new misfit.
struct ItemRec {
    long long c;  // chain (-1: no more work)
    int move, cur, k_old, k_new, rm, ins;
    double lpr, nx, ny;
    FinishIn fin;
    FinishPre pre;
};

// block-wide barrier of the step body: all 512 threads (one block per chain), or the 480 workers of the fused kernel
// (named barrier 1; the proposer warp never joins it)
template <bool FUSED>
__device__ __forceinline__ void cs_sync() {
    if constexpr (FUSED) asm volatile("bar.sync 1, %0;" ::"n"(CS_MAIN) : "memory");
    else __syncthreads();
}

struct PrepEntry {
    int j0, len;  // entries of the (merged) column inside the current ray tile; the level rides in len's top 2 bits
    double dlt;   // change of the cell function at the column's points
};

// 16 consecutive nearest-site indices of one chain (one or two 16-byte loads)
template <typename IdxT>
struct IdxVec {
    uint4 w[sizeof(IdxT)];
    __device__ __forceinline__ void load(const IdxT *p) {
#pragma unroll
        for (int i = 0; i < (int)sizeof(IdxT); ++i) w[i] = reinterpret_cast<const uint4 *>(p)[i];
    }
    __device__ __forceinline__ int at(int q) const { return (int)reinterpret_cast<const IdxT *>(w)[q]; }
    // all 16 indices equal
    __device__ __forceinline__ bool uniform() const {
        const unsigned x = w[0].x;
        bool u = x == __funnelshift_l(x, x, 8 * (int)sizeof(IdxT)) && w[0].y == x && w[0].z == x && w[0].w == x;
#pragma unroll
        for (int i = 1; i < (int)sizeof(IdxT); ++i) u = u && w[i].x == x && w[i].y == x && w[i].z == x && w[i].w == x;
        return u;
    }
};

// multi-tile problems also keep a bitmap of the touched rays (one bit per ray of the target)
__host__ __device__ inline size_t chain_step_smem(int tile_rays, int K, long long n_data = 0, int list_cap = CS_LIST_SMEM) {
    const bool multi = n_data > tile_rays;
    return (size_t)tile_rays * 8 + (size_t)K * 4 * 8 + (size_t)CS_CHUNK * sizeof(PrepEntry) + (size_t)list_cap * 4 +
           (multi ? (size_t)((n_data + 31) / 32) * 4 : 0);
}

// An SM divides 256 KB between its L1 cache and shared memory in steps (shared-memory carve-outs of 100, 132, 164, 196,
// 228 KB, ...), and every resident block costs 1 KB on top of what it asks for.  Two blocks that need a few bytes more
// than a step are given the next one and lose 32 KB of L1: BASELINE configs[2] (10 000 rays, K = 200) sat 144 bytes above
// the 196 KB step, and trimming the shared-memory head of the change list by 64 entries took the chain kernel from 0.0928 to
// 0.0880 ms (back to back 0.085 -> 0.080 ms per iteration).  Single-tile problems therefore shrink that head (it only
// decides how many entries spill to the global list) when that brings two blocks under a step.
static int fit_list_cap(size_t smem_without_list, size_t static_smem) {
    if (const char *cap = getenv("BBB_CHAIN_LIST_CAP")) {  // tests: force the spill to the global list
        const long v = atol(cap);
        return (int)(v < 0 ? 0 : (v > CS_LIST_SMEM ? CS_LIST_SMEM : v));
    }
    const size_t per_block_extra = static_smem + 1024;
    const size_t steps_kb[] = {100, 132, 164, 196};
    const size_t full = smem_without_list + (size_t)CS_LIST_SMEM * 4 + per_block_extra;
    for (size_t kb : steps_kb) {
        const size_t budget = kb * 1024 / 2;  // two blocks per SM
        if (full <= budget) return CS_LIST_SMEM;  // fits this step as it is
        const size_t floor_need = smem_without_list + 256 * 4 + per_block_extra;
        if (floor_need <= budget) {
            const size_t cap = (budget - smem_without_list - per_block_extra) / 4;
            return (int)(cap & ~(size_t)63);
        }
    }
    return CS_LIST_SMEM;
}

long long chain_local_tile_rays(long long n_data, int kmax, long long n_points) {
    if (kmax > (1 << CS_M_BITS) || n_points > (1LL << CS_G_BITS) || n_data < 1) return 0;
    const long long fixed = (long long)chain_step_smem(0, kmax) + 1024;
    long long max_tile = (CS_SMEM_BLOCK - fixed) / 8;
    if (max_tile > 65535) max_tile = 65535;  // csc_rows is uint16
    if (max_tile < 1024) return 0;
    if (const char *cap = getenv("BBB_CHAIN_TILE_RAYS")) {  // tests: force the multi-tile path on small problems
        const long long v = atoll(cap);
        if (v >= 1 && v < max_tile) max_tile = v;
    }
    if (n_data <= max_tile) return n_data;  // one tile
    // several tiles: room for the touched-ray bitmap, tiles of whole 32-ray words
    max_tile = (CS_SMEM_BLOCK - fixed - 4 * ((n_data + 31) / 32)) / 8;
    if (const char *cap = getenv("BBB_CHAIN_TILE_RAYS")) {
        const long long v = atoll(cap);
        if (v >= 1 && v < max_tile) max_tile = v;
    }
    max_tile &= ~31LL;
    if (max_tile < 32) max_tile = 32;
    if (max_tile > 65504) max_tile = 65504;
    const long long n_tiles = (n_data + max_tile - 1) / max_tile;
    return (((n_data + n_tiles - 1) / n_tiles) + 31) & ~31LL;
}

template <bool FUSED>
__device__ __forceinline__ double block_sum_to_thread0(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    cs_sync<FUSED>();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < CS_MAIN_WARPS; ++w) tot += s_red[w];  // (the misfit terms live on the first CS_MAIN_WARPS warps)
    return tot;
}

// 64-bit fixed-point accumulators kept as two 32-bit planes so that both halves use the native shared-memory
// integer atomic (a 64-bit shared atomicAdd compiles to a compare-and-swap loop).  The low word's returned value
// tells whether this addition wrapped; the wraps are carried into the high word.  Every addition is exact modulo
// 2^64, so the final (hi, lo) does not depend on the order of the additions.
__device__ __forceinline__ void fx_add(unsigned *lo, unsigned *hi, int r, long long v) {
#ifdef CS_CAS64
    // experiment: planes interleaved as one 64-bit word per ray
    atomicAdd(reinterpret_cast<unsigned long long *>(lo) + r, (unsigned long long)v);
    return;
#endif
    const unsigned xl = (unsigned)v, xh = (unsigned)((unsigned long long)v >> 32);
    unsigned carry = 0;
    if (xl != 0) {
        const unsigned old = atomicAdd(lo + r, xl);
        carry = ((unsigned)(old + xl) < xl) ? 1u : 0u;
    }
    const unsigned ah = xh + carry;
    if (ah != 0) atomicAdd(hi + r, ah);
}

// Developer instrumentation (-DCS_TIMELINE): SM clock at the phase boundaries of every block, read back by
// bbb_debug_timeline (tools/timeline.py).  Not compiled into the product library.
#ifdef CS_TIMELINE
__device__ long long g_timeline[4096][24];
#define CS_TL(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) g_timeline[blockIdx.x][i] = clock64(); } while (0)
__device__ __forceinline__ long long cs_globaltimer() { long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
__device__ __forceinline__ int cs_smid() { int v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#define CS_TL_G(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) g_timeline[blockIdx.x][i] = cs_globaltimer(); } while (0)
#define CS_TL_T(i, thr) do { if (threadIdx.x == (thr) && blockIdx.x < 4096) g_timeline[blockIdx.x][i] = clock64(); } while (0)
#else
#define CS_TL(i) do {} while (0)
#define CS_TL_G(i) do {} while (0)
#define CS_TL_T(i, thr) do {} while (0)
#endif

// Native shared-memory atomics on 32-bit shared-window addresses (computed once per block): the generic-pointer
// atomicAdd re-derives the window base in front of every atomic.
// (predicated inside the asm statement: an `if` around it becomes a branch per atomic)
__device__ __forceinline__ unsigned atoms_add_if(unsigned on, unsigned saddr, unsigned v) {
    unsigned old = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
                 : "+r"(old) : "r"(saddr), "r"(v), "r"(on) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add_if(unsigned on, unsigned saddr, unsigned v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 ::"r"(saddr), "r"(v), "r"(on) : "memory");
}
// Four entries of one column: the four low-word additions are issued back to back, the carries they return go into
// the four high-word additions (adding zero leaves an untouched ray's planes zero, so nothing is tested per entry).
__device__ __forceinline__ void fx_add4(unsigned lo_s, unsigned hi_s, const int (&r)[4], const double (&w)[4], double dlt,
                                        double scale, int base_j, int len, int lane) {
    unsigned xl[4], xh[4], old[4];
    unsigned on[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        on[u] = base_j + 32 * u + lane < len ? 1u : 0u;
        const long long v = __double2ll_rn(__dmul_rn(w[u], dlt) * scale);
        xl[u] = (unsigned)v;
        xh[u] = (unsigned)((unsigned long long)v >> 32);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        old[u] = atoms_add_if(on[u], lo_s + 4u * (unsigned)r[u], xl[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned carry = ((unsigned)(old[u] + xl[u]) < xl[u]) ? 1u : 0u;
        reds_add_if(on[u], hi_s + 4u * (unsigned)r[u], xh[u] + carry);
    }
}

// developer switch: persistent launch of the chain kernel (measured slower, see k_chain_step)
#ifndef CS_PERSISTENT
#define CS_PERSISTENT 0
#endif

// One chain's step by one block (everything below the scalars of the proposal); see k_chain_step.
// MULTI: the target's rays need several accumulator tiles (csc_n_tiles > 1); problems with one tile get an
// instantiation without any of the multi-tile code.
// FUSED: the body runs on the 480 workers of k_chain_run with the proposal handed over in shared memory (`rec`).
// Returns what the accumulator planes hold afterwards: n >= 0 -- only the n rays listed in the touched-ray list (the
// `prep` area) are non-zero; -1 -- anything; -2 -- untouched (still zero).
// HOSTED: the mapped hosted step (hosted.cuh) -- the chain's uploaded record is read from pinned host memory and
// compared with the resident state before anything is committed (a chain that differs is flagged and left untouched),
// and the new state is written to the chain's record of the out image.
template <typename IdxT, bool MULTI, bool FUSED, bool HOSTED = false>
__device__ __forceinline__ int chain_step_body(const EngineParams &ep, int t, long long c, const int32_t *guard,
                                               unsigned char *cs_smem, ItemRec *rec, const HostedIO &io, int proposer_finished = 0,
                                               int list_cap = CS_LIST_SMEM) {
    constexpr int NW = FUSED ? CS_MAIN_WARPS : CS_WARPS;  // warps in the body
    constexpr int NTH = NW * 32;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const bbb_target_spec &T = P.targets[t];
    const int s = T.space;
    const bbb_space_spec &S = P.spaces[s];
    const long long C = B.n_chains;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    CS_TL(0);
    CS_TL_G(16);
#ifdef CS_TIMELINE
    if (tid == 0 && blockIdx.x < 4096) g_timeline[blockIdx.x][18] = cs_smid();
#endif

    __shared__ int s_front, s_back, s_accept;
    __shared__ int s_wtot[CS_WARPS];
    __shared__ double s_red[CS_WARPS];
    __shared__ FinishIn s_fin;
    __shared__ FinishPre s_pre;
    constexpr int RW = CS_MAIN_WARPS;  // warps of the rasterisation pass; the last warp prepares the decision meanwhile
    FinishIn &fin = FUSED ? rec->fin : s_fin;
    FinishPre &pre = FUSED ? rec->pre : s_pre;

    int move, cur, k_old, k_new, rm, ins;
    double lpr, nx, ny;
    if constexpr (FUSED) {
        move = rec->move; cur = rec->cur; k_old = rec->k_old; k_new = rec->k_new; rm = rec->rm; ins = rec->ins;
        lpr = rec->lpr; nx = rec->nx; ny = rec->ny;
    } else {
        // every scalar of the chain's proposal is requested before the first one is used: ONE memory round trip
        // instead of a chain of three (move -> selector -> cell counts)
        move = B.move[c];
        lpr = B.log_prob_ratio[c];
        cur = B.sel[s][c];
        const int k_slot0 = k_ref(B, s, 0, c), k_slot1 = k_ref(B, s, 1, c);
        rm = B.edit[c]; ins = B.edit[C + c];
        nx = B.edit_site[c]; ny = B.edit_site[C + c];
        k_old = cur ? k_slot1 : k_slot0; k_new = cur ? k_slot0 : k_slot1;
        // hosted step: stale caches (see k_propose) -- the whole launch is a no-op
        if (guard != nullptr && *guard != 0) return -2;
        if constexpr (HOSTED) {
            if (io.only != nullptr && io.only[c] == 0) return -2;  // redo pass: only the flagged chains
        }
        if (!target_touched(P, move, lpr, t)) {  // noise move, other space, or rejected by the prior: no forward
            if constexpr (HOSTED) {
                const int words_in = hosted_words_in_use(ep, io, c);
                if (__syncthreads_or(hosted_compare(ep, io, io.in + c * io.stride, c, tid, CS_THREADS) ? 1 : 0)) {
                    if (tid == 0) { io.mismatch[c] = 1; *io.host_flag = 1; hosted_signal_done(io, c); }
                    return -2;
                }
                if (tid == 0) {
                    Rng rng;
                    rng_begin(rng, ep, c, STREAM_STEP, true);
                    finish_chain(ep, c, rng);
                    rng_end(rng, ep, c);
                    hosted_signal_done(io, c);
                }
                __syncthreads();
                hosted_write(ep, io, c, tid, CS_THREADS, words_in);
                return -2;
            }
            if (tid == 0 && !proposer_finished) {  // (else: decided by the warp that drew the proposal, k_propose)
                Rng rng;
                rng_begin(rng, ep, c, STREAM_STEP, true);
                finish_chain(ep, c, rng);
                rng_end(rng, ep, c);
                CS_TL_G(17);
            }
            return -2;
        }
        if (tid == CS_THREADS - 1) finish_gather(ep, c, s_fin);  // thread 0 decides from these after the forward
    }
    int h_words = 0;
    unsigned long long h_pref = 0;
    if constexpr (HOSTED) {
        // The uploaded record is requested now (one word per thread, held in a register) and compared after the first
        // rasterisation pass, so that the trip over the bus overlaps the tables and the pass: 0.125 ms per call.  (Measured
        // alternatives, all slower on B200: an L2 prefetch of the record's lines instead of the register, 0.131 ms;
        // comparing after the second pass, 0.132; comparing right here, 0.137; asynchronous copies of the record into a
        // shared-memory staging area with the resident cells captured next to it in the tables phase, 0.142.)
        h_words = hosted_words_in_use(ep, io, c);
        // (space 0's cells follow the header: one contiguous run of the record, thread tid requests word tid of it now)
        const int run = io.hrw + k_ref(B, 0, (int)B.sel[0][c], c) * (P.spaces[0].ndim + P.spaces[0].n_params);
        if (tid < run) h_pref = io.in[c * io.stride + tid];
    }
    const int prop = cur ^ 1;
    (void)move;

    const int K = S.kmax, RT = T.csc_tile_rays, NT = MULTI ? T.csc_n_tiles : 1, G = T.n_points, R = T.n_data;
    unsigned *acc_lo = reinterpret_cast<unsigned *>(cs_smem);
    unsigned *acc_hi = acc_lo + RT;
    double *sx = reinterpret_cast<double *>(cs_smem + (size_t)RT * 8);
    double *sy = sx + K, *fo = sy + K, *fn = fo + K;
    PrepEntry *prep = reinterpret_cast<PrepEntry *>(fn + K);
    // the head of the change list and the whole re-derivation queue (it aliases `prep`, which is idle until the
    // forward) live in shared memory: a global round trip less between the phases that write and read them
    unsigned *s_list = reinterpret_cast<unsigned *>(prep + CS_CHUNK);
    unsigned *s_queue = reinterpret_cast<unsigned *>(prep);
    constexpr int CS_QUEUE_SMEM = CS_CHUNK * (int)sizeof(PrepEntry) / 4;
    unsigned *s_touch = s_list + list_cap;  // multi-tile: bitmap of the touched rays

    const int inner = move & 3;
    const bool recip = T.transform == BBB_TRANSFORM_RECIPROCAL;
    IdxT *row = reinterpret_cast<IdxT *>(B.idx_cm[t]) + c * ((long long)(G + 15) & ~15LL);
    // change list [0, G) and, behind it, the queue of 16-point groups with points to re-derive
    unsigned *list = B.change_list[t] + c * (long long)(G + (G + 15) / 16);
    unsigned *queue = list + G;
    auto list_at = [&](int pos) -> unsigned & { return pos < (CS_AB_LIST ? list_cap : 0) ? s_list[pos] : list[pos]; };
    auto queue_at = [&](int pos) -> unsigned & { return pos < (CS_AB_LIST ? CS_QUEUE_SMEM : 0) ? s_queue[pos] : queue[pos]; };
    const int n_levels = T.csc_n_levels;
    IdxVec<IdxT> wv_next;
    if (warp < RW) wv_next.load(row + min(tid * CS_PTS, G - 1) / CS_PTS * CS_PTS);  // in flight while the tables load

    // ---- tables: proposed sites, cell function of the current (fo) and proposed (fn) state --------------------
    double fmax = 0.0;
    for (int j = tid; j < k_new; j += NTH) {
        sx[j] = site_ref(B, S, s, prop, 0, j, c);
        sy[j] = site_ref(B, S, s, prop, 1, j, c);
        double v = value_ref(B, S, s, prop, T.param, j, c);
        if (recip) v = 1.0 / v;
        fn[j] = v;
        fmax = fmax > fabs(v) ? fmax : fabs(v);
        if (!(fabs(v) <= 1.79769313486231570815e308)) fmax = INFINITY;  // NaN / inf
    }
    for (int j = tid; j < k_old; j += NTH) {
        double v = value_ref(B, S, s, cur, T.param, j, c);
        if (recip) v = 1.0 / v;
        fo[j] = v;
        fmax = fmax > fabs(v) ? fmax : fabs(v);
        if (!(fabs(v) <= 1.79769313486231570815e308)) fmax = INFINITY;
    }
    if (tid == 0) { s_front = 0; s_back = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, fmax, o);
        fmax = fmax > other ? fmax : other;
    }
    if (lane == 0) s_red[warp] = fmax;
    cs_sync<FUSED>();
    CS_TL(1);
    fmax = 0.0;
    for (int w = 0; w < NW; ++w) fmax = fmax > s_red[w] ? fmax : s_red[w];
    // (s_red is next written by the misfit sum, several barriers from here)
    // fixed-point scale: |sum of the terms of one ray| <= csc_bound * 2 fmax < 2^e  ->  quantum 2^(e - 61)
    const bool finite = fmax <= 1.79769313486231570815e308 && T.csc_bound * (2.0 * fmax) <= 1.79769313486231570815e308;
    int e2 = 0;
    if (finite && fmax > 0.0) (void)frexp(T.csc_bound * (2.0 * fmax), &e2);
    int sh = 61 - e2;
    sh = sh > 900 ? 900 : (sh < -900 ? -900 : sh);
    const double scale = ldexp(1.0, sh), inv_scale = ldexp(1.0, -sh);

    // ---- change-local rasterisation: which grid points get a new nearest site --------------------------------
    // (same rule as k_raster2d_inc in tomo_kernels.cu: a surviving owner is only challenged by the inserted cell;
    // points of the removed/moved cell are re-derived over all proposed sites; strict '<' with ascending index and
    // separately rounded products keep the lowest-index tie rule of numpy argmin / KDTree.query).
    // List entries (back << 31 | level << 29 | group << 12 | cell): points captured by the inserted cell carry their
    // OLD owner; points that lost their owner (back = 1, after the re-derivation) their NEW owner, the old one being
    // the removed cell.
    if (!FUSED && tid == CS_THREADS - 1) {
        // everything of the decision that does not need the new misfit (logarithms, the generator, other targets):
        // off the critical path, on the one warp without rasterisation work
        Rng rng;
        rng_resume(rng, ep, c, s_fin);
        finish_pre(ep, c, s_fin, rng, t, s_pre);
        rng_end(rng, ep, c);
        CS_TL_T(9, CS_THREADS - 1);
    }
    // Pass 1, one thread per group of 16 points (one 16-byte load of the index row): a value move lists the cell's own
    // points; a geometry move only SORTS the groups -- those that hold points of the removed / moved cell (found by
    // comparing indices) and those the inserted cell might reach (every group the bounding-box test cannot rule out)
    // are queued for pass 2, where 16 lanes share a group's points.  No distance is evaluated per point here, so the
    // pass costs the same whatever the move and wherever the new cell lies.
    for (int g0 = tid * CS_PTS; g0 < G && warp < RW; g0 += RW * 32 * CS_PTS) {
        const IdxVec<IdxT> wv = wv_next;
        if (g0 + RW * 32 * CS_PTS < G) wv_next.load(row + g0 + RW * 32 * CS_PTS);
        unsigned scn = 0;
        if (inner == BBB_MOVE_VALUES) {
            unsigned chg = 0;
#pragma unroll
            for (int q = 0; q < CS_PTS; ++q)
                if (g0 + q < G && wv.at(q) == rm) chg |= 1u << q;
            if (chg) {
                // merged groups: all 16 points, or an aligned group of 4, of the cell
                unsigned quads = 0;
                const bool hex = n_levels >= 3 && chg == 0xFFFFu;
                if (!hex && n_levels >= 2) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        if (((chg >> (4 * qd)) & 0xFu) == 0xFu) {
                            quads |= 1u << qd;
                            chg &= ~(0xFu << (4 * qd));
                        }
                    }
                }
                if (hex) {
                    const int pos = atomicAdd(&s_front, 1);
                    list_at(pos) = (2u << 29) | ((unsigned)(g0 >> 4) << CS_M_BITS) | (unsigned)rm;
                } else {
                    int pos = atomicAdd(&s_front, __popc(quads) + __popc(chg));
                    while (quads) {
                        const int qd = __ffs(quads) - 1;
                        quads &= quads - 1;
                        list_at(pos++) = (1u << 29) | ((unsigned)((g0 >> 2) + qd) << CS_M_BITS) | (unsigned)rm;
                    }
                    while (chg) {
                        const int q = __ffs(chg) - 1;
                        chg &= chg - 1;
                        list_at(pos++) = ((unsigned)(g0 + q) << CS_M_BITS) | (unsigned)rm;
                    }
                }
            }
            continue;
        }
        if (rm >= 0) {
#pragma unroll
            for (int q = 0; q < CS_PTS; ++q)
                if (g0 + q < G && wv.at(q) == rm) scn |= 1u << q;
        }
        // Pruning: a group is decided from its bounding box.  With o0 the current owner of ANY of its points,
        // d_own(g) <= d(g, o0) for every point g of the group (the owner is the nearest site), hence
        // d_new(g) - d_own(g) >= d_new(g) - d(g, o0) = |n|^2 - |o0|^2 - 2 g.(n - o0), which is affine in g: its
        // minimum over the box is taken at a corner, and when that minimum is positive by a margin far above any
        // rounding of the per-point test (1e-9 of the squared coordinates against 1e-16), no point of the group can
        // be captured by the inserted cell.
        bool reach = ins >= 0;
        if (ins >= 0 && T.group_box != nullptr && g0 + CS_PTS <= G) {
            int o0 = wv.at(0);
            if (rm >= 0 && o0 == rm) o0 = wv.at(CS_PTS - 1);
            if (!(rm >= 0 && o0 == rm)) {
                const double2 bx = __ldg(reinterpret_cast<const double2 *>(T.group_box) + 2 * (g0 >> 4));
                const double2 by = __ldg(reinterpret_cast<const double2 *>(T.group_box) + 2 * (g0 >> 4) + 1);
                int mm = o0 - ((rm >= 0 && o0 > rm) ? 1 : 0);
                mm += (mm >= ins) ? 1 : 0;
                const double ox = sx[mm], oy = sy[mm];
                const double ddx = nx - ox, ddy = ny - oy;
                const double gmax = (ddx > 0.0 ? bx.y : bx.x) * ddx + (ddy > 0.0 ? by.y : by.x) * ddy;
                const double nn = nx * nx + ny * ny, oo = ox * ox + oy * oy;
                const double ax = ::fmax(fabs(bx.x), fabs(bx.y)), ay = ::fmax(fabs(by.x), fabs(by.y));
                reach = !((nn - oo) - 2.0 * gmax > 1e-9 * (nn + oo + ax * ax + ay * ay));
            }
        }
        // queue item: group << 17 | reach << 16 | mask of the points that lost their owner
        if (scn || reach) queue_at(atomicAdd(&s_back, 1)) = ((unsigned)(g0 >> 4) << 17) | (reach ? 1u << 16 : 0u) | scn;
    }
    bool h_bad = false;
    if constexpr (HOSTED) h_bad = hosted_compare_prefetched(ep, io, io.in + c * io.stride, c, tid, CS_THREADS, h_pref);
    CS_TL(10);
    cs_sync<FUSED>();
    CS_TL(2);
    // (the queue is complete)
    // Pass 2: 16 lanes per queued group, lane q = point q of the group.  A point of the removed / moved cell is
    // re-derived over all proposed sites; any other point of a group within reach is tested against the inserted cell
    // alone (its owner survives: the same rule as k_raster2d_inc in tomo_kernels.cu -- strict '<' with ascending index and
    // separately rounded products keep the lowest-index tie rule of numpy argmin / KDTree.query).  List entries
    // (back << 31 | level << 29 | group << 12 | cell): points captured by the inserted cell carry their OLD owner; points
    // that lost their owner (back = 1) their NEW owner, the old one being the removed cell (a moved cell that still owns a
    // point changes nothing there: 2-D site move, rm == ins, values kept).  Aligned groups of 4 / 16 points that change
    // between the same two cells become one entry of a merged level.
    {
        const int n_items = s_back;
        const int q = tid & 15, half = lane >> 4;
        for (int base = 0; base < n_items; base += NTH / 16) {
            const int it = base + (tid >> 4);
            const unsigned item = it < n_items ? queue_at(it) : 0u;
            const int g0 = (int)(item >> 17) << 4;
            const bool live = it < n_items && g0 + q < G;
            const bool lost = live && ((item >> q) & 1u);
            const bool test = live && !lost && ((item >> 16) & 1u);
            int bi = -1;   // new owner of a point that lost its owner
            int old = -1;  // old owner of a point the inserted cell captures
            if (lost || test) {
                const double gx = T.points_x[g0 + q], gy = T.points_y[g0 + q];
                if (lost) {
                    double best = INFINITY;
                    for (int j = 0; j < k_new; ++j) {
                        const double dx = sx[j] - gx, dy = sy[j] - gy;
                        const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        if (d < best) { best = d; bi = j; }
                    }
                    if (rm == ins && bi == ins) bi = -1;  // unchanged
                } else {
                    const int o = (int)row[g0 + q];
                    int mm = o - ((rm >= 0 && o > rm) ? 1 : 0);
                    mm += (mm >= ins) ? 1 : 0;
                    const double dx = sx[mm] - gx, dy = sy[mm] - gy;
                    const double d_own = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    const double ex = nx - gx, ey = ny - gy;
                    const double d_new = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                    if (d_new < d_own || (d_new == d_own && ins < mm)) old = o;
                }
            }
            const unsigned full = 0xffffffffu;
            // the half-warp's 16 points: merged entries need every point of the aligned group to change alike
            const int key = bi >= 0 ? bi : (old >= 0 ? (old | 0x10000) : -1);  // (new owner) or (old owner, captured)
            const int k0 = __shfl_sync(full, key, half * 16);
            const unsigned all_same = __ballot_sync(full, key >= 0 && key == k0);
            const bool hex = n_levels >= 3 && ((all_same >> (half * 16)) & 0xFFFFu) == 0xFFFFu;
            const int kq = __shfl_sync(full, key, (lane & ~3));
            const unsigned quad_same = __ballot_sync(full, key >= 0 && key == kq);
            const bool quad = !hex && n_levels >= 2 && ((quad_same >> (lane & ~3)) & 0xFu) == 0xFu;
            const unsigned tagged = bi >= 0 ? (CS_BACK | (unsigned)bi) : (unsigned)(old >= 0 ? old : 0);
            unsigned entry = CS_NONE;
            if (hex) {
                if (q == 0) entry = tagged | (2u << 29) | ((unsigned)(g0 >> 4) << CS_M_BITS);
            } else if (quad) {
                if ((q & 3) == 0) entry = tagged | (1u << 29) | ((unsigned)((g0 + q) >> 2) << CS_M_BITS);
            } else if (key >= 0) {
                entry = tagged | ((unsigned)(g0 + q) << CS_M_BITS);
            }
            const unsigned emit = __ballot_sync(full, entry != CS_NONE);
            int pos = 0;
            if (lane == 0 && emit) pos = atomicAdd(&s_front, __popc(emit));
            pos = __shfl_sync(full, pos, 0);
            if (entry != CS_NONE) list_at(pos + __popc(emit & ((1u << lane) - 1u))) = entry;
        }
    }
    cs_sync<FUSED>();
    CS_TL(3);
    const int n_list = s_front;

    // ---- change-local forward: fixed-point accumulation of the ray changes of one ray tile ---------------------
    const int32_t *cp0 = T.csc_colptr[0], *cp1 = T.csc_colptr[1], *cp2 = T.csc_colptr[2];
    const unsigned short *rw0 = T.csc_rows[0], *rw1 = T.csc_rows[1], *rw2 = T.csc_rows[2];
    const double *dt0 = T.csc_data[0], *dt1 = T.csc_data[1], *dt2 = T.csc_data[2];
    const unsigned lo_s = (unsigned)__cvta_generic_to_shared(acc_lo), hi_s = (unsigned)__cvta_generic_to_shared(acc_hi);
    auto accumulate = [&](int tile) {
        for (int base = 0; base < n_list; base += NTH) {
            const int e = base + tid;
            PrepEntry pe{0, 0, 0.0};
            if (e < n_list) {
                const unsigned pk = list_at(e);
                {
                    const int lvl = (int)((pk >> 29) & 3u);
                    const unsigned grp = (pk >> CS_M_BITS) & CS_G_MASK;
                    const int m = (int)(pk & CS_M_MASK);
                    const int n_groups = (G + (1 << (2 * lvl)) - 1) >> (2 * lvl);
                    const int32_t *cp = (lvl == 0 ? cp0 : (lvl == 1 ? cp1 : cp2)) + (long long)tile * (n_groups + 1);
                    pe.j0 = cp[grp];
                    pe.dlt = (pk & CS_BACK) ? fn[m] - fo[rm] : fn[ins] - fo[m];
                    pe.len = (pe.dlt != 0.0) ? ((cp[grp + 1] - pe.j0) | (lvl << 30)) : 0;
                }
            }
            prep[tid] = pe;
            cs_sync<FUSED>();
            if (base == 0) CS_TL(4);
            const int n = min(NTH, n_list - base);
            // two columns per warp and round, four entries per lane and column: all loads of a round are issued
            // before its first atomic
            for (int i = warp; i < n; i += 2 * NW) {
                PrepEntry qa = prep[i];
                PrepEntry qb{0, 0, 0.0};
                if (i + NW < n) qb = prep[i + NW];
                const int la = qa.len >> 30, lb = qb.len >> 30;
                qa.len &= 0x3FFFFFFF;
                qb.len &= 0x3FFFFFFF;
                const unsigned short *rows_a = la == 0 ? rw0 : (la == 1 ? rw1 : rw2), *rows_b = lb == 0 ? rw0 : (lb == 1 ? rw1 : rw2);
                const double *data_a = la == 0 ? dt0 : (la == 1 ? dt1 : dt2), *data_b = lb == 0 ? dt0 : (lb == 1 ? dt1 : dt2);
                const int len_max = max(qa.len, qb.len);
                for (int base_j = 0; base_j < len_max; base_j += 128) {
                    // (all conditions on base_j are warp-uniform: sub-batches past a column's end are skipped, not
                    // predicated off)
                    int ra[4], rb[4];
                    double wa[4], wb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = base_j + 32 * u + lane;
                        ra[u] = rb[u] = 0;
                        wa[u] = wb[u] = 0.0;
                        if (base_j + 32 * u < qa.len && jj < qa.len) {
                            ra[u] = (int)rows_a[qa.j0 + jj];
                            wa[u] = data_a[qa.j0 + jj];
                        }
                        if (base_j + 32 * u < qb.len && jj < qb.len) {
                            rb[u] = (int)rows_b[qb.j0 + jj];
                            wb[u] = data_b[qb.j0 + jj];
                        }
                    }
                    if (base_j < qa.len) fx_add4(lo_s, hi_s, ra, wa, qa.dlt, scale, base_j, qa.len, lane);
                    if (base_j < qb.len) fx_add4(lo_s, hi_s, rb, wb, qb.dlt, scale, base_j, qb.len, lane);
                }
            }
            cs_sync<FUSED>();
        }
    };
    auto zero_acc = [&]() {
        uint4 *acc4 = reinterpret_cast<uint4 *>(acc_lo);  // (several tiles: RT is a multiple of 32)
#pragma unroll 4
        for (int i = tid; i < (2 * RT) >> 2; i += NTH) acc4[i] = make_uint4(0, 0, 0, 0);
        for (int r = ((2 * RT) & ~3) + tid; r < 2 * RT; r += NTH) acc_lo[r] = 0;
        cs_sync<FUSED>();
    };

    // ---- misfit: phi' - phi = sum_r w_r d_r (2 res_r + d_r) over the touched rays -----------------------------
    // Single-tile problems leave the NEW residual of every touched ray in the accumulator planes (untouched rays
    // keep their zeros), so an accepted proposal is committed with stores only.
    double *res = B.res[t] + c * (long long)R;
    double dphi = 0.0;
    // single-tile problems whose planes can be read four rays at a time: the touched rays are first compacted into a
    // list (the idle `prep` area), so that the arithmetic below runs on dense warps
    constexpr int CS_TOUCH_SMEM = CS_CHUNK * (int)sizeof(PrepEntry) / 2;
    unsigned short *touch = reinterpret_cast<unsigned short *>(prep);
    int n_touch = -1;  // >= 0: `touch` lists the touched rays of the (only) tile
    for (int tile = 0; tile < NT; ++tile) {
        if (MULTI) {
            // The parking pass below reads the residual of every touched ray of this tile.  A good part of the rays are
            // touched and 16 rays share a 128-byte line, so practically every line of the tile's stretch of the row is read
            // anyway: all of them are requested now, one prefetch per line, and the reads find them in the L2 instead of
            // waiting for DRAM (C5: 0.800 -> 0.791 ms per iteration; no effect on single-tile problems, where it is off).
            const int pr0 = tile * RT, pnr = min(RT, R - pr0);
            for (int i = tid; 16 * i < pnr; i += NTH) asm volatile("prefetch.global.L2 [%0];" ::"l"(res + pr0 + 16 * i));
        }
        if (tile > 0) zero_acc();
        accumulate(tile);
        CS_TL(5);
        const int r0 = tile * RT, nr = min(RT, R - r0);
        if (MULTI) {
            // several ray tiles recycle the accumulators: the new residuals of the touched rays are parked in the
            // chain's row of the (otherwise idle) d_pred scratch, the touched rays remembered in a bitmap
            double *park = B.dpred_scratch[t] + c * (long long)R;
            // eight 32-ray words per warp and round: the residuals of all eight are requested before the first is
            // used (one word at a time, every round waited for its own load); each thread still adds its rays in
            // ascending order, so the sum is the same
            const int n_words = (nr + 31) / 32;
            for (int w0 = warp; w0 < n_words && warp < CS_MAIN_WARPS; w0 += 8 * CS_MAIN_WARPS) {
                unsigned lo[8], hi[8];
                double rr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int w = w0 + u * CS_MAIN_WARPS, r = 32 * w + lane;
                    const bool in = w < n_words && r < nr;
                    lo[u] = in ? acc_lo[r] : 0u;
                    hi[u] = in ? acc_hi[r] : 0u;
                    const bool touched = (lo[u] | hi[u]) != 0;
                    const unsigned m = __ballot_sync(0xffffffffu, touched);
                    if (lane == 0 && w < n_words) s_touch[(r0 >> 5) + w] = m;
                    rr[u] = touched ? res[r0 + r] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if ((lo[u] | hi[u]) == 0) continue;
                    const int r = 32 * (w0 + u * CS_MAIN_WARPS) + lane;
                    const double d = (double)(long long)(((unsigned long long)hi[u] << 32) | lo[u]) * inv_scale;
                    const double wr = (T.cov_kind == BBB_COV_VECTOR) ? T.cov_vec[r0 + r] : 1.0;
                    dphi += wr * (d * (2.0 * rr[u] + d));
                    park[r0 + r] = rr[u] + d;
                }
            }
            cs_sync<FUSED>();
            continue;
        }
        const bool pairs = ((R | r0 | nr | RT) & 1) == 0;  // rays two at a time: 16-byte loads of the residual row
        auto one_ray = [&](int r, double rr) {
            const unsigned lo = acc_lo[r], hi = acc_hi[r];
            if ((lo | hi) != 0) {
                const double d = (double)(long long)(((unsigned long long)hi << 32) | lo) * inv_scale;
                const double wr = (T.cov_kind == BBB_COV_VECTOR) ? T.cov_vec[r0 + r] : 1.0;
                dphi += wr * (d * (2.0 * rr + d));  // w (res + d)^2 - w res^2
                double nv = rr + d;
                if (nv == 0.0) nv = -0.0;  // all-zero planes mean "untouched"; -0.0 is the same residual
                acc_lo[r] = (unsigned)__double2loint(nv);
                acc_hi[r] = (unsigned)__double2hiint(nv);
            }
        };
        if (!MULTI && (RT & 3) == 0 && RT <= 8 * 4 * CS_MAIN) {
            // (a) every thread of the first CS_MAIN_WARPS warps scans 4 consecutive rays per round, bit (4 i + j) of m:
            // ray 4 tid + 4 CS_MAIN i + j touched
            unsigned m = 0;
#pragma unroll 8
            for (int i = 0; i < 8; ++i) {
                const int r4 = 4 * tid + 4 * CS_MAIN * i;
                if (r4 < nr && tid < CS_MAIN) {
                    const uint4 lo4 = *reinterpret_cast<const uint4 *>(acc_lo + r4), hi4 = *reinterpret_cast<const uint4 *>(acc_hi + r4);
                    m |= (((lo4.x | hi4.x) != 0 ? 1u : 0u) | ((lo4.y | hi4.y) != 0 ? 2u : 0u) | ((lo4.z | hi4.z) != 0 ? 4u : 0u) |
                          ((lo4.w | hi4.w) != 0 ? 8u : 0u)) << (4 * i);
                }
            }
            // (b) list positions from a scan over (warp, lane): the order of the list -- and with it the order of the
            // misfit sum -- does not depend on thread timing
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_wtot[warp] = incl;
            cs_sync<FUSED>();
            int pos = incl - cnt;
            n_touch = 0;
#pragma unroll
            for (int w = 0; w < CS_MAIN_WARPS; ++w) {
                const int wt = s_wtot[w];
                if (w < warp) pos += wt;
                n_touch += wt;
            }
            if (n_touch <= CS_TOUCH_SMEM) {
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    touch[pos++] = (unsigned short)(4 * tid + 4 * CS_MAIN * (b >> 2) + (b & 3));
                }
                cs_sync<FUSED>();
                // (c) the touched rays, four per thread in flight
                for (int i0 = tid; i0 < n_touch && tid < CS_MAIN; i0 += 4 * CS_MAIN) {
                    int rq[4];
                    double rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * CS_MAIN;
                        rq[u] = i < n_touch ? (int)touch[i] : -1;
                        rr[u] = rq[u] >= 0 ? res[r0 + rq[u]] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (rq[u] >= 0) one_ray(rq[u], rr[u]);
                }
            } else {  // (more touched rays than the list holds: every thread handles the rays it scanned)
                n_touch = -1;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int r = 4 * tid + 4 * CS_MAIN * (b >> 2) + (b & 3);
                    one_ray(r, res[r0 + r]);
                }
            }
        } else if (pairs) {
            // CS_MISFIT_U 16-byte loads of the residual row in flight per thread (C3: the whole row in one round)
            for (int rb = 2 * tid; rb < nr && tid < CS_MAIN; rb += 2 * CS_MISFIT_U * CS_MAIN) {
                double2 rr[CS_MISFIT_U];
#pragma unroll
                for (int u = 0; u < CS_MISFIT_U; ++u) {
                    const int r = rb + 2 * u * CS_MAIN;
                    rr[u] = r < nr ? *reinterpret_cast<const double2 *>(res + r0 + r) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < CS_MISFIT_U; ++u) {
                    const int r = rb + 2 * u * CS_MAIN;
                    if (r >= nr) continue;
                    const uint2 lo2 = *reinterpret_cast<const uint2 *>(acc_lo + r), hi2 = *reinterpret_cast<const uint2 *>(acc_hi + r);
                    if ((lo2.x | lo2.y | hi2.x | hi2.y) == 0) continue;
                    one_ray(r, rr[u].x);
                    one_ray(r + 1, rr[u].y);
                }
            }
        } else {
            for (int rb = tid; rb < nr && tid < CS_MAIN; rb += 4 * CS_MAIN) {
                double rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * CS_MAIN;
                    rr[u] = r < nr ? res[r0 + r] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * CS_MAIN;
                    if (r < nr) one_ray(r, rr[u]);
                }
            }
        }
        if (MULTI) cs_sync<FUSED>();
    }
    CS_TL(11);
    dphi = block_sum_to_thread0<FUSED>(dphi, s_red);
    CS_TL(6);

    // ---- likelihood ratio, decision, statistics (thread 0), then the commit by the whole block ----------------
    if constexpr (HOSTED) {
        // the uploaded state differs from the one the caches belong to: nothing is decided, nothing committed
        if (__syncthreads_or(h_bad ? 1 : 0)) {
            if (tid == 0) { io.mismatch[c] = 1; *io.host_flag = 1; hosted_signal_done(io, c); }
            return -1;
        }
    }
    if (tid == 0) {
        s_accept = finish_post(ep, c, fin, pre, t, finite ? fin.phi_old[t].s0 + dphi : NAN) ? 1 : 0;
        // the chain's new STATE is complete here (the commit below only updates the caches of the next chain kernel):
        // the next iteration's proposal warp for this chain may go
        if constexpr (HOSTED) hosted_signal_done(io, c);
    }
    cs_sync<FUSED>();
    CS_TL(7);
    if (tid == 0 && blockIdx.x < 4096) {
#ifdef CS_TIMELINE
        g_timeline[blockIdx.x][12] = n_list; g_timeline[blockIdx.x][13] = s_back; g_timeline[blockIdx.x][14] = s_accept;
        g_timeline[blockIdx.x][15] = k_new; g_timeline[blockIdx.x][19] = move;
#endif
    }
    CS_TL_G(17);
    const int planes = MULTI ? -1 : n_touch;
    if (!s_accept) {
        if constexpr (HOSTED) hosted_write(ep, io, c, tid, CS_THREADS, h_words, s, cur, k_old);
        return planes;
    }

    if (!MULTI && n_touch >= 0) {
        for (int i = tid; i < n_touch; i += NTH) {
            const int r = (int)touch[i];
            res[r] = __hiloint2double((int)acc_hi[r], (int)acc_lo[r]);
        }
    } else if (!MULTI) {
        for (int r = tid; r < R; r += NTH) {
            const unsigned lo = acc_lo[r], hi = acc_hi[r];
            if ((lo | hi) != 0) res[r] = __hiloint2double((int)hi, (int)lo);
        }
    } else {
        const double *park = B.dpred_scratch[t] + c * (long long)R;
        // (eight words per warp and round: all loads of a round ahead of its stores)
        const int n_words = (R + 31) / 32;
        for (int w0 = warp; w0 < n_words; w0 += 8 * NW) {
            double v[8];
            bool on[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int w = w0 + u * NW;
                on[u] = w < n_words && ((s_touch[w] >> lane) & 1u);
                v[u] = on[u] ? park[32 * w + lane] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (on[u]) res[32 * (w0 + u * NW) + lane] = v[u];
        }
    }
    // nearest-site index: renumber the survivors when the edit shifts cell indices (death), then the changed points
    constexpr int VEC = 16 / (int)sizeof(IdxT);
    const bool renumber = (rm >= 0 || ins >= 0) && rm != ins && !(rm < 0 && ins >= k_old);
    if (renumber) {
        for (int g0 = tid * VEC; g0 < G; g0 += NTH * VEC) {
            uint4 wv = *reinterpret_cast<const uint4 *>(row + g0);
            IdxT *ow = reinterpret_cast<IdxT *>(&wv);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const int o = (int)ow[q];
                int mm = o - ((rm >= 0 && o > rm) ? 1 : 0);
                mm += (ins >= 0 && mm >= ins) ? 1 : 0;
                ow[q] = (IdxT)mm;
            }
            *reinterpret_cast<uint4 *>(row + g0) = wv;
        }
        cs_sync<FUSED>();
    }
    for (int e = tid; e < n_list; e += NTH) {
        const unsigned pk = list_at(e);
        const int lvl = (int)((pk >> 29) & 3u);
        const int g_first = (int)((pk >> CS_M_BITS) & CS_G_MASK) << (2 * lvl);
        const IdxT owner = (IdxT)((pk & CS_BACK) ? (pk & CS_M_MASK) : (unsigned)ins);
        for (int q = 0; q < (1 << (2 * lvl)); ++q) row[g_first + q] = owner;
    }
    if constexpr (HOSTED) hosted_write(ep, io, c, tid, CS_THREADS, h_words, s, prop, k_new);
    CS_TL(8);
    CS_TL_G(17);
    return planes;
}

// One block per chain.  -DCS_PERSISTENT=1 (developer switch) launches 2 blocks per SM that take chains from a work
// counter (the 16th word of the launch's block-order class counts, zeroed by the proposal kernel) in longest-first
// order; measured SLOWER than letting the hardware dispatch one block per chain (chain kernel 0.111 vs 0.103 ms: the
// atomic, two more barriers per chain and the in-loop zeroing cost more than the block dispatch they replace).
template <typename IdxT, bool MULTI, bool HOSTED>
__global__ void __launch_bounds__(CS_THREADS, 2) k_chain_step(const __grid_constant__ EngineParams ep, int t, long long c_first,
                                                              int order_slot, const int32_t *guard, int n_launch,
                                                              const __grid_constant__ HostedIO io, int proposer_finished, int list_cap) {
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char cs_smem[];
    __shared__ int s_work, s_cum[BBB_ORDER_BUCKETS + 1];
    // Programmatic dependent launch: the proposal kernel releases its dependents as soon as it starts, so this block
    // may be resident while the proposals are still being drawn.  Everything that does not read the proposal is done
    // here, ahead of the dependency wait: zeroing the ray accumulators.
    auto zero_planes = [&]() {
        const int n_words = 2 * ep.spec.targets[t].csc_tile_rays;
        uint4 *acc4 = reinterpret_cast<uint4 *>(cs_smem);
        const int n4 = n_words >> 2;
#pragma unroll 4
        for (int i = tid; i < n4; i += CS_THREADS) acc4[i] = make_uint4(0, 0, 0, 0);
        unsigned *acc = reinterpret_cast<unsigned *>(cs_smem);
        for (int q = (n4 << 2) + tid; q < n_words; q += CS_THREADS) acc[q] = 0;
    };
    zero_planes();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // longest-first block order (common.cuh): work item b is the b-th chain in cost-class order; class counts that do
    // not add up to the launch (a proposal kernel that ran twice, a caller that binds no order buffer) mean index order
    bool ordered = false;
    int32_t *counter = nullptr;
    if (order_slot >= 0) {
        int32_t *cnt = order_counts(B.block_order, C, order_slot >> 2, order_slot & 3);
        counter = cnt + 15;
        if (tid == 0) {
            int tot = 0;
            for (int i = 0; i < BBB_ORDER_BUCKETS; ++i) { s_cum[i] = tot; tot += cnt[i]; }
            s_cum[BBB_ORDER_BUCKETS] = tot;
        }
        if (blockIdx.x == 0 && tid < 15) order_counts(B.block_order, C, (order_slot >> 2) ^ 1, order_slot & 3)[tid] = 0;
        __syncthreads();
        ordered = s_cum[BBB_ORDER_BUCKETS] == n_launch;
    }
    if constexpr (HOSTED) {
        // the next iteration's proposal kernel is this launch's programmatic dependent: it may become resident once every
        // block of this grid has got here (so it can never keep a block of this grid from being scheduled); its warps wait
        // for their own chain's completion signal below, not for the whole grid
        asm volatile("griddepcontrol.launch_dependents;");
    }
    const bool persistent = CS_PERSISTENT && counter != nullptr && (int)gridDim.x < n_launch;
    for (int iter = 0;; ++iter) {
        int b = (int)blockIdx.x;
        if (persistent) {
            if (tid == 0) s_work = atomicAdd(counter, 1);
            __syncthreads();
            b = s_work;
            if (b >= n_launch) break;
            if (iter > 0) zero_planes();
        } else if (iter > 0) {
            break;
        }
        long long c = c_first + b;
        if (ordered) {
            int cls = 0;
#pragma unroll
            for (int i = 1; i < BBB_ORDER_BUCKETS; ++i) cls += b >= s_cum[i] ? 1 : 0;
            c = B.block_order[(long long)cls * C + c_first + (b - s_cum[cls])];
        }
        chain_step_body<IdxT, MULTI, false, HOSTED>(ep, t, c, guard, cs_smem, nullptr, io, proposer_finished, list_cap);
        if (persistent) __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// The fused chain-local run: proposal + chain step + decision of MANY iterations in ONE persistent launch
// (`BaseMarkovChain.advance_chain`, src/bayesbay/_markov_chain.py:284-295 -- the per-chain loop -- for all chains).
//
// Two blocks per SM stay resident for the whole launch, so there is no per-iteration launch, no wave tail per
// iteration, and the slow chains of one iteration overlap the fast chains of the next.  Work = (chain, its next
// iteration); a chain is READY when its previous iteration is complete.  Ready chains wait in a global FIFO (ctl):
// tickets from a `head` counter, entries published with release stores by whoever completes a chain's iteration
// (ticket h < n_chains is chain h itself: every chain starts ready).  A chain is therefore never claimed before it
// can run, whatever the number of proposals in flight.
// A block is warp-specialised:
//   * the last CS_PROPOSERS warps, the PROPOSERS, each take a ticket when one of their two records is free, draw
//     the chain's proposal exactly as k_propose does, and
//       - decide proposals that need no forward (noise moves, proposals the prior rejects) on their own,
//       - hand the others to the workers (record index into a block-level FIFO in shared memory, mbarrier per FIFO
//         position / per record), together with everything of the decision that does not need the new misfit;
//     (CS_N_PROPOSERS is a compile-time choice: one proposer needs ~34 us of its own time per forward item against
//     the workers' ~22 us; two are slower still, because every further instruction stream lowers the hit rate of the
//     SM's instruction cache for all of them -- see the measurement note below);
//   * the first CS_MAIN_WARPS warps, the WORKERS, run chain_step_body on the records in arrival order: change-local
//     rasterisation, forward, misfit, decision, commit.  They clear only the accumulator words the item touched.
// Results are bit-identical to k_propose + k_chain_step (same draws, same arithmetic, same summation order); they do
// not depend on which block runs which item.
// MEASURED (B200, C3, 1024 chains, profiles/r2_fused_run_ncu.md): 0.109 ms per iteration with one proposer per block,
// 0.122 with two, against 0.083 for k_propose + k_chain_step pipelined on four streams.  The kernel is bound by
// instruction fetch: the proposal is ~3 000 executed instructions of straight-line code per item (7 000 in the
// binary), the SM's L1.5 instruction cache holds 2 000, and proposers and workers evict each other (instruction-cache
// hit rate 52 %, `no_instruction` the second largest stall reason, GPC instruction-fetch path at 77 % of its peak).
// It therefore stays behind the developer switch BBB_FUSED_RUN=1; the default step keeps the two kernels apart.
// ctl (device int32 words, zero when the launch starts, zero again when it ends): [0] head (tickets), [1] blocks done,
// [2] tail (entries published), [4 + c] iterations chain c completed in this launch, then (8-byte aligned) n_chains
// 64-bit FIFO entries (ticket tag << 32 | chain).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
// (polling with acquire loads would invalidate the SM's L1 on every probe and slow every other warp of the SM down:
// the FIFO entry is polled with relaxed loads, ONE fence follows the hit)
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifdef CS_TIMELINE
__device__ long long g_runstats[1024][16];
#define RS_ADD(i, v) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 1024) atomicAdd((unsigned long long *)&g_runstats[blockIdx.x][i], (unsigned long long)(v)); } while (0)
#define RS_CLK() clock64()
#else
#define RS_ADD(i, v) do { (void)(v); } while (0)
#define RS_CLK() 0LL
#endif

constexpr int CS_RECS = 2 * CS_PROPOSERS;  // records of a block: two per proposer (one at / waiting for the workers, one being drawn)

struct RunShared {
    unsigned long long ready[CS_RECS];  // block FIFO position filled
    unsigned long long empty[CS_RECS];  // record released by the workers
    int fifo[CS_RECS];
    int tail;
    int last;
    double draw[CS_PROPOSERS][2][32];
    Edit edit[CS_PROPOSERS];  // the proposers' edit scripts (all lanes of a proposer build the same one)
    ItemRec rec[CS_RECS];
};

__host__ __device__ inline long long run_ctl_queue_offset(long long n_chains_engine) { return (4 + n_chains_engine + 1) & ~1LL; }

// a chain's iteration is complete (one thread, after a barrier / warp sync behind the item's last global store):
// count it and, unless it was the chain's last one of this launch, publish the chain as ready
__device__ __forceinline__ void run_complete(int32_t *ctl, unsigned long long *queue, long long c, int n_chains, int n_iters) {
    const int done = ctl[4 + c] + 1;
    ctl[4 + c] = done;
    if (done >= n_iters) return;
    __threadfence();
    const unsigned j = atomicAdd(reinterpret_cast<unsigned *>(ctl + 2), 1u);
    st_release_gpu_u64(queue + (j % (unsigned)n_chains), ((unsigned long long)(j + 1u) << 32) | (unsigned long long)(unsigned)c);
}

__device__ __noinline__ void run_proposer(const EngineParams &ep, int t, long long c_first, int n_chains, int n_iters, int32_t *ctl,
                                          RunShared &sh, int pw) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const long long C = B.n_chains;
    const int lane = threadIdx.x & 31;
    const unsigned total = (unsigned)n_chains * (unsigned)n_iters;
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(ctl + run_ctl_queue_offset(C));
    int produced = 0;
    auto hand_over = [&](int r) {  // lane 0: record r into the block's FIFO
        const int pos = atomicAdd(&sh.tail, 1) % CS_RECS;
        sh.fifo[pos] = r;
        mbar_arrive(&sh.ready[pos]);
    };
    for (;;) {
        // a ticket is taken only when a record is free for the chain it brings
        const int r = 2 * pw + (produced & 1);
        if (produced >= 2) {
            const long long t0 = RS_CLK();
            mbar_wait(&sh.empty[r], (unsigned)((produced >> 1) - 1) & 1u);
            RS_ADD(3, RS_CLK() - t0);
        }
        ItemRec &rec = sh.rec[r];
        unsigned h = 0;
        if (lane == 0) h = atomicAdd(reinterpret_cast<unsigned *>(ctl), 1u);
        h = __shfl_sync(0xffffffffu, h, 0);
        if (h >= total) {
            if (lane == 0) {  // end mark
                rec.c = -1;
                hand_over(r);
            }
            return;
        }
        RS_ADD(4, 1);
        long long c = c_first + h;
        if (h >= (unsigned)n_chains) {  // the (h - n_chains)-th chain to become ready again
            const long long t0 = RS_CLK();
            const unsigned j = h - (unsigned)n_chains;
            unsigned cc = 0;
            if (lane == 0) {
                const unsigned long long *slot = queue + (j % (unsigned)n_chains);
                unsigned long long v;
                while ((unsigned)((v = ld_relaxed_gpu_u64(slot)) >> 32) != j + 1u) __nanosleep(200);
                cc = (unsigned)v;
                __threadfence();  // acquire: the chain's state as its last iteration left it
            }
            c = (long long)__shfl_sync(0xffffffffu, cc, 0);
            RS_ADD(0, RS_CLK() - t0);
        }
        const long long t_prop = RS_CLK();
        FastRng rng;
        rng.begin(sh.draw[pw][0], sh.draw[pw][1], B.tape != nullptr ? B.tape + c * B.tape_stride : nullptr,
                  B.tape != nullptr ? B.tape_cursor[c] : 0, B.tape_stride, P.seed, (uint64_t)(B.chain_id0 + c), (uint64_t)B.iteration[c],
                  STREAM_STEP);
        Edit &ed = sh.edit[pw];
        propose_chain(P, B, c, rng, false, ed);
        const long long cursor = rng.after();
        if (lane == 0) B.tape_cursor[c] = cursor;
        __syncwarp();
        const int move = ed.move;  // (every lane built the same script)
        const double lpr = ed.lpr;
        RS_ADD(1, RS_CLK() - t_prop);
        if (!target_touched(P, move, lpr, t)) {  // no forward: the whole decision is this warp's
            const long long t0 = RS_CLK();
            RS_ADD(5, 1);
            if (lane == 0) {  // (finish_chain with its working set in the free record instead of local memory)
                rec.fin.move = move;
                rec.fin.lpr = lpr;
                rec.fin.cursor = cursor;
                finish_gather_rest(ep, c, rec.fin);
                Rng r2;
                rng_resume(r2, ep, c, rec.fin);
                finish_pre(ep, c, rec.fin, r2, -1, rec.pre);
                rng_end(r2, ep, c);
                finish_post(ep, c, rec.fin, rec.pre, -1, 0.0);
                run_complete(ctl, queue, c, n_chains, n_iters);
            }
            __syncwarp();
            RS_ADD(6, RS_CLK() - t0);
            continue;
        }
        const long long t_fill = RS_CLK();
        if (lane == 0) {
            rec.c = c;
            rec.move = move;
            rec.lpr = lpr;
            rec.cur = ed.cur;
            rec.k_old = ed.k_old;
            rec.k_new = ed.k_new;
            rec.rm = ed.rm;
            rec.ins = ed.ins;
            rec.nx = ed.site[0];
            rec.ny = ed.site[1];
            rec.fin.move = move;
            rec.fin.lpr = lpr;
            rec.fin.cursor = cursor;
            finish_gather_rest(ep, c, rec.fin);
            Rng r2;
            rng_resume(r2, ep, c, rec.fin);
            finish_pre(ep, c, rec.fin, r2, t, rec.pre);
            rng_end(r2, ep, c);
        }
        __syncwarp();  // (the lanes' stores of the proposed slot are ordered before the hand-over)
        RS_ADD(2, RS_CLK() - t_fill);
        if (lane == 0) hand_over(r);
        ++produced;
    }
}

template <typename IdxT, bool MULTI>
__global__ void __launch_bounds__(CS_THREADS, 2) k_chain_run(const __grid_constant__ EngineParams ep, int t, long long c_first,
                                                             int n_chains, int n_iters, int32_t *ctl) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    __shared__ RunShared sh;
    const int tid = threadIdx.x;
    const int n_words = 2 * ep.spec.targets[t].csc_tile_rays;
    auto zero_planes = [&](int n_threads) {
        uint4 *acc4 = reinterpret_cast<uint4 *>(cs_smem);
        const int n4 = n_words >> 2;
#pragma unroll 4
        for (int i = tid; i < n4; i += n_threads) acc4[i] = make_uint4(0, 0, 0, 0);
        unsigned *acc = reinterpret_cast<unsigned *>(cs_smem);
        for (int q = (n4 << 2) + tid; q < n_words; q += n_threads) acc[q] = 0;
    };
    zero_planes(CS_THREADS);
    if (tid == 0) {
        for (int i = 0; i < CS_RECS; ++i) { mbar_init(&sh.ready[i], 1); mbar_init(&sh.empty[i], 1); }
        sh.tail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(ctl + run_ctl_queue_offset(ep.buf.n_chains));
    if (tid >= CS_MAIN) {
        run_proposer(ep, t, c_first, n_chains, n_iters, ctl, sh, (tid - CS_MAIN) >> 5);
    } else {
        unsigned *acc_lo = reinterpret_cast<unsigned *>(cs_smem);
        unsigned *acc_hi = acc_lo + ep.spec.targets[t].csc_tile_rays;
        // (the touched-ray list of the body: the `prep` area behind the planes and the four cell tables)
        const unsigned short *touch = reinterpret_cast<const unsigned short *>(
            cs_smem + (size_t)ep.spec.targets[t].csc_tile_rays * 8 + (size_t)ep.spec.spaces[ep.spec.targets[t].space].kmax * 32);
        const long long t_start = RS_CLK();
        int ended = 0;
        for (int i = 0; ended < CS_PROPOSERS; ++i) {
            const int pos = i % CS_RECS;
            // (one thread watches the barrier of the FIFO position, the others sleep in the named barrier instead of polling)
            const long long t0 = RS_CLK();
            if (tid == 0) mbar_wait(&sh.ready[pos], (unsigned)(i / CS_RECS) & 1u);
            cs_sync<true>();
            const long long t1 = RS_CLK();
            const int r = sh.fifo[pos];
            ItemRec *rec = &sh.rec[r];
            const long long c = rec->c;
            if (c < 0) {  // a proposer's end mark
                ++ended;
                cs_sync<true>();  // (everyone has read the position before its next use)
                continue;
            }
            const int planes = chain_step_body<IdxT, MULTI, true>(ep, t, c, nullptr, cs_smem, rec, HostedIO{});
            const long long t2 = RS_CLK();
            // leave the planes zero for the next item: only the words this item touched, when the body listed them
            if (planes >= 0) {
                for (int q = tid; q < planes; q += CS_MAIN) {
                    const int rr = (int)touch[q];
                    acc_lo[rr] = 0;
                    acc_hi[rr] = 0;
                }
            } else if (planes == -1) {
                zero_planes(CS_MAIN);
            }
            cs_sync<true>();  // every worker is done with the record, the lists and its global stores
            if (tid == 0) {
                run_complete(ctl, queue, c, n_chains, n_iters);
                mbar_arrive(&sh.empty[r]);
                RS_ADD(8, t1 - t0);
                RS_ADD(9, t2 - t1);
                RS_ADD(10, RS_CLK() - t2);
                RS_ADD(11, 1);
            }
        }
        if (tid == 0) RS_ADD(12, RS_CLK() - t_start);
    }
    __syncthreads();
    // the last block to finish leaves the control words zero for the next launch
    if (tid == 0) {
        __threadfence();
        sh.last = atomicAdd(reinterpret_cast<unsigned *>(ctl + 1), 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (sh.last) {
        __threadfence();
        const long long n_ctl = run_ctl_queue_offset(ep.buf.n_chains) + 2 * (long long)ep.buf.n_chains;
        for (long long i = tid; i < n_ctl; i += CS_THREADS) ctl[i] = 0;
    }
}

long long chain_run_ctl_words(long long n_chains_engine) { return run_ctl_queue_offset(n_chains_engine) + 2 * n_chains_engine; }

int launch_chain_run(const EngineParams &dev, long long c_first, long long c_end, long long n_iters, int t, int tile_rays,
                     long long n_data, int K, int idx_bytes, int32_t *ctl, int n_blocks, cudaStream_t st) {
    const long long n_chains = c_end - c_first;
    if (n_chains < 1 || n_iters < 1) return 0;
    if (n_chains * n_iters >= (1LL << 31) - 65536) return (int)cudaErrorInvalidValue;
    const size_t smem = chain_step_smem(tile_rays, K, n_data);
    const bool multi = n_data > tile_rays;
    long long grid = n_blocks;
    if (grid > n_chains) grid = n_chains;
    auto launch = [&](auto kernel) -> int {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        kernel<<<dim3((unsigned)grid), dim3(CS_THREADS), smem, st>>>(dev, t, c_first, (int)n_chains, (int)n_iters, ctl);
        return (int)cudaGetLastError();
    };
    if (idx_bytes == 1) return multi ? launch(k_chain_run<uint8_t, true>) : launch(k_chain_run<uint8_t, false>);
    return multi ? launch(k_chain_run<uint16_t, true>) : launch(k_chain_run<uint16_t, false>);
}

#ifdef CS_TIMELINE
extern "C" int bbb_debug_timeline(long long *host_out, int n_blocks) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 24 * (size_t)n_blocks);
}
extern "C" int bbb_debug_runstats(long long *host_out, int clear) {
    int rc = (int)cudaMemcpyFromSymbol(host_out, g_runstats, sizeof(g_runstats));
    if (rc == 0 && clear) {
        void *p = nullptr;
        cudaGetSymbolAddress(&p, g_runstats);
        rc = (int)cudaMemset(p, 0, sizeof(g_runstats));
    }
    return rc;
}
extern "C" int bbb_debug_timeline_clear() {
    void *p = nullptr;
    cudaGetSymbolAddress(&p, g_timeline);
    return (int)cudaMemset(p, 0, sizeof(g_timeline));
}
#endif

int launch_chain_step(const EngineParams &dev, long long c_first, long long c_end, int t, int tile_rays, long long n_data, int K,
                      int idx_bytes, int order_slot, bool dependent_launch, const int32_t *guard, cudaStream_t st, const HostedIO *hosted,
                      bool proposer_finished) {
    if (dev.buf.block_order == nullptr) order_slot = -1;
    // programmatic dependent launch behind the proposal kernel of the same stream (see the kernel's prologue)
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    // persistent launch: two blocks per SM (what registers and shared memory allow) take the chains from a work counter
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long n_launch = c_end - c_first;
    const bool persistent = CS_PERSISTENT && order_slot >= 0 && n_launch > 2LL * n_sm;
    cfg.gridDim = dim3((unsigned)(persistent ? 2 * n_sm : n_launch));
    cfg.blockDim = dim3(CS_THREADS);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = dependent_launch ? 1 : 0;
    const bool multi = n_data > tile_rays;
    const HostedIO hio = hosted != nullptr ? *hosted : HostedIO{};
    auto launch = [&](auto kernel) -> int {
        // (multi-tile problems size their ray tiles to fill the block's budget: nothing to trim there)
        int list_cap = CS_LIST_SMEM;
        if (!multi) {
            static size_t static_smem = (size_t)-1;  // (one per kernel instantiation: the lambda is generic)
            if (static_smem == (size_t)-1) {
                cudaFuncAttributes fa{};
                cudaError_t e0 = cudaFuncGetAttributes(&fa, kernel);
                if (e0 != cudaSuccess) return (int)e0;
                static_smem = fa.sharedSizeBytes;
            }
            list_cap = fit_list_cap(chain_step_smem(tile_rays, K, n_data, 0), static_smem);
        }
        const size_t smem = chain_step_smem(tile_rays, K, n_data, list_cap);
        cfg.dynamicSmemBytes = smem;
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        return (int)cudaLaunchKernelEx(&cfg, kernel, dev, t, c_first, order_slot, guard, (int)n_launch, hio, proposer_finished ? 1 : 0, list_cap);
    };
    int rc;
    if (hosted != nullptr) {
        if (idx_bytes == 1) rc = multi ? launch(k_chain_step<uint8_t, true, true>) : launch(k_chain_step<uint8_t, false, true>);
        else rc = multi ? launch(k_chain_step<uint16_t, true, true>) : launch(k_chain_step<uint16_t, false, true>);
    } else if (idx_bytes == 1) rc = multi ? launch(k_chain_step<uint8_t, true, false>) : launch(k_chain_step<uint8_t, false, false>);
    else rc = multi ? launch(k_chain_step<uint16_t, true, false>) : launch(k_chain_step<uint16_t, false, false>);
    if (rc != 0) return rc;
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Layout changes between the chain-fastest arrays of the SpMM / from-scratch rasteriser and the chain-major
// caches of the chain-local step: out[j * out_stride + i] = in[i * in_stride + j] (- sub[i]).
// ------------------------------------------------------------------------------------------------
template <typename Tv, bool SUB>
__global__ void k_transpose(const Tv *in, Tv *out, long long rows, long long cols, long long in_stride,
                            long long out_stride, const Tv *sub) {
    __shared__ Tv tile[32][33];
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long i = i0 + y, j = j0 + threadIdx.x;
        if (i < rows && j < cols) {
            Tv v = in[i * in_stride + j];
            if (SUB) v = v - sub[i];
            tile[y][threadIdx.x] = v;
        }
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long j = j0 + y, i = i0 + threadIdx.x;
        if (i < rows && j < cols) out[j * out_stride + i] = tile[threadIdx.x][y];
    }
}

template <typename Tv, bool SUB>
static int launch_transpose_t(const Tv *in, Tv *out, long long rows, long long cols, long long in_stride,
                              long long out_stride, const Tv *sub, cudaStream_t st) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (grid.y > 65535) return (int)cudaErrorInvalidConfiguration;
    k_transpose<Tv, SUB><<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols, in_stride, out_stride, sub);
    return (int)cudaGetLastError();
}

int launch_transpose_idx(const void *in, void *out, int idx_bytes, long long rows, long long cols, long long in_stride,
                         long long out_stride, cudaStream_t st) {
    if (idx_bytes == 1)
        return launch_transpose_t<uint8_t, false>((const uint8_t *)in, (uint8_t *)out, rows, cols, in_stride, out_stride, nullptr, st);
    return launch_transpose_t<uint16_t, false>((const uint16_t *)in, (uint16_t *)out, rows, cols, in_stride, out_stride, nullptr, st);
}

// out[j][i] = in[i][j] (dense)
int launch_transpose_f64(const double *in, double *out, long long rows, long long cols, cudaStream_t st) {
    return launch_transpose_t<double, false>(in, out, rows, cols, cols, rows, nullptr, st);
}

// res[c][r] = dpred[r][c] - dobs[r]
int launch_residual_transpose(const double *dpred, double *res, long long R, long long C, const double *dobs, cudaStream_t st) {
    return launch_transpose_t<double, true>(dpred, res, R, C, C, R, dobs, st);
}

}  // namespace bbb
