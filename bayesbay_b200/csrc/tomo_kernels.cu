// F1 (2-D Voronoi nearest-site rasterisation) and F2 (straight-ray forward as a CSR ray-matrix
// SpMM with chains as the dense, coalesced dimension, fused with the Gaussian misfit reduction).
//
// Layout: every per-chain array is [..][C] with the chain index fastest, so a warp = 32 chains
// reads/writes one contiguous 32 B (uint8), 128 B (int32) or 256 B (fp64) segment per access.
#include "kernels.h"

namespace bbb {

// ------------------------------------------------------------------------------------------------
// F1: brute-force nearest site.  One lane = one chain, one warp = 32 chains x a strip of grid
// points; the 32 chains' site rows come from L1/L2 as coalesced 256 B segments.  Squared distance
// dx*dx + dy*dy with separately rounded products and sum (no FMA contraction) and strict '<' so
// the lowest index wins ties: identical to numpy argmin / scipy KDTree.query on the same sites.
// ------------------------------------------------------------------------------------------------
constexpr int RASTER_POINTS_PER_LANE = 4;
constexpr int RASTER_WARPS = 8;

template <typename IdxT>
__global__ void __launch_bounds__(RASTER_WARPS * 32)
k_raster2d(RasterArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * CHAIN_TILE + lane;
    const bool in_range = c < a.C;
    bool active = in_range;
    int sslot = 0, islot = 0;
    if (in_range) {
        if (a.move != nullptr) {
            const int mv = a.move[c];
            active = !isinf(a.lpr[c]) && mv != 4 * a.n_spaces && (mv >> 2) == a.space &&
                     (mv & 3) != BBB_MOVE_VALUES;
        }
        if (a.sel != nullptr) sslot = a.sel[c] ^ a.slot_xor;
        islot = a.slot_xor;  // current states -> slot 0, proposals -> slot 1
    }
    if (!__syncthreads_or(active ? 1 : 0)) return;
    const int k = active ? a.k[(long long)sslot * a.C + c] : 0;
    int kmax = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    const double *sx = a.sites + ((long long)sslot * 2 + 0) * a.K * a.C + (in_range ? c : 0);
    const double *sy = a.sites + ((long long)sslot * 2 + 1) * a.K * a.C + (in_range ? c : 0);
    IdxT *out = reinterpret_cast<IdxT *>(a.idx) + (long long)islot * a.G * a.C + (in_range ? c : 0);

    const int g_begin = blockIdx.y * a.points_per_block;
    const int g_end = min(a.G, g_begin + a.points_per_block);
    for (int g0 = g_begin + warp * RASTER_POINTS_PER_LANE; g0 < g_end; g0 += RASTER_WARPS * RASTER_POINTS_PER_LANE) {
        double gx[RASTER_POINTS_PER_LANE], gy[RASTER_POINTS_PER_LANE], best[RASTER_POINTS_PER_LANE];
        int bi[RASTER_POINTS_PER_LANE];
#pragma unroll
        for (int q = 0; q < RASTER_POINTS_PER_LANE; ++q) {
            const int g = min(g0 + q, a.G - 1);
            gx[q] = a.px[g];
            gy[q] = a.py[g];
            best[q] = INFINITY;
            bi[q] = 0;
        }
        for (int j = 0; j < kmax; ++j) {
            if (j < k) {
                const double x = sx[(long long)j * a.C];
                const double y = sy[(long long)j * a.C];
#pragma unroll
                for (int q = 0; q < RASTER_POINTS_PER_LANE; ++q) {
                    const double dx = x - gx[q];
                    const double dy = y - gy[q];
                    const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    if (d < best[q]) { best[q] = d; bi[q] = j; }
                }
            }
        }
        if (active) {
#pragma unroll
            for (int q = 0; q < RASTER_POINTS_PER_LANE; ++q)
                if (g0 + q < g_end) out[(long long)(g0 + q) * a.C] = (IdxT)bi[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// F1, change-local: a proposal edits ONE cell (remove `rm` and/or insert at `ins`), so the nearest
// site of a grid point only has to be re-derived from its current owner `o`:
//   * o != rm: the owner survives (renumbered by the edit); if a cell was inserted, the new owner is
//     the closer of (owner, inserted cell), the lower index on an exact tie -- 2 distance evaluations;
//   * o == rm: the owner was removed or moved -> full scan of the proposed sites for this point.
// This is exactly argmin over the proposed sites with the lowest-index tie rule (removing a
// non-owner cannot change an argmin and the renumbering preserves index order), so the indices are
// bit-identical to the brute-force kernel; cost ~ (2 + fraction_rescanned * k) evaluations per point
// instead of k.
//
// Index slots: slot 0 holds the CURRENT index of every chain, slot 1 the PROPOSED index of every chain
// (so the ray kernel reads one slot for all chains); k_idx_commit_copy has just made slot 1 a copy of
// slot 0, this kernel overwrites the chains that propose a geometry change.  The sites of the block's
// 32 chains are staged in shared memory as [cell][lane] so the per-point owner lookup is a
// conflict-free LDS instead of a 32-sector gather.
// ------------------------------------------------------------------------------------------------
template <typename IdxT, bool STAGE>
__global__ void __launch_bounds__(RASTER_WARPS * 32)
k_raster2d_inc(RasterArgs a) {
    extern __shared__ double s_sites[];  // STAGE: [2][K_block][32]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * CHAIN_TILE + lane;
    const bool in_range = c < a.C;
    bool active = false;
    int cur = 0, rm = -1, ins = -1;
    double nx = 0.0, ny = 0.0;
    if (in_range) {
        const int mv = a.move[c];
        active = !isinf(a.lpr[c]) && mv != 4 * a.n_spaces && (mv >> 2) == a.space &&
                 (mv & 3) != BBB_MOVE_VALUES;
        cur = a.sel[c];
        if (active) {
            rm = a.edit[c];
            ins = a.edit[a.C + c];
            nx = a.edit_site[c];
            ny = a.edit_site[a.C + c];
        }
    }
    if (!__syncthreads_or(active ? 1 : 0)) return;  // k_idx_commit_copy already wrote this tile's proposed slot
    const long long cc = in_range ? c : 0;
    const int k_new = active ? a.k[(long long)(cur ^ 1) * a.C + c] : 0;
    const double *sx = a.sites + ((long long)(cur ^ 1) * 2 + 0) * a.K * a.C + cc;
    const double *sy = a.sites + ((long long)(cur ^ 1) * 2 + 1) * a.K * a.C + cc;
    int kmax_blk = k_new;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax_blk = max(kmax_blk, __shfl_xor_sync(0xffffffffu, kmax_blk, o));
    double *tx = s_sites + lane;
    double *ty = s_sites + (long long)a.k_rows * 32 + lane;
    if (STAGE) {
        for (int j = warp; j < kmax_blk; j += RASTER_WARPS) {
            tx[j * 32] = (j < k_new) ? sx[(long long)j * a.C] : 0.0;
            ty[j * 32] = (j < k_new) ? sy[(long long)j * a.C] : 0.0;
        }
        __syncthreads();
    }
    IdxT *slot0 = reinterpret_cast<IdxT *>(a.idx) + cc;
    IdxT *slot1 = slot0 + (long long)a.G * a.C;

    constexpr int PPL = RASTER_POINTS_PER_LANE;
    const int g_begin = blockIdx.y * a.points_per_block;
    const int g_end = min(a.G, g_begin + a.points_per_block);
    for (int g0 = g_begin + warp * PPL; g0 < g_end; g0 += RASTER_WARPS * PPL) {
        double gx[PPL], gy[PPL];
        int m[PPL];
        bool scan[PPL];
        bool any_scan = false;
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int g = min(g0 + q, a.G - 1);
            gx[q] = a.px[g];
            gy[q] = a.py[g];
            const int o = active ? (int)slot0[(long long)g * a.C] : 0;
            scan[q] = active && rm >= 0 && o == rm;
            int mm = o - ((rm >= 0 && o > rm) ? 1 : 0);
            mm += (ins >= 0 && mm >= ins) ? 1 : 0;
            m[q] = mm;
            any_scan |= scan[q];
        }
        if (active && ins >= 0) {
            double ox[PPL], oy[PPL];
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int mm = scan[q] ? 0 : m[q];
                ox[q] = STAGE ? tx[mm * 32] : sx[(long long)mm * a.C];
                oy[q] = STAGE ? ty[mm * 32] : sy[(long long)mm * a.C];
            }
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const double dx = ox[q] - gx[q], dy = oy[q] - gy[q];
                const double d_own = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const double ex = nx - gx[q], ey = ny - gy[q];
                const double d_new = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                if (!scan[q] && (d_new < d_own || (d_new == d_own && ins < m[q]))) m[q] = ins;
            }
        }
        // points whose owner was removed/moved need a full scan of the proposed sites.  Only ~1/k of a proposing
        // chain's points do, so scanning here would idle 99 % of the lanes: queue them per chain instead and let
        // k_raster2d_rescan do the scans with lanes = sites (warp-shuffle argmin).
        (void)any_scan;
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            if (scan[q] && g0 + q < g_end) {
                const int pos = atomicAdd(a.rescan_count + c, 1);
                a.rescan_list[(long long)c * a.G + pos] = (unsigned)(g0 + q);
            }
        }
        if (active) {
#pragma unroll
            for (int q = 0; q < PPL; ++q)
                if (g0 + q < g_end && !scan[q]) slot1[(long long)(g0 + q) * a.C] = (IdxT)m[q];
        }
    }
}

// Second phase of the change-local update: one block per chain drains the chain's queue of grid points that
// lost their owner.  The chain's proposed sites are staged in shared memory; one warp per point evaluates the
// sites with lanes = sites and reduces (distance, index) lexicographically with shuffles, so the lowest index
// wins ties exactly as in the brute-force kernel.
constexpr int RESCAN_WARPS = 8;
constexpr int RESCAN_SPLIT = 4;  // blocks per chain (cell sizes vary a lot: keeps the largest cell off the critical path)

template <typename IdxT>
__global__ void __launch_bounds__(RESCAN_WARPS * 32)
k_raster2d_rescan(RasterArgs a) {
    extern __shared__ double s_xy[];  // [2][k_new]
    const long long c = blockIdx.x;
    const int n = a.rescan_count[c];
    if (n == 0) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int prop = a.sel[c] ^ 1;
    const int k_new = a.k[(long long)prop * a.C + c];
    const double *sx = a.sites + ((long long)prop * 2 + 0) * a.K * a.C + c;
    const double *sy = a.sites + ((long long)prop * 2 + 1) * a.K * a.C + c;
    for (int j = threadIdx.x; j < k_new; j += blockDim.x) {
        s_xy[j] = sx[(long long)j * a.C];
        s_xy[a.k_rows + j] = sy[(long long)j * a.C];
    }
    __syncthreads();
    IdxT *slot1 = reinterpret_cast<IdxT *>(a.idx) + (long long)a.G * a.C + c;
    const unsigned *list = a.rescan_list + c * a.G;
    for (int item = blockIdx.y * RESCAN_WARPS + warp; item < n; item += RESCAN_WARPS * RESCAN_SPLIT) {
        const unsigned g = list[item];
        const double gx = a.px[g], gy = a.py[g];
        double best = INFINITY;
        int bi = 0x7fffffff;
        for (int j = lane; j < k_new; j += 32) {
            const double dx = s_xy[j] - gx, dy = s_xy[a.k_rows + j] - gy;
            const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d < best) { best = d; bi = j; }
        }
        // lexicographic (distance, index) minimum over the warp with three REDUX steps: non-negative doubles
        // order like their bit patterns, so min the high word, then the low word among the lanes that hold the
        // minimal high word, then the index among the lanes that hold the minimal distance
        const unsigned hi = (unsigned)__double2hiint(best), lo = (unsigned)__double2loint(best);
        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
        bi = (int)__reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned)bi : 0xffffffffu);
        if (lane == 0) slot1[(long long)g * a.C] = (IdxT)bi;
    }
}

// Index-slot maintenance, one streaming pass per step over [G][C]: chains whose accepted geometry is still
// pending in slot 1 (idx_sel == 1) are committed to slot 0, then slot 1 := slot 0 for EVERY chain, so that the
// proposed slot is complete (the change-local kernel overwrites the chains with a geometry proposal).
// 16 bytes (16 uint8 / 8 uint16 chains) per thread.
template <typename IdxT>
__global__ void k_idx_commit_copy(IdxT *idx, const uint8_t *idx_sel, int G, long long C) {
    constexpr int V = 16 / (int)sizeof(IdxT);
    const long long vec_per_row = C / V;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= vec_per_row * G) return;
    const long long g = i / vec_per_row, cv = (i % vec_per_row) * V;
    uint4 *p0 = reinterpret_cast<uint4 *>(idx + g * C + cv);
    uint4 *p1 = reinterpret_cast<uint4 *>(idx + ((long long)G + g) * C + cv);
    uint4 w = *p0;
    unsigned pend = 0;
    {
        uint8_t flags[V];
        if (V == 16) *reinterpret_cast<uint4 *>(flags) = *reinterpret_cast<const uint4 *>(idx_sel + cv);
        else *reinterpret_cast<uint2 *>(flags) = *reinterpret_cast<const uint2 *>(idx_sel + cv);
#pragma unroll
        for (int v = 0; v < V; ++v) pend |= (flags[v] ? 1u : 0u) << v;
    }
    if (pend) {
        const uint4 n = *p1;
        IdxT *wb = reinterpret_cast<IdxT *>(&w);
        const IdxT *nb = reinterpret_cast<const IdxT *>(&n);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if ((pend >> v) & 1u) wb[v] = nb[v];
        *p0 = w;
    }
    *p1 = w;
}
template <typename IdxT>
__global__ void k_idx_commit_copy_scalar(IdxT *idx, const uint8_t *idx_sel, int G, long long C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * G) return;
    const long long c = i % C;
    IdxT v = idx[i];
    if (idx_sel[c]) {
        v = idx[(long long)G * C + i];
        idx[i] = v;
    }
    idx[(long long)G * C + i] = v;
}

int launch_idx_commit_copy(void *idx, int idx_bytes, const uint8_t *idx_sel, int G, long long C, cudaStream_t stream) {
    const int V = 16 / idx_bytes;
    if (C % V == 0) {
        const long long n = C / V * G;
        const int blocks = (int)((n + 255) / 256);
        if (idx_bytes == 1) k_idx_commit_copy<uint8_t><<<blocks, 256, 0, stream>>>((uint8_t *)idx, idx_sel, G, C);
        else k_idx_commit_copy<uint16_t><<<blocks, 256, 0, stream>>>((uint16_t *)idx, idx_sel, G, C);
    } else {
        const long long n = C * G;
        const int blocks = (int)((n + 255) / 256);
        if (idx_bytes == 1) k_idx_commit_copy_scalar<uint8_t><<<blocks, 256, 0, stream>>>((uint8_t *)idx, idx_sel, G, C);
        else k_idx_commit_copy_scalar<uint16_t><<<blocks, 256, 0, stream>>>((uint16_t *)idx, idx_sel, G, C);
    }
    return (int)cudaGetLastError();
}

template <typename IdxT>
static int launch_inc_variant(const RasterArgs &a, dim3 grid, cudaStream_t stream) {
    const size_t smem = (size_t)a.k_rows * 32 * 2 * sizeof(double);  // rows = upper bound of k over the chains
    if (smem <= 100 * 1024) {  // at least two resident blocks per SM
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(k_raster2d_inc<IdxT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
        }
        k_raster2d_inc<IdxT, true><<<grid, RASTER_WARPS * 32, smem, stream>>>(a);
    } else {
        k_raster2d_inc<IdxT, false><<<grid, RASTER_WARPS * 32, 0, stream>>>(a);
    }
    return (int)cudaGetLastError();
}

int launch_raster2d_inc(const RasterArgs &a, cudaStream_t stream) {
    const int chain_tiles = (int)((a.C + CHAIN_TILE - 1) / CHAIN_TILE);
    dim3 grid(chain_tiles, (a.G + a.points_per_block - 1) / a.points_per_block);
    int rc = a.idx_bytes == 1 ? launch_inc_variant<uint8_t>(a, grid, stream) : launch_inc_variant<uint16_t>(a, grid, stream);
    if (rc != 0) return rc;
    const size_t smem = (size_t)a.k_rows * 2 * sizeof(double);
    dim3 rgrid((unsigned)a.C, RESCAN_SPLIT);
    if (a.idx_bytes == 1) k_raster2d_rescan<uint8_t><<<rgrid, RESCAN_WARPS * 32, smem, stream>>>(a);
    else k_raster2d_rescan<uint16_t><<<rgrid, RESCAN_WARPS * 32, smem, stream>>>(a);
    rc = (int)cudaGetLastError();
    if (rc != 0) return rc;
    return (int)cudaMemsetAsync(a.rescan_count, 0, (size_t)a.C * sizeof(int32_t), stream);  // queues drained
}

int launch_raster2d(const RasterArgs &a, cudaStream_t stream) {
    const int chain_tiles = (int)((a.C + CHAIN_TILE - 1) / CHAIN_TILE);
    // enough blocks for >= ~4 waves of 148 SMs, each block at least 64 points
    int ppb = a.points_per_block;
    dim3 grid(chain_tiles, (a.G + ppb - 1) / ppb);
    if (a.idx_bytes == 1) k_raster2d<uint8_t><<<grid, RASTER_WARPS * 32, 0, stream>>>(a);
    else k_raster2d<uint16_t><<<grid, RASTER_WARPS * 32, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

int raster_points_per_block(int G, long long C) {
    const long long chain_tiles = (C + CHAIN_TILE - 1) / CHAIN_TILE;
    const long long target_blocks = 148LL * 8;
    long long per_tile = (target_blocks + chain_tiles - 1) / chain_tiles;
    if (per_tile < 1) per_tile = 1;
    long long ppb = (G + per_tile - 1) / per_tile;
    const int quantum = RASTER_WARPS * RASTER_POINTS_PER_LANE;
    ppb = ((ppb + quantum - 1) / quantum) * quantum;
    if (ppb < quantum) ppb = quantum;
    return (int)ppb;
}

// ------------------------------------------------------------------------------------------------
// F2: d_pred[r][c] = sum_j data[j] * f(value[cell_idx[indices[j]][c]][c]);  phi[c] = sum_r w_r res^2.
//
// Persistent blocks: grid = chain tiles x ray groups, sized to one block per SM.  A block owns 32*CPL
// chains (each lane CPL consecutive chains) and stages their (reciprocal) cell values ONCE in shared
// memory as CPL planes table_q[cell][lane] (256 B rows, bank-conflict free), then its 16 warps walk
// the block's rays.  Per CSR non-zero a warp issues
//   * 1 + 2 shuffles (column pre-multiplied by C, fp64 weight) -- the CSR entries were loaded 32 at a
//     time, coalesced, and the broadcast is amortised over 32*CPL chains,
//   * 2 coalesced 128 B gathers: the 32-bit word holding the nearest-site indices of the lane's CPL
//     chains (4 x uint8 or 2 x uint16) in the current and in the proposed index slot, merged per chain
//     with one byte permute,
//   * per chain: one byte permute that BUILDS the shared-memory offset (cell * 256 + lane * 8), one
//     LDS with the plane as immediate offset, one DFMA.
// Eight entries are in flight per warp to overlap the L2 round trips of the gathers.
// The block's misfit goes to partial[group][c] and is summed in group order by the accept stage:
// bit-reproducible run to run (no atomics).  d_pred is only written on request.
// ------------------------------------------------------------------------------------------------
constexpr int SPMM_WARPS = 16;
constexpr int SPMM_BATCH = 8;   // CSR entries whose index gathers are in flight together per warp
constexpr int SPMM_CHUNK = SPMM_WARPS;  // rays handed out per scheduling chunk (one per warp)

template <int CPL> struct SpmmPlane { static constexpr int ROWS = (CPL == 4) ? 208 : (CPL == 2 ? 416 : 832); };
constexpr int SPMM_STAGE_BYTES = SPMM_WARPS * 32 * 16;  // per-warp staging of 32 CSR entries {weight, column}

template <typename IdxT, int CPL, bool MERGE>
__global__ void __launch_bounds__(SPMM_WARPS * 32, 1)
k_ray_spmm(SpmmArgs a) {
    extern __shared__ __align__(256) unsigned char smem_raw[];  // CPL planes of ROWS x 32 doubles, then CSR staging
    constexpr int PLANE_BYTES = SpmmPlane<CPL>::ROWS * 256;
    uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + CPL * PLANE_BYTES) + (threadIdx.x >> 5) * 32;
    constexpr int IB = (int)sizeof(IdxT);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c0 = (long long)blockIdx.x * (32 * CPL) + (long long)lane * CPL;
    bool active[CPL];
    int vslot[CPL], islot[CPL], k[CPL];
    bool any_active = false;
    int kmax = 0;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const long long c = c0 + q;
        active[q] = c < a.C;
        vslot[q] = 0;
        islot[q] = 0;
        if (c < a.C) {
            bool geom = false;
            if (a.move != nullptr) {
                const int mv = a.move[c];
                active[q] = !isinf(a.lpr[c]) && mv != 4 * a.n_spaces && (mv >> 2) == a.space;
                geom = (mv & 3) != BBB_MOVE_VALUES;
            }
            if (a.sel != nullptr) vslot[q] = a.sel[c] ^ a.slot_xor;
            // proposals: every chain's proposed index is in slot 1; current states: slot idx_sel
            if (a.move != nullptr) islot[q] = 1;
            else if (a.idx_sel != nullptr) islot[q] = a.idx_sel[c];
            (void)geom;
        }
        k[q] = active[q] ? a.k[(long long)vslot[q] * a.C + c] : 0;
        any_active |= active[q];
        kmax = max(kmax, k[q]);
    }
    if (!__syncthreads_or(any_active ? 1 : 0)) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    // stage the (transformed) cell values of this block's chains: plane q, row = cell, column = lane
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const long long c = min(c0 + q, a.C - 1);
        const double *vals = a.values + ((long long)vslot[q] * a.n_params + a.param) * a.K * a.C + c;
        double *plane = reinterpret_cast<double *>(smem_raw + q * PLANE_BYTES);
        for (int j = warp; j < kmax; j += SPMM_WARPS) {
            double v = (j < k[q]) ? vals[(long long)j * a.C] : 1.0;
            if (a.transform == BBB_TRANSFORM_RECIPROCAL) v = 1.0 / v;
            plane[j * 32 + lane] = v;
        }
    }
    __syncthreads();

    // 32-bit word view of the index slots; the lane's word of grid point `col` is word_base[col * words_per_point]
    const bool lane_ok = c0 + CPL - 1 < a.C;
    const unsigned words_per_point = (unsigned)(a.C * IB / 4);
    // uniform bases + 32-bit per-thread word offsets keep the address arithmetic of the gathers short
    const unsigned *w_slot0 = reinterpret_cast<const unsigned *>(a.idx);
    const unsigned *w_slot1 = w_slot0 + (size_t)a.G * words_per_point;
    const IdxT *e_slot = reinterpret_cast<const IdxT *>(a.idx) + (size_t)islot[0] * a.G * a.C;  // CPL == 1
    // offsets in 32-bit words (CPL * IB == 4) or in elements (CPL == 1)
    const unsigned point_stride = (CPL * IB == 4) ? words_per_point : (unsigned)a.C;
    const unsigned lane_off = lane_ok ? (unsigned)((CPL * IB == 4) ? c0 * IB / 4 : c0) : 0u;
    unsigned merge_sel = 0;  // per chain: take its bytes from slot 0 (nibble b) or slot 1 (nibble b + 4)
    unsigned off_sel[CPL];   // builds (cell << 8) | (lane * 8) from the merged word and lane8
    const unsigned lane8 = (unsigned)lane * 8u;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
#pragma unroll
        for (int b = 0; b < IB; ++b) {
            const int byte = q * IB + b;
            merge_sel |= (unsigned)(byte + 4 * islot[q]) << (4 * byte);
        }
        // result byte0 = lane8.byte0 (index 4), byte1 = cell low byte, byte2 = cell high byte (u16) or 0, byte3 = 0
        off_sel[q] = 0x5004u | ((unsigned)(q * IB) << 4) | ((IB == 2 ? (unsigned)(q * IB + 1) : 5u) << 8);
    }

    double ssq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) ssq[q] = 0.0;
    const int n_chunks = (a.R + SPMM_CHUNK - 1) / SPMM_CHUNK;
    for (int chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
        const int r = chunk * SPMM_CHUNK + warp;
        if (r >= a.R) continue;
        const int j0 = a.indptr[r];
        const int j1 = a.indptr[r + 1];
        double acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = 0.0;
        for (int jb = j0; jb < j1; jb += 32) {
            // 32 CSR entries: one coalesced load per lane, parked in the warp's staging row; each entry is then
            // fetched with ONE broadcast LDS.128 (cheaper on the shared-memory pipe than three shuffles)
            const int jj = jb + lane;
            const unsigned my_col = (jj < j1) ? (unsigned)a.indices[jj] * point_stride : 0u;
            const double my_val = (jj < j1) ? a.data[jj] : 0.0;
            __syncwarp();
            stage[lane] = make_uint4((unsigned)__double2loint(my_val), (unsigned)__double2hiint(my_val), my_col, 0u);
            __syncwarp();
            const int n = min(32, j1 - jb);
            for (int u0 = 0; u0 < n; u0 += SPMM_BATCH) {
                unsigned w[SPMM_BATCH];
                double v[SPMM_BATCH];
#pragma unroll
                for (int u = 0; u < SPMM_BATCH; ++u) {
                    const uint4 e = stage[u0 + u];
                    v[u] = __hiloint2double((int)e.y, (int)e.x);
                    const unsigned col = e.z + lane_off;
                    if (CPL * IB == 4) {
                        if (MERGE) w[u] = __byte_perm(w_slot0[col], w_slot1[col], merge_sel);
                        else w[u] = w_slot1[col];
                    } else {  // CPL == 1: one index per lane, read from the chain's own slot
                        w[u] = (unsigned)e_slot[col];
                    }
                }
#pragma unroll
                for (int u = 0; u < SPMM_BATCH; ++u) {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const unsigned off = __byte_perm(w[u], lane8, off_sel[q]);
                        const double t = *reinterpret_cast<const double *>(smem_raw + off + q * PLANE_BYTES);
                        acc[q] = fma(v[u], t, acc[q]);
                    }
                }
            }
        }
        const double dob = a.dobs[r];
        const double wr = (a.w != nullptr) ? a.w[r] : 1.0;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            if (active[q]) {
                if (a.dpred != nullptr) a.dpred[(long long)r * a.C + c0 + q] = acc[q];
                const double res = acc[q] - dob;
                ssq[q] += wr * (res * res);
            }
        }
    }
    __syncthreads();  // the table is dead: reuse plane 0 as red[warp][q][lane]
    double *red = reinterpret_cast<double *>(smem_raw);
#pragma unroll
    for (int q = 0; q < CPL; ++q) red[(warp * CPL + q) * 32 + lane] = ssq[q];
    __syncthreads();
    if (warp == 0 && a.partial != nullptr) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            if (!active[q]) continue;
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < SPMM_WARPS; ++w) tot += red[(w * CPL + q) * 32 + lane];
            a.partial[(long long)blockIdx.y * a.C + c0 + q] = tot;
        }
    }
}

// chains per lane the shared-memory value table allows for this K (0: K too large)
static int spmm_cpl(int K, long long C, int idx_bytes) {
    if (idx_bytes == 1) {
        if (C % 4 == 0 && K <= SpmmPlane<4>::ROWS) return 4;
    } else {
        if (C % 2 == 0 && K <= SpmmPlane<2>::ROWS) return 2;
    }
    return K <= SpmmPlane<1>::ROWS ? 1 : 0;
}

int spmm_ray_groups(int R, long long C, int K, int idx_bytes, int n_sm) {
    const int cpl = spmm_cpl(K, C, idx_bytes);
    if (cpl == 0) return 0;
    const long long chain_tiles = (C + 32 * cpl - 1) / (32 * cpl);
    long long groups = n_sm / chain_tiles;  // one persistent block per SM
    const long long max_groups = (R + RAY_TILE - 1) / RAY_TILE;  // rows of the partial buffer
    if (groups > max_groups) groups = max_groups;
    if (groups < 1) groups = 1;
    return (int)groups;
}

template <typename IdxT, int CPL, bool MERGE>
static int launch_spmm_merge(const SpmmArgs &a, cudaStream_t stream) {
    const int chain_tiles = (int)((a.C + 32 * CPL - 1) / (32 * CPL));
    dim3 grid(chain_tiles, a.n_groups);
    const size_t smem = (size_t)CPL * SpmmPlane<CPL>::ROWS * 256 + SPMM_STAGE_BYTES;
    cudaError_t err = cudaFuncSetAttribute(k_ray_spmm<IdxT, CPL, MERGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    k_ray_spmm<IdxT, CPL, MERGE><<<grid, SPMM_WARPS * 32, smem, stream>>>(a);
    return (int)cudaGetLastError();
}
template <typename IdxT, int CPL>
static int launch_spmm_variant(const SpmmArgs &a, cudaStream_t stream) {
    // proposals read slot 1 only; current states may sit in either slot (merge the two words)
    return (a.move != nullptr) ? launch_spmm_merge<IdxT, CPL, false>(a, stream) : launch_spmm_merge<IdxT, CPL, true>(a, stream);
}

int launch_ray_spmm(const SpmmArgs &a, cudaStream_t stream) {
    if ((long long)a.G * a.C * a.idx_bytes * 2 / 4 >= (1LL << 32)) return (int)cudaErrorInvalidConfiguration;
    const int cpl = spmm_cpl(a.k_rows, a.C, a.idx_bytes);
    if (a.idx_bytes == 1) {
        if (cpl == 4) return launch_spmm_variant<uint8_t, 4>(a, stream);
        if (cpl == 1) return launch_spmm_variant<uint8_t, 1>(a, stream);
    } else {
        if (cpl == 2) return launch_spmm_variant<uint16_t, 2>(a, stream);
        if (cpl == 1) return launch_spmm_variant<uint16_t, 1>(a, stream);
    }
    return (int)cudaErrorInvalidConfiguration;  // n_dimensions_max too large for the shared-memory value table
}

}  // namespace bbb
