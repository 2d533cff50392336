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
        if (a.idx_sel != nullptr) islot = a.idx_sel[c] ^ a.slot_xor;
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
// instead of k.  Reads the current index slot, writes the proposed one (every point: no copy needed
// on accept, the slot bit is flipped).
// ------------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(RASTER_WARPS * 32)
k_raster2d_inc(RasterArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * CHAIN_TILE + lane;
    const bool in_range = c < a.C;
    bool active = false;
    int cur = 0, icur = 0, rm = -1, ins = -1;
    double nx = 0.0, ny = 0.0;
    if (in_range) {
        const int mv = a.move[c];
        active = !isinf(a.lpr[c]) && mv != 4 * a.n_spaces && (mv >> 2) == a.space && (mv & 3) != BBB_MOVE_VALUES;
        cur = a.sel[c];
        icur = a.idx_sel[c];
        if (active) {
            rm = a.edit[c];
            ins = a.edit[a.C + c];
            nx = a.edit_site[c];
            ny = a.edit_site[a.C + c];
        }
    }
    if (!__syncthreads_or(active ? 1 : 0)) return;
    const long long cc = in_range ? c : 0;
    const int k_new = active ? a.k[(long long)(cur ^ 1) * a.C + c] : 0;
    const double *sx = a.sites + ((long long)(cur ^ 1) * 2 + 0) * a.K * a.C + cc;
    const double *sy = a.sites + ((long long)(cur ^ 1) * 2 + 1) * a.K * a.C + cc;
    const IdxT *in = reinterpret_cast<const IdxT *>(a.idx) + (long long)icur * a.G * a.C + cc;
    IdxT *out = reinterpret_cast<IdxT *>(a.idx) + (long long)(icur ^ 1) * a.G * a.C + cc;

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
            const int o = active ? (int)in[(long long)g * a.C] : 0;
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
                ox[q] = sx[(long long)mm * a.C];
                oy[q] = sy[(long long)mm * a.C];
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
        if (__any_sync(0xffffffffu, any_scan)) {
            const int kk = any_scan ? k_new : 0;
            int kmax = kk;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            double best[PPL];
            int bi[PPL];
#pragma unroll
            for (int q = 0; q < PPL; ++q) { best[q] = INFINITY; bi[q] = 0; }
            for (int j = 0; j < kmax; ++j) {
                if (j < kk) {
                    const double x = sx[(long long)j * a.C];
                    const double y = sy[(long long)j * a.C];
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        const double dx = x - gx[q], dy = y - gy[q];
                        const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        if (d < best[q]) { best[q] = d; bi[q] = j; }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < PPL; ++q)
                if (scan[q]) m[q] = bi[q];
        }
        if (active) {
#pragma unroll
            for (int q = 0; q < PPL; ++q)
                if (g0 + q < g_end) out[(long long)(g0 + q) * a.C] = (IdxT)m[q];
        }
    }
}

int launch_raster2d_inc(const RasterArgs &a, cudaStream_t stream) {
    const int chain_tiles = (int)((a.C + CHAIN_TILE - 1) / CHAIN_TILE);
    dim3 grid(chain_tiles, (a.G + a.points_per_block - 1) / a.points_per_block);
    if (a.idx_bytes == 1) k_raster2d_inc<uint8_t><<<grid, RASTER_WARPS * 32, 0, stream>>>(a);
    else k_raster2d_inc<uint16_t><<<grid, RASTER_WARPS * 32, 0, stream>>>(a);
    return (int)cudaGetLastError();
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
// Block = 32*CPL chains x RAY_TILE rays.  Each lane owns CPL consecutive chains, so the nearest-site
// indices of its chains at one grid point are ONE 32-bit word (4 x uint8 or 2 x uint16): per CSR
// non-zero a warp issues one coalesced 128 B gather per index slot (the current and the proposed
// slot, merged per chain with a byte permute), CPL conflict-free LDS from the per-block table of
// (reciprocal) cell values table[cell][q][lane], and CPL DFMAs.  The CSR entries are loaded 32 at a
// time (coalesced) and broadcast by shuffles, which is amortised over 32*CPL chains; eight entries'
// gathers are issued back to back to overlap their L2 round trips.
// The per-block misfit goes to partial[tile][c] and is summed in tile order by the accept stage:
// bit-reproducible run to run (no atomics).  d_pred is only written on request.
// ------------------------------------------------------------------------------------------------
constexpr int SPMM_WARPS = 16;

template <typename IdxT, int CPL>
__global__ void __launch_bounds__(SPMM_WARPS * 32)
k_ray_spmm(SpmmArgs a) {
    extern __shared__ double table[];  // [K_block][CPL][32], reused as the cross-warp reduction scratch
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
            if (a.idx_sel != nullptr) islot[q] = a.idx_sel[c] ^ ((a.move != nullptr && geom) ? 1 : 0);
        }
        k[q] = active[q] ? a.k[(long long)vslot[q] * a.C + c] : 0;
        any_active |= active[q];
        kmax = max(kmax, k[q]);
    }
    if (!__syncthreads_or(any_active ? 1 : 0)) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    // stage the (transformed) cell values of this block's chains
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const long long c = min(c0 + q, a.C - 1);
        const double *vals = a.values + ((long long)vslot[q] * a.n_params + a.param) * a.K * a.C + c;
        for (int j = warp; j < kmax; j += SPMM_WARPS) {
            double v = (j < k[q]) ? vals[(long long)j * a.C] : 1.0;
            if (a.transform == BBB_TRANSFORM_RECIPROCAL) v = 1.0 / v;
            table[(j * CPL + q) * 32 + lane] = v;
        }
    }
    __syncthreads();

    const long long slot_stride = (long long)a.G * a.C;
    const bool lane_ok = c0 + CPL - 1 < a.C;  // C is a multiple of CPL (checked by the launcher)
    const IdxT *idx0 = reinterpret_cast<const IdxT *>(a.idx) + (lane_ok ? c0 : 0);
    unsigned selector = 0;  // byte-permute selector merging the two slots' words per chain
    if (CPL * sizeof(IdxT) == 4) {
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
            for (int b = 0; b < (int)sizeof(IdxT); ++b) {
                const int byte = q * (int)sizeof(IdxT) + b;
                selector |= (unsigned)(byte + 4 * islot[q]) << (4 * byte);
            }
    }
    const double *tab = table + lane;

    const int r_begin = blockIdx.y * RAY_TILE;
    const int r_end = min(a.R, r_begin + RAY_TILE);
    double ssq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) ssq[q] = 0.0;
    for (int r = r_begin + warp; r < r_end; r += SPMM_WARPS) {
        const int j0 = a.indptr[r];
        const int j1 = a.indptr[r + 1];
        double acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[q] = 0.0;
        for (int jb = j0; jb < j1; jb += 32) {
            const int jj = jb + lane;
            const int my_col = (jj < j1) ? a.indices[jj] : 0;
            const double my_val = (jj < j1) ? a.data[jj] : 0.0;
            const int n = min(32, j1 - jb);
            for (int u0 = 0; u0 < n; u0 += 8) {
                unsigned w[8];
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long long col = __shfl_sync(0xffffffffu, my_col, u0 + u);
                    v[u] = __shfl_sync(0xffffffffu, my_val, u0 + u);
                    const IdxT *p = idx0 + col * a.C;
                    if (CPL * sizeof(IdxT) == 4) {
                        const unsigned w0 = lane_ok ? *reinterpret_cast<const unsigned *>(p) : 0u;
                        const unsigned w1 = lane_ok ? *reinterpret_cast<const unsigned *>(p + slot_stride) : 0u;
                        w[u] = __byte_perm(w0, w1, selector);
                    } else {
                        w[u] = lane_ok ? (unsigned)p[(long long)islot[0] * slot_stride] : 0u;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const unsigned cell = (CPL == 1) ? w[u] : ((w[u] >> (8 * (int)sizeof(IdxT) * q)) & ((1u << (8 * sizeof(IdxT))) - 1u));
                        acc[q] += v[u] * tab[(cell * CPL + q) * 32];
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
    __syncthreads();  // the table is dead: reuse it as red[warp][q][lane]
#pragma unroll
    for (int q = 0; q < CPL; ++q) table[(warp * CPL + q) * 32 + lane] = ssq[q];
    __syncthreads();
    if (warp == 0 && a.partial != nullptr) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            if (!active[q]) continue;
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < SPMM_WARPS; ++w) tot += table[(w * CPL + q) * 32 + lane];
            a.partial[(long long)blockIdx.y * a.C + c0 + q] = tot;
        }
    }
}

template <typename IdxT, int CPL>
static int launch_spmm_variant(const SpmmArgs &a, cudaStream_t stream) {
    const int chain_tiles = (int)((a.C + 32 * CPL - 1) / (32 * CPL));
    const int ray_tiles = (a.R + RAY_TILE - 1) / RAY_TILE;
    dim3 grid(chain_tiles, ray_tiles);
    size_t smem = (size_t)a.K * CPL * 32 * sizeof(double);
    const size_t red = (size_t)SPMM_WARPS * CPL * 32 * sizeof(double);
    if (smem < red) smem = red;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(k_ray_spmm<IdxT, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
    }
    k_ray_spmm<IdxT, CPL><<<grid, SPMM_WARPS * 32, smem, stream>>>(a);
    return (int)cudaGetLastError();
}

int launch_ray_spmm(const SpmmArgs &a, cudaStream_t stream) {
    const size_t budget = 200 * 1024;  // dynamic shared memory available to the value table
    const size_t row = 32 * sizeof(double);
    if (a.idx_bytes == 1) {
        if (a.C % 4 == 0 && (size_t)a.K * 4 * row <= budget) return launch_spmm_variant<uint8_t, 4>(a, stream);
        if ((size_t)a.K * row <= 227 * 1024 - 1024) return launch_spmm_variant<uint8_t, 1>(a, stream);
    } else {
        if (a.C % 2 == 0 && (size_t)a.K * 2 * row <= budget) return launch_spmm_variant<uint16_t, 2>(a, stream);
        if ((size_t)a.K * row <= 227 * 1024 - 1024) return launch_spmm_variant<uint16_t, 1>(a, stream);
    }
    return (int)cudaErrorInvalidConfiguration;  // n_dimensions_max too large for the shared-memory value table
}

}  // namespace bbb
