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
// Resident per-chain caches (chain-major, unlike the [..][C] arrays of the other kernels, because here one block
// walks one chain): idx_cm [C][G'] nearest-site index, res [C][R] residual d_pred - d_obs, phi = sum w res^2.
// Rounding cannot accumulate without bound: the engine re-evaluates res and phi from scratch with the full SpMM
// every `resync_every` steps (abi.cu).
#include "finish.cuh"
#include "kernels.h"

namespace bbb {

constexpr int CS_THREADS = 512;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_CHUNK = CS_THREADS;       // change-list entries prepared per round
constexpr unsigned CS_NONE = 0xFFFFFFFFu;  // list entry: point re-derived, nearest site unchanged
constexpr int CS_M_BITS = 12;              // list entry = (grid point << 12) | new nearest site
constexpr int CS_SMEM_BLOCK = 112 * 1024;  // two resident blocks per SM

struct PrepEntry {
    int j0, len;  // entries of the point's column inside the current ray tile
    double dlt;   // change of the cell function at the point
};

__host__ __device__ inline size_t chain_step_smem(int tile_rays, int K) {
    return (size_t)tile_rays * 8 + (size_t)K * 4 * 8 + (size_t)CS_CHUNK * sizeof(PrepEntry);
}

long long chain_local_tile_rays(long long n_data, int kmax, long long n_points) {
    if (kmax > (1 << CS_M_BITS) || n_points > (1LL << (32 - CS_M_BITS)) - 1 || n_data < 1) return 0;
    const long long fixed = (long long)chain_step_smem(0, kmax) + 1024;
    long long max_tile = (CS_SMEM_BLOCK - fixed) / 8;
    if (max_tile > 65535) max_tile = 65535;  // csc_rows is uint16
    if (max_tile < 1024) return 0;
    const long long n_tiles = (n_data + max_tile - 1) / max_tile;
    return (n_data + n_tiles - 1) / n_tiles;
}

__device__ __forceinline__ double block_sum_to_thread0(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < CS_WARPS; ++w) tot += s_red[w];
    return tot;
}

template <typename IdxT>
__global__ void __launch_bounds__(CS_THREADS, 2) k_chain_step(const EngineParams *epp, int t) {
    const EngineParams &ep = *epp;
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    const bbb_target_spec &T = P.targets[t];
    const int s = T.space;
    const bbb_space_spec &S = P.spaces[s];
    const long long C = B.n_chains;
    const long long c = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ int s_front, s_back, s_accept;
    __shared__ double s_red[CS_WARPS];

    const int move = B.move[c];
    const double lpr = B.log_prob_ratio[c];
    if (!target_touched(P, move, lpr, t)) {  // noise move, other space, or rejected by the prior: no forward
        if (tid == 0) {
            Rng rng;
            rng_begin(rng, ep, c, STREAM_STEP, true);
            finish_chain(ep, c, rng);
            rng_end(rng, ep, c);
        }
        return;
    }

    extern __shared__ __align__(16) unsigned char cs_smem[];
    const int K = S.kmax, RT = T.csc_tile_rays, NT = T.csc_n_tiles, G = T.n_points, R = T.n_data;
    long long *acc = reinterpret_cast<long long *>(cs_smem);
    double *sx = reinterpret_cast<double *>(cs_smem + (size_t)RT * 8);
    double *sy = sx + K, *fo = sy + K, *fn = fo + K;
    PrepEntry *prep = reinterpret_cast<PrepEntry *>(fn + K);

    const int cur = B.sel[s][c], prop = cur ^ 1;
    const int k_old = k_ref(B, s, cur, c), k_new = k_ref(B, s, prop, c);
    const int inner = move & 3;
    const int rm = B.edit[c], ins = B.edit[C + c];
    const double nx = B.edit_site[c], ny = B.edit_site[C + c];
    const bool recip = T.transform == BBB_TRANSFORM_RECIPROCAL;

    // ---- tables: proposed sites, cell function of the current (fo) and proposed (fn) state --------------------
    double fmax = 0.0;
    for (int j = tid; j < k_new; j += CS_THREADS) {
        sx[j] = site_ref(B, S, s, prop, 0, j, c);
        sy[j] = site_ref(B, S, s, prop, 1, j, c);
        double v = value_ref(B, S, s, prop, T.param, j, c);
        if (recip) v = 1.0 / v;
        fn[j] = v;
        fmax = fmax > fabs(v) ? fmax : fabs(v);
        if (!(fabs(v) <= 1.79769313486231570815e308)) fmax = INFINITY;  // NaN / inf
    }
    for (int j = tid; j < k_old; j += CS_THREADS) {
        double v = value_ref(B, S, s, cur, T.param, j, c);
        if (recip) v = 1.0 / v;
        fo[j] = v;
        fmax = fmax > fabs(v) ? fmax : fabs(v);
        if (!(fabs(v) <= 1.79769313486231570815e308)) fmax = INFINITY;
    }
    for (int r = tid; r < RT; r += CS_THREADS) acc[r] = 0;
    if (tid == 0) { s_front = 0; s_back = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, fmax, o);
        fmax = fmax > other ? fmax : other;
    }
    if (lane == 0) s_red[warp] = fmax;
    __syncthreads();
    fmax = 0.0;
    for (int w = 0; w < CS_WARPS; ++w) fmax = fmax > s_red[w] ? fmax : s_red[w];
    __syncthreads();  // s_red is reused below
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
    // separately rounded products keep the lowest-index tie rule of numpy argmin / KDTree.query)
    IdxT *row = reinterpret_cast<IdxT *>(B.idx_cm[t]) + c * ((long long)(G + 15) & ~15LL);
    unsigned *list = B.change_list[t] + c * (long long)G;  // changed points from the front, re-derive queue from the back
    constexpr int VEC = 16 / (int)sizeof(IdxT);
    for (int g0 = tid * VEC; g0 < G; g0 += CS_THREADS * VEC) {
        const uint4 wv = *reinterpret_cast<const uint4 *>(row + g0);
        const IdxT *ow = reinterpret_cast<const IdxT *>(&wv);
        unsigned chg = 0, scn = 0;
        if (inner == BBB_MOVE_VALUES) {
#pragma unroll
            for (int q = 0; q < VEC; ++q)
                if (g0 + q < G && (int)ow[q] == rm) chg |= 1u << q;
        } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const int g = g0 + q;
                if (g >= G) continue;
                const int o = (int)ow[q];
                if (rm >= 0 && o == rm) {
                    scn |= 1u << q;
                } else if (ins >= 0) {
                    int mm = o - ((rm >= 0 && o > rm) ? 1 : 0);
                    mm += (mm >= ins) ? 1 : 0;
                    const double gx = T.points_x[g], gy = T.points_y[g];
                    const double dx = sx[mm] - gx, dy = sy[mm] - gy;
                    const double d_own = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    const double ex = nx - gx, ey = ny - gy;
                    const double d_new = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                    if (d_new < d_own || (d_new == d_own && ins < mm)) chg |= 1u << q;
                }
            }
        }
        if (chg) {
            int pos = atomicAdd(&s_front, __popc(chg));
            while (chg) {
                const int q = __ffs(chg) - 1;
                chg &= chg - 1;
                list[pos++] = ((unsigned)(g0 + q) << CS_M_BITS) | (unsigned)ins;
            }
        }
        if (scn) {
            int pos = atomicAdd(&s_back, __popc(scn));
            while (scn) {
                const int q = __ffs(scn) - 1;
                scn &= scn - 1;
                list[G - 1 - pos++] = (unsigned)(g0 + q);
            }
        }
    }
    __syncthreads();
    const int n_front = s_front, n_back = s_back;
    for (int i = tid; i < n_back; i += CS_THREADS) {
        const unsigned g = list[G - 1 - i];
        const double gx = T.points_x[g], gy = T.points_y[g];
        double best = INFINITY;
        int bi = 0;
        for (int j = 0; j < k_new; ++j) {
            const double dx = sx[j] - gx, dy = sy[j] - gy;
            const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d < best) { best = d; bi = j; }
        }
        // a moved cell that still owns the point: nothing changes (2-D site move: rm == ins, values kept)
        list[G - 1 - i] = (rm == ins && bi == ins) ? CS_NONE : ((g << CS_M_BITS) | (unsigned)bi);
    }
    __syncthreads();
    const int n_list = n_front + n_back;

    // ---- change-local forward: fixed-point accumulation of the ray changes of one ray tile ---------------------
    auto accumulate = [&](int tile) {
        const int32_t *cp = T.csc_colptr + (long long)tile * (G + 1);
        for (int base = 0; base < n_list; base += CS_CHUNK) {
            const int e = base + tid;
            PrepEntry pe{0, 0, 0.0};
            if (e < n_list) {
                const unsigned pk = e < n_front ? list[e] : list[G - 1 - (e - n_front)];
                if (pk != CS_NONE) {
                    const unsigned g = pk >> CS_M_BITS;
                    const int m = (int)(pk & ((1u << CS_M_BITS) - 1u));
                    pe.dlt = fn[m] - fo[(int)row[g]];
                    pe.j0 = cp[g];
                    pe.len = (pe.dlt != 0.0) ? cp[g + 1] - pe.j0 : 0;
                }
            }
            prep[tid] = pe;
            __syncthreads();
            const int n = min(CS_CHUNK, n_list - base);
            for (int i = warp; i < n; i += CS_WARPS) {
                const PrepEntry q = prep[i];
                for (int j = lane; j < q.len; j += 32) {
                    const int r = (int)T.csc_rows[q.j0 + j];
                    const double term = __dmul_rn(T.csc_data[q.j0 + j], q.dlt);
                    atomicAdd(reinterpret_cast<unsigned long long *>(acc + r), (unsigned long long)__double2ll_rn(term * scale));
                }
            }
            __syncthreads();
        }
    };

    double *res = B.res[t] + c * (long long)R;
    double dphi = 0.0;
    for (int tile = 0; tile < NT; ++tile) {
        if (tile > 0) {
            for (int r = tid; r < RT; r += CS_THREADS) acc[r] = 0;
            __syncthreads();
        }
        accumulate(tile);
        const int r0 = tile * RT, nr = min(RT, R - r0);
        for (int r = tid; r < nr; r += CS_THREADS) {
            const long long a = acc[r];
            if (a != 0) {
                const double d = (double)a * inv_scale;
                const double wr = (T.cov_kind == BBB_COV_VECTOR) ? T.cov_vec[r0 + r] : 1.0;
                dphi += wr * (d * (2.0 * res[r0 + r] + d));  // w (res + d)^2 - w res^2
            }
        }
        if (NT > 1) __syncthreads();
    }
    dphi = block_sum_to_thread0(dphi, s_red);

    // ---- likelihood ratio, decision, statistics (thread 0), then the commit by the whole block ----------------
    if (tid == 0) {
        const Phi3 old = load_phi(B, t, 0, c);
        store_phi(B, t, 1, c, Phi3{finite ? old.s0 + dphi : NAN, 0.0, 0.0});
        Rng rng;
        rng_begin(rng, ep, c, STREAM_STEP, true);
        s_accept = finish_chain(ep, c, rng) ? 1 : 0;
        rng_end(rng, ep, c);
    }
    __syncthreads();
    if (!s_accept) return;

    for (int tile = 0; tile < NT; ++tile) {
        if (NT > 1) {  // the accumulators of a multi-tile problem were recycled: rebuild the tile's
            for (int r = tid; r < RT; r += CS_THREADS) acc[r] = 0;
            __syncthreads();
            accumulate(tile);
        }
        const int r0 = tile * RT, nr = min(RT, R - r0);
        for (int r = tid; r < nr; r += CS_THREADS) {
            const long long a = acc[r];
            if (a != 0) res[r0 + r] += (double)a * inv_scale;
        }
        if (NT > 1) __syncthreads();
    }
    // nearest-site index: renumber the survivors when the edit shifts cell indices (death), then the changed points
    const bool renumber = (rm >= 0 || ins >= 0) && rm != ins && !(rm < 0 && ins >= k_old);
    if (renumber) {
        for (int g0 = tid * VEC; g0 < G; g0 += CS_THREADS * VEC) {
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
        __syncthreads();
    }
    for (int e = tid; e < n_list; e += CS_THREADS) {
        const unsigned pk = e < n_front ? list[e] : list[G - 1 - (e - n_front)];
        if (pk != CS_NONE) row[pk >> CS_M_BITS] = (IdxT)(pk & ((1u << CS_M_BITS) - 1u));
    }
}

int launch_chain_step(const EngineParams *dev, long long C, int t, int tile_rays, int K, int idx_bytes, cudaStream_t st) {
    const size_t smem = chain_step_smem(tile_rays, K);
    cudaError_t err;
    if (idx_bytes == 1) {
        err = cudaFuncSetAttribute(k_chain_step<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        k_chain_step<uint8_t><<<(unsigned)C, CS_THREADS, smem, st>>>(dev, t);
    } else {
        err = cudaFuncSetAttribute(k_chain_step<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        k_chain_step<uint16_t><<<(unsigned)C, CS_THREADS, smem, st>>>(dev, t);
    }
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

// res[c][r] = dpred[r][c] - dobs[r]
int launch_residual_transpose(const double *dpred, double *res, long long R, long long C, const double *dobs, cudaStream_t st) {
    return launch_transpose_t<double, true>(dpred, res, R, C, C, R, dobs, st);
}

}  // namespace bbb
