// Internal launch interface between the C ABI (abi.cu) and the kernels.
#pragma once
#include "common.cuh"

namespace bbb {

// Arguments of the 2-D rasterisation kernel.  `move == nullptr` means "all chains, current slots";
// otherwise only chains whose proposal changed the geometry of `space` are rasterised, reading the
// proposed slot (sel ^ 1) and writing the proposed index slot (idx_sel ^ 1).
struct RasterArgs {
    const double *sites;      // [slots][2][K][C]
    const int32_t *k;         // [slots][C]
    int K;                    // pad of the site arrays (n_dimensions_max)
    int k_rows;               // upper bound of k over the chains (<= K): sizes the shared-memory staging
    long long C;
    const uint8_t *sel;       // [C] or nullptr (slot 0)
    const uint8_t *idx_sel;   // [C] or nullptr (slot 0)
    int slot_xor;             // 1: proposed slots, 0: current slots
    const int32_t *move;      // [C] or nullptr
    const double *lpr;        // [C]
    const int32_t *edit;      // [2][C] (rm, ins)            -- incremental kernel only
    const double *edit_site;  // [2][C] inserted site        -- incremental kernel only
    int n_spaces, space;
    const double *px, *py;    // [G]
    int G;
    void *idx;                // [slots][G][C]
    int idx_bytes;            // 1 or 2
    int points_per_block;
    unsigned *rescan_list;    // [C][G] queue of grid points that lost their owner   -- incremental kernel only
    int32_t *rescan_count;    // [C] queue lengths (zero between steps)
};

struct SpmmArgs {
    const void *idx;          // [slots][G][C]
    int idx_bytes;
    const uint8_t *idx_sel;   // [C] or nullptr
    const double *values;     // [slots][n_params][K][C]
    const int32_t *k;         // [slots][C]
    const uint8_t *sel;       // [C] or nullptr
    int slot_xor;
    int n_params, param, K;
    int k_rows;               // upper bound of k over the chains (<= K): picks the chains-per-lane variant
    long long C;
    int G, R;
    const int32_t *indptr, *indices;
    const double *data;
    int transform;
    const double *dobs;
    const double *w;          // [R] or nullptr
    const int32_t *move;      // [C] or nullptr (all chains)
    const double *lpr;
    int n_spaces, space;
    double *dpred;            // [R][C] or nullptr
    double *partial;          // [n_groups][C] or nullptr
    int n_groups;             // ray groups = gridDim.y (spmm_ray_groups)
};

int raster_points_per_block(int G, long long C);
int launch_raster2d(const RasterArgs &a, cudaStream_t stream);
int launch_raster2d_inc(const RasterArgs &a, cudaStream_t stream);
int launch_idx_commit_copy(void *idx, int idx_bytes, const uint8_t *idx_sel, int G, long long C, cudaStream_t stream);
int launch_ray_spmm(const SpmmArgs &a, cudaStream_t stream);
int spmm_ray_groups(int R, long long C, int K, int idx_bytes, int n_sm);

// chain_kernels.cu
long long chain_local_tile_rays(long long n_data, int kmax, long long n_points);
struct HostedIO;  // hosted.cuh: the mapped hosted step (records in pinned host memory read / written by the kernel)
int launch_chain_step(const EngineParams &host_params, long long c_first, long long c_end, int target, int tile_rays,
                      long long n_data, int K, int idx_bytes, int order_slot, bool dependent_launch, const int32_t *guard,
                      cudaStream_t stream, const HostedIO *hosted = nullptr, bool proposer_finished = false);
long long chain_run_ctl_words(long long n_chains_engine);
// the fused persistent run (k_chain_run): n_iters iterations of the chains [c_first, c_end) in one launch of n_blocks
// blocks; ctl = chain_run_ctl_words(n_chains of the engine) int32 control words, zero before the launch (left zero by it)
int launch_chain_run(const EngineParams &host_params, long long c_first, long long c_end, long long n_iters, int target,
                     int tile_rays, long long n_data, int K, int idx_bytes, int32_t *ctl, int n_blocks, cudaStream_t stream);
int launch_transpose_idx(const void *in, void *out, int idx_bytes, long long rows, long long cols, long long in_stride,
                         long long out_stride, cudaStream_t stream);
int launch_residual_transpose(const double *dpred, double *res, long long R, long long C, const double *dobs,
                              cudaStream_t stream);

// step_kernels.cu
// (the step kernels take the engine parameters BY VALUE, as a __grid_constant__ kernel argument: every scalar of the
// problem spec and every buffer pointer is then a constant-bank read instead of a dependent global load)
// order_slot = parity * 4 + part of the block-order class counts (common.cuh), -1: no block order
// guard (device, may be NULL): a non-zero word makes the launch a no-op (hosted step, abi.cu)
// wait_done != NULL: launched as the programmatic dependent of the hosted chain kernel in front of it; the warp of chain c
// waits (bounded; *wait_failed is raised otherwise) until wait_done[c] == wait_seq
// finish_no_forward (chain-local engines): chains whose proposal needs no forward are decided by the proposal kernel
// itself; the chain kernel behind it must then be launched with proposer_finished = true
int launch_propose(const EngineParams &host_params, long long c_first, long long c_end, int order_slot, const int32_t *guard,
                   cudaStream_t stream, const int32_t *wait_done = nullptr, int wait_seq = 0, int32_t *wait_failed = nullptr,
                   bool finish_no_forward = false);
int launch_finish(const EngineParams &host_params, long long C, cudaStream_t stream);
int launch_corr_sums(const EngineParams *dev_params, long long C, int target, int proposal, cudaStream_t stream);
int launch_fused_steps(const EngineParams &host_params, long long C, long long n_steps, cudaStream_t stream);
int launch_init_states(const EngineParams *dev_params, long long C, cudaStream_t stream);
int launch_refresh_finish(const EngineParams *dev_params, long long C, cudaStream_t stream);
int launch_gather_space(const EngineParams *dev_params, long long C, int space, int kmax, int32_t *k_out,
                        double *sites_out, double *values_out, cudaStream_t stream);
long long state_image_words(const EngineParams &host_params, const int32_t *kb);
long long state_image_rows(const EngineParams &host_params, const int32_t *kb);
// chains [c_first, c_end): the image has rows of (c_end - c_first) words
int launch_state_image(const EngineParams &host_params, int mode, const int32_t *kb, void *image, int32_t *differ,
                       long long c_first, long long c_end, cudaStream_t stream);
// ragged image of the chains [c_first, c_end) (only the cells in use; layout: step_kernels.cu); mode as above
int ragged_header_rows(const EngineParams &host_params);
int launch_state_image_ragged(const EngineParams &host_params, int mode, void *image, int32_t *differ, long long c_first,
                              long long c_end, cudaStream_t stream);
// records of the mapped hosted step (hosted.cuh): all chains -> io.out; flagged chains of io.in -> engine slot 0
int launch_hosted_pack(const EngineParams &host_params, const HostedIO &io, cudaStream_t stream);
int launch_hosted_unpack(const EngineParams &host_params, const HostedIO &io, cudaStream_t stream);
int launch_max_k(const EngineParams *dev_params, long long C, int32_t *out /* device [n_spaces], zeroed */, cudaStream_t stream);
// save point: states + noise + temperature sections of the chain-major image (step_kernels.cu), d_pred of a chain-local
// ray target from its resident residuals
long long save_point_state_words(const EngineParams &host_params, const int32_t *kb);
int launch_save_states(const EngineParams &host_params, const int32_t *kb, void *image, cudaStream_t stream);
int launch_save_dpred_local(const double *res, const double *dobs, long long C, int R, double *out, cudaStream_t stream);
int launch_transpose_f64(const double *in, double *out, long long rows, long long cols, cudaStream_t stream);
int launch_small_dpred(const EngineParams *dev_params, long long C, int target, double *dpred_out,
                       cudaStream_t stream);
int launch_misfit_logdet(const EngineParams *dev_params, long long C, double *misfit_out,
                         double *logdet_out, cudaStream_t stream);
int launch_pt_swap(const double *gathered, long long n_total, uint64_t seed, long long round, long long id0,
                   long long n_local, double *temperature_out, int32_t *n_accepted_out, const double *tape,
                   cudaStream_t stream);
int launch_raster1d(const double *sites, const double *values, const int32_t *k, int K, long long C,
                    const double *x, int R, int32_t *idx_out, double *field_out, cudaStream_t stream);
int launch_dense_forward(const double *G, const double *m, int R, int M, long long C, double *dpred_out,
                         cudaStream_t stream);
int launch_loglike(const double *dpred, const double *dobs, const double *sigma, int R, long long C,
                   double *misfit_out, double *logdet_out, cudaStream_t stream);
int launch_logprior(const bbb_space_spec &space, const double *values, const int32_t *k, long long C,
                    double *out, cudaStream_t stream);
int launch_phi_from_partial(const double *partial, int n_tiles, long long C, double *phi_out, cudaStream_t stream);

}  // namespace bbb
