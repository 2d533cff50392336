/*
 * bbb200.h -- C ABI of libbbb200.so, the B200-native batched reversible-jump MCMC engine.
 *
 * This is the drop-in boundary for ONE path of BayesBay 0.3.11 (reference paths relative to
 * /root/reference): the chain pool + chain step that `Sampler.advance_chain`
 * (src/bayesbay/samplers/_samplers.py:191-239) farms out to one OS process per chain, i.e.
 * `BaseMarkovChain.advance_chain/_next_iteration` (src/bayesbay/_markov_chain.py:204-297) and
 * everything it calls (perturbations/, discretization/, prior/, likelihood/, _utils_1d.pyx).
 * A `Sampler` subclass whose `run()` calls these entry points instead of `advance_chain` is a
 * plug-in of the reference's own extension seam (`Sampler.run` is abstract,
 * samplers/_samplers.py:152-162; used by `BaseBayesianInversion.run`, _bayes_inversion.py:236-249).
 *
 * Conventions
 *   - plain C types only; every pointer named "device" is a CUDA device pointer owned by the
 *     caller (PyTorch allocates, this library never frees caller memory);
 *   - every entry returns 0 on success or a negative bbb_status; bbb_last_error() gives the
 *     message (thread-local);
 *   - all launches are asynchronous on the `stream` argument (a cudaStream_t passed as void*);
 *   - chains are the fastest-varying (coalesced) dimension of every per-chain array: an array
 *     documented as [A][B][C] is row-major with C = n_chains innermost.
 */
#ifndef BBB200_H
#define BBB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBB_ABI_VERSION 6
#define BBB_ORDER_BUCKETS 10

#define BBB_MAX_SPACES 4
#define BBB_MAX_PARAMS 8
#define BBB_MAX_TARGETS 4
#define BBB_MAX_LINEAR_PARAMS 16
#define BBB_CSC_LEVELS 3    /* merged-column levels of the CSC copy: groups of 1, 4 and 16 consecutive columns */

typedef enum {
    BBB_OK = 0,
    BBB_ERR_INVALID = -1,   /* invalid spec / argument */
    BBB_ERR_SHAPE = -2,     /* buffer shape / missing buffer */
    BBB_ERR_CUDA = -3,      /* a CUDA runtime call failed (message carries the CUDA error) */
    BBB_ERR_UNSUPPORTED = -4
} bbb_status;

/* parameter-space kinds: ParameterSpace (parameterization/_parameter_space.py), Voronoi1D,
 * Voronoi2D (discretization/_voronoi.py:471, :1283) */
enum { BBB_SPACE_PLAIN = 0, BBB_SPACE_VORONOI1D = 1, BBB_SPACE_VORONOI2D = 2 };
/* prior kinds: UniformPrior(a=vmin,b=vmax), GaussianPrior(a=mean,b=std), LaplacePrior(a=mean,b=scale)
 * (prior/_prior.py:300, :484, :685) */
enum { BBB_PRIOR_UNIFORM = 0, BBB_PRIOR_GAUSSIAN = 1, BBB_PRIOR_LAPLACE = 2 };
enum { BBB_BIRTH_FROM_PRIOR = 0, BBB_BIRTH_FROM_NEIGHBOUR = 1 };
/* forward operators (the user forward functions of the tutorials) */
enum { BBB_FWD_RAY2D = 0, BBB_FWD_LOOKUP1D = 1, BBB_FWD_LINEAR = 2, BBB_FWD_POLYNOMIAL = 3 };
enum { BBB_TRANSFORM_IDENTITY = 0, BBB_TRANSFORM_RECIPROCAL = 1 };
/* data covariance: hierarchical sigma (Target(std_min,...)), fixed scalar / per-datum vector / full matrix inverse
 * covariance (likelihood/_target.py:136-170: `covariance_mat_inv * vector` or, for ndim == 2, `covariance_mat_inv @ vector`) */
enum { BBB_COV_HIERARCHICAL = 0, BBB_COV_SCALAR = 1, BBB_COV_VECTOR = 2, BBB_COV_MATRIX = 3 };
/* inner move kinds; move id = 4 * space + kind, noise move id = 4 * n_spaces */
enum { BBB_MOVE_BIRTH = 0, BBB_MOVE_DEATH = 1, BBB_MOVE_VALUES = 2, BBB_MOVE_SITES = 3 };

typedef struct {
    int32_t kind;            /* BBB_SPACE_* */
    int32_t ndim;            /* 0, 1, 2 */
    int32_t trans_d;
    int32_t kmin, kmax;      /* kmax is also the pad K of the SoA arrays */
    int32_t kinit_max;       /* int((kmax-kmin)*init_range + kmin), discretization/_voronoi.py:159 */
    int32_t birth_from;      /* BBB_BIRTH_FROM_* */
    int32_t n_params;
    double vmin[2], vmax[2], site_std[2];
    double w_birth, w_death, w_values, w_sites; /* inner weights 1,1,3,1 (0 = move absent) */
    double w_outer;          /* weight of the space in the chain's outer choice; 0 = the sum of the inner weights (what
                              * BayesianInversion builds, _bayes_inversion.py:458-466); another value after
                              * BaseBayesianInversion.set_perturbation_funcs (:117-171) */
    int32_t prior_kind[BBB_MAX_PARAMS];
    double prior_a[BBB_MAX_PARAMS], prior_b[BBB_MAX_PARAMS];
    double perturb_std[BBB_MAX_PARAMS], perturb_std_birth[BBB_MAX_PARAMS];
    /* position-dependent hyper-parameters (Prior(..., position=...), prior/_prior.py:257-297; Voronoi1D only):
     * n_pos[p] > 0 -> pos_table[p] is a device array [5][n_pos] with rows position (ascending), a, b,
     * perturb_std, perturb_std_birth, linearly interpolated at the cell's site (interpolate_linear_1d,
     * _utils_1d.pyx:68-84) and overriding the scalars above */
    int32_t n_pos[BBB_MAX_PARAMS];
    const double *pos_table[BBB_MAX_PARAMS];
} bbb_space_spec;

typedef struct {
    int32_t n_data;
    int32_t cov_kind;        /* BBB_COV_* */
    double std_min, std_max, std_perturb_std;
    int32_t correlated;      /* hierarchical only: Target(noise_is_correlated=True), tridiagonal inverse covariance
                              * inverse_covariance(sigma, r, n) (_utils_1d.pyx:198-211, likelihood/_target.py:167-170) */
    double corr_min, corr_max, corr_perturb_std;
    double cov_scalar;
    const double *dobs;      /* device [n_data] */
    const double *cov_vec;   /* BBB_COV_VECTOR: device [n_data]; BBB_COV_MATRIX: device [n_data][n_data] row-major; else NULL */
    int32_t fwd_kind;        /* BBB_FWD_* */
    int32_t space;           /* parameter space the forward reads */
    int32_t param;           /* ray2d / lookup1d: parameter index inside the space */
    int32_t transform;       /* ray2d: BBB_TRANSFORM_* applied to the cell value */
    /* ray2d: CSR ray matrix shared by all chains + query points of the rasterisation */
    int32_t n_points;
    int64_t nnz;
    const int32_t *indptr;   /* device [n_data + 1] */
    const int32_t *indices;  /* device [nnz] */
    const double *data;      /* device [nnz] */
    const double *points_x;  /* device [n_points] */
    const double *points_y;  /* device [n_points] */
    /* lookup1d: data positions; polynomial: abscissae of d_pred[r] = sum_j values[param][j] * x[r]^j over the
     * k current cells of a (possibly trans-dimensional) plain space: np.vander(x, k, True) @ m
     * (docs/source/tutorials/03_transd_polyfit.ipynb cell 5) */
    const double *x;         /* device [n_data] */
    /* linear: d_pred = G @ concat(values[p] for p in lin_params), each of length kmax (fixed k) */
    int32_t n_lin_params;
    int32_t lin_params[BBB_MAX_LINEAR_PARAMS];
    const double *G;         /* device [n_data][n_lin_params * kmax] row-major */
    /* ray2d, optional (csc_tile_rays > 0): the SAME ray matrix in ray-tile-blocked CSC form, which enables the
     * chain-local step (one thread block per chain; a proposal edits one cell, so only the columns of the grid
     * points whose nearest site changed are visited: d_pred' = d_pred + sum_g G[:, g] * (f(v'(g)) - f(v(g))),
     * the change-local form of `jacobian @ (1 / vel[nearest])`, 31_sw_tomography.ipynb cell 12).
     * Rays are split into csc_n_tiles tiles of csc_tile_rays rays (bbb_chain_local_tile_rays(); the last tile may
     * be shorter).  Level l (0 <= l < csc_n_levels <= BBB_CSC_LEVELS) holds the matrix with every aligned group of
     * 4^l consecutive columns summed into one (n_l = ceil(n_points / 4^l) columns; level 0 is the matrix itself):
     * entries sorted by (tile, column, row), the entries of column b inside tile q are
     * [csc_colptr[l][q * (n_l + 1) + b], csc_colptr[l][q * (n_l + 1) + b + 1]).  Merged levels pay off when
     * consecutive point indices are spatial neighbours (the Python engine orders the points along a Morton curve). */
    int32_t csc_tile_rays, csc_n_tiles, csc_n_levels;
    const int32_t *csc_colptr[BBB_CSC_LEVELS]; /* device [csc_n_tiles][n_l + 1] */
    const uint16_t *csc_rows[BBB_CSC_LEVELS];  /* device [nnz_l] ray index inside its tile */
    const double *csc_data[BBB_CSC_LEVELS];    /* device [nnz_l] */
    double csc_bound;          /* max over rays of sum_g |data|: bounds any change of d_pred per unit change of the
                                * cell function; scales the fixed-point ray accumulators */
    /* optional (NULL: none): bounding boxes (xmin, xmax, ymin, ymax) of the aligned groups of 16 consecutive grid
     * points.  The change-local rasterisation skips a group that has one owner when the inserted cell cannot be
     * nearer than that owner anywhere in the group's box (d_new - d_own is affine in the point). */
    const double *group_box;   /* device [ceil(n_points / 16)][4] */
} bbb_target_spec;

typedef struct {
    int32_t n_spaces, n_targets;
    bbb_space_spec spaces[BBB_MAX_SPACES];
    bbb_target_spec targets[BBB_MAX_TARGETS];
    uint64_t seed;           /* Philox key */
    double w_noise;          /* weight of the noise perturbation in the outer choice; < 0 = the default (1 when a target
                              * is hierarchical), 0 = no noise move (set_perturbation_funcs without it) */
} bbb_problem_spec;

/* Device buffers of one engine (all owned by the caller).  "slot" arrays are double-buffered:
 * sel[s][c] names the slot holding chain c's CURRENT state of space s, the other slot receives
 * the proposal; accepting a proposal flips the bit (no copy). */
typedef struct {
    int64_t n_chains;        /* chains resident on this device */
    int64_t chain_id0;       /* global id of local chain 0 (Philox counter word) */
    /* per parameter space */
    double *sites[BBB_MAX_SPACES];    /* [2][ndim][K][C] */
    double *values[BBB_MAX_SPACES];   /* [2][n_params][K][C] */
    int32_t *k[BBB_MAX_SPACES];       /* [2][C] */
    uint8_t *sel[BBB_MAX_SPACES];     /* [C] */
    /* per target */
    double *sigma[BBB_MAX_TARGETS];   /* [2][C]  (0 = current, 1 = proposed) hierarchical only */
    double *corr[BBB_MAX_TARGETS];    /* [2][C]  noise correlation (correlated targets only) */
    double *phi[BBB_MAX_TARGETS];     /* [2][3][C] residual sums with the noise parameters factored out:
                                       * sum_i w_i r_i^2, r_0^2 + r_{n-1}^2, sum_i r_i r_{i+1} (the last two only
                                       * for correlated noise) */
    double *dpred_scratch[BBB_MAX_TARGETS]; /* [n_data][C], ray2d targets with correlated noise only */
    void *cell_idx[BBB_MAX_TARGETS];  /* ray2d: [2][G][C] uint8 (K <= 256) or uint16; slot 0 = current, slot 1 = proposed */
    uint8_t *idx_sel[BBB_MAX_TARGETS];/* ray2d: [C] 1 = the current index is still in slot 1 (accepted, not yet committed) */
    double *partial[BBB_MAX_TARGETS]; /* ray2d: [n_ray_tiles][C], n_ray_tiles = bbb_ray_tiles(n_data) */
    uint32_t *rescan_list[BBB_MAX_TARGETS]; /* ray2d: [C][G] scratch queue of the change-local rasteriser */
    int32_t *rescan_count[BBB_MAX_TARGETS]; /* ray2d: [C], must be zero when bound */
    /* ray2d, chain-local step (all three or none): chain-major caches of the CURRENT state */
    void *idx_cm[BBB_MAX_TARGETS];    /* [C][bbb_idx_row_stride(G)] uint8 (K <= 256) or uint16 nearest-site index */
    double *res[BBB_MAX_TARGETS];     /* [C][n_data] residual d_pred - d_obs */
    uint32_t *change_list[BBB_MAX_TARGETS]; /* [C][G + ceil(G / 16)] scratch: grid points whose nearest site changes */
    /* chain scalars */
    double *temperature;     /* [C] */
    int64_t *iteration;      /* [C] iterations done (Philox counter) */
    /* last proposal / decision (inputs of the accept stage, outputs for step-replay parity) */
    int32_t *move;           /* [C] move id */
    int32_t *edit;           /* [2][C] (remove index or -1, insert index or -1) */
    double *edit_site;       /* [2][C] site of the inserted cell */
    double *edit_vals;       /* [BBB_MAX_PARAMS][C] values of the inserted cell */
    double *log_prob_ratio;  /* [C] */
    double *log_like_ratio;  /* [C] */
    int32_t *accepted;       /* [C] */
    /* statistics: _save_statistics (src/bayesbay/_markov_chain.py:136-157) */
    int64_t *n_proposed;     /* [4*n_spaces+1][C] */
    int64_t *n_accepted;     /* [4*n_spaces+1][C] */
    /* optional (NULL: not counted), ZERO-INITIALISED by the caller: what the reference would have reported as
     * exceptions (`statistics["exceptions"]`, src/bayesbay/_markov_chain.py:205-247).  Row 0: proposals whose likelihood
     * ratio was not finite (NaN / infinite forward or misfit; rejected).  Row 1: rejection-sampling loops of a site or
     * noise move (unbounded in the reference, _voronoi.py:171-188, _data_noise.py:61-76) that gave up after 100 000
     * tries and kept the old value. */
    int64_t *n_exceptions;   /* [2][C] */
    /* optional recorded uniforms (step-replay parity); NULL selects Philox */
    const double *tape;      /* [C][tape_stride] */
    int64_t tape_stride;
    int64_t *tape_cursor;    /* [C] tape position (tape mode) / draws consumed in the step (Philox mode) */
    /* chain-local step, optional (NULL: blocks take the chains in index order): scratch for the longest-first block
     * order of the chain kernel -- the proposal kernel files every chain under one of BBB_ORDER_BUCKETS cost classes.
     * [BBB_ORDER_BUCKETS][C] chain ids followed by [2][4][16] class counts; ZERO-INITIALISED by the caller.  The order
     * never changes a result (chains do not interact inside a step), only when a chain's block is scheduled. */
    int32_t *block_order;    /* [BBB_ORDER_BUCKETS * C + 128] */
} bbb_buffers;

typedef struct bbb_engine bbb_engine;

int bbb_abi_version(void);
const char *bbb_last_error(void);
/* number of ray tiles the partial-misfit buffer must hold for a ray target with n_data rays */
int64_t bbb_ray_tiles(int64_t n_data);
/* chain-local step: elements per chain row of idx_cm (G rounded up to a 16-byte multiple for either index type) */
int64_t bbb_idx_row_stride(int64_t n_points);
/* chain-local step: rays per CSC tile for a ray target with n_data rays whose space pads to kmax cells (the
 * per-chain ray accumulators live in shared memory); 0 if the chain-local step cannot run this shape */
int64_t bbb_chain_local_tile_rays(int64_t n_data, int32_t kmax, int64_t n_points);

/* Engine life cycle.  Replaces the chain pool `Sampler.advance_chain`
 * (samplers/_samplers.py:191-239).  The spec is copied; device pointers inside it are borrowed. */
int bbb_engine_create(const bbb_problem_spec *spec, int device, bbb_engine **out);
int bbb_engine_bind_buffers(bbb_engine *e, const bbb_buffers *buffers);
void bbb_engine_destroy(bbb_engine *e);

/* Draw starting states on the device: `MarkovChain._init_starting_state`
 * (src/bayesbay/_markov_chain.py:364-395) -> Voronoi._initialize / ParameterSpace._initialize /
 * Prior.initialize / Target.initialize.  Also evaluates the forward of every chain. */
int bbb_engine_init_states(bbb_engine *e, void *stream);
/* Re-evaluate rasterisation, forward and misfit of the CURRENT states (after the caller wrote
 * starting states, temperatures, ... into the bound buffers). */
int bbb_engine_refresh(bbb_engine *e, void *stream);
/* Performance hint: an upper bound of the current number of cells over all resident chains of `space`
 * (kmin <= bound <= kmax).  The engine sizes shared-memory staging by it and raises it by one per step on its
 * own; callers may tighten it at a synchronisation point.  A bound BELOW the true maximum of the resident chains
 * (measured on the device by the call; it blocks) is refused with BBB_ERR_INVALID.
 * bbb_engine_sync_k_bound measures the true maximum of every trans-dimensional space on the device (one kernel, blocks
 * on `stream`), sets the bounds to it and reports them in bounds_out[n_spaces] (may be NULL). */
int bbb_engine_set_k_bound(bbb_engine *e, int space, int32_t bound);
int bbb_engine_sync_k_bound(bbb_engine *e, int32_t *bounds_out, void *stream);
/* Advance every chain n_steps iterations: `BaseMarkovChain.advance_chain`
 * (src/bayesbay/_markov_chain.py:249-297) for all chains at once. */
int bbb_engine_step(bbb_engine *e, int64_t n_steps, void *stream);
/* One stage of one step, for per-kernel timing and profiling: 0 proposal, 1 rasterisation of the
 * proposed geometry, 2 ray forward + misfit, 3 likelihood ratio / accept / commit.  Calling stages
 * 0,1,2,3 in order is exactly one iteration of bbb_engine_step (ray-tomography problems only).
 * Chain-local engines do rasterisation, forward, misfit and accept in ONE kernel: stage 2 runs it, stages 1
 * and 3 launch nothing. */
int bbb_engine_stage(bbb_engine *e, int stage, void *stream);
/* 1 if this engine runs the chain-local step (one ray target with CSC arrays and chain-major caches bound,
 * uncorrelated noise, kmax <= 4096, n_points <= 2^17, BBB_FULL_FORWARD != 1), else 0. */
int bbb_engine_chain_local(bbb_engine *e);
/* Chain-local engines update residuals and misfit incrementally; every `every` steps (default 256, 0 = never)
 * bbb_engine_step re-evaluates forward and misfit of the current states from scratch (full SpMM) so rounding
 * cannot accumulate.  `steps_since` positions the schedule (checkpoint/resume). */
int bbb_engine_set_resync(bbb_engine *e, int64_t every, int64_t steps_since);
int64_t bbb_engine_steps_since_resync(bbb_engine *e);
/* The from-scratch re-evaluation itself (forward + misfit of the current states; the nearest-site index is kept).
 * Also brings the [G][C] index slot 0 in line with the chain-major index.  No-op for other engines. */
int bbb_engine_resync(bbb_engine *e, void *stream);
/* Simulated-annealing schedule `SimulatedAnnealing.on_begin_iteration` (samplers/_samplers.py:402-414):
 * while enabled (temperature_start > 0) every chain's temperature of iteration `it` is
 * max(1, temperature_start * exp(-cooling_rate * it)) for it < min(cooling_iterations, burnin_iterations)
 * and 1 afterwards.  temperature_start == 0 disables the schedule.  Synchronises the device. */
int bbb_engine_set_annealing(bbb_engine *e, double temperature_start, double cooling_rate,
                             int64_t cooling_iterations, int64_t burnin_iterations);
/* Compact the CURRENT state of every chain into caller arrays (`_save_state`,
 * src/bayesbay/_markov_chain.py:115-134): k [C] int32, sites [ndim][K][C], values [P][K][C]. */
int bbb_engine_gather_space(bbb_engine *e, int space, int32_t *k_out, double *sites_out,
                            double *values_out, void *stream);
/* Save point (`_save_state`, src/bayesbay/_markov_chain.py:115-134, for all chains at once): the CURRENT state of
 * every chain as ONE chain-major image of 8-byte words in device memory, so that a save point is one asynchronous
 * device -> host copy and a chain's sample is a contiguous slice of the host copy.  Sections, in this order:
 *   per space s: k [C] (int64) | sites [C][kb[s]][ndim] | values [C][n_params][kb[s]]   (cells j >= k hold 0);
 *   per hierarchical target: sigma [C] | correlation [C] (correlated targets only);   temperature [C];
 *   if want_dpred: per target d_pred [C][n_data] (the "{target}.dpred" entries the reference saves by default).
 * kb[s] >= the largest k of space s (e.g. n_dimensions_max).  `scratch` ([n_data][C] doubles, may be NULL) is only
 * needed for d_pred of targets other than the ray target of a chain-local engine. */
int64_t bbb_engine_save_point_bytes(bbb_engine *e, const int32_t *kb, int32_t want_dpred);
int bbb_engine_save_point(bbb_engine *e, const int32_t *kb, int32_t want_dpred, void *image, double *scratch, void *stream);
/* Chain-state image: the CURRENT state of every chain packed into ONE buffer of 8-byte words, so that moving the
 * chains between host and device -- what the reference's pool does with pickles on every advance_chain
 * (samplers/_samplers.py:225-237) -- is one copy each way.  Rows of C words each: per space k (int64), sites
 * [ndim][kb[s]] and values [n_params][kb[s]] (cells j >= k are 0; kb[s] <= kmax is the number of cell rows that
 * travel, at least the largest k); per hierarchical target sigma (and correlation); temperature; iteration (int64).
 *   pack     engine -> image (device buffer of bbb_engine_state_image_bytes());
 *   compare  sets *differ (device int32, caller zeroes it) when some chain's state is not bit-identical to the image:
 *            a caller that kept the chains resident uses it to learn whether its derived caches are still valid;
 *   unpack   image -> engine (slot 0), then re-derives rasterisation, forward and misfit (bbb_engine_refresh). */
int64_t bbb_engine_state_image_bytes(bbb_engine *e, const int32_t *kb);
int bbb_engine_pack_states(bbb_engine *e, const int32_t *kb, void *image, void *stream);
int bbb_engine_compare_states(bbb_engine *e, const int32_t *kb, const void *image, int32_t *differ, void *stream);
int bbb_engine_unpack_states(bbb_engine *e, const int32_t *kb, const void *image, void *stream);
/* The same image for the chains [c_first, c_end) only: rows of (c_end - c_first) words. */
int bbb_engine_pack_states_range(bbb_engine *e, const int32_t *kb, int64_t c_first, int64_t c_end, void *image,
                                 void *stream);
/* Hosted step: ONE iteration on chains the CALLER holds in pinned host memory -- the reference's
 * `Sampler.advance_chain` round trip (samplers/_samplers.py:225-237: chains pickled to the pool, advanced, pickled
 * back), with the copies overlapped with the kernels.  The chain set is split into bbb_engine_hosted_parts() parts
 * (first_chain[0..n], n <= 4; chain-local engines only); a HOSTED image is the concatenation of the parts' images
 * (bbb_engine_pack_states_range layout; part h starts at word rows(kb) * first_chain[h]).  Per part, on its own
 * internal stream: host_in -> dev_stage, compared on the device with the resident states the derived caches belong
 * to, proposal kernel + chain kernel (no-ops for a part whose states differ), packed with kb_out rows into dev_out,
 * dev_out -> host_out.  kb_out[s] (written by the call) = largest k in host_in + 1 (the cell a step can add) rounded
 * up to a multiple of 8, capped at kmax.  The work of a call is replayed from a CUDA graph captured the first time a
 * combination of image sizes and buffers is seen (BBB_HOSTED_GRAPH=0: enqueued call by call).  host_in / host_out (pinned) and dev_stage / dev_out (device) hold bbb_engine_state_image_bytes(kb_in) /
 * (kb_out) bytes and must be four different buffers.  Asynchronous: bbb_engine_step_hosted_wait synchronises
 * `stream`, and if the uploaded states of some parts differed from the resident ones writes those parts into the
 * engine, re-derives the caches (bbb_engine_refresh), advances those parts and downloads them before it returns
 * (*n_redone = number of such parts; 0 on the fast path).  Every chain has advanced exactly one iteration and
 * host_out holds the new states when the wait returns. */
int bbb_engine_hosted_parts(bbb_engine *e, int64_t *first_chain /* [5] */);
int bbb_engine_step_hosted(bbb_engine *e, const int32_t *kb_in, const void *host_in, int32_t *kb_out, void *host_out,
                           void *dev_stage, void *dev_out, void *stream);
int bbb_engine_step_hosted_wait(bbb_engine *e, int32_t *n_redone, void *stream);
/* The hosted step on a RAGGED image: only the cells in use travel, as in the pickles the reference's pool moves
 * (samplers/_samplers.py:225-237: a chain's arrays hold k cells, not n_dimensions_max).  A ragged hosted image is
 * bbb_engine_hosted_ragged_bytes() bytes; part h (bbb_engine_hosted_parts) starts at byte h * bytes / n_parts and
 * holds, for its n_c chains, 8-byte words:
 *   header, rows of n_c words: k of every space (int64) | sigma (and correlation) per hierarchical target |
 *           temperature | iteration (int64);
 *   then per space, chain after chain: the chain's k cells, (ndim + n_params) words each (site coordinates, then
 *           the parameter values) -- chain lc of space s starts (ndim + n_params) * (k[0] + ... + k[lc - 1]) words into
 *           the space's block, and the blocks of the spaces follow one another without gaps.
 * bbb_engine_pack_states_ragged writes the current states of all chains in this format (device buffer).
 * bbb_engine_step_hosted_ragged is bbb_engine_step_hosted for such images: per part the header's k rows (read on the
 * host) say how many words go up; what comes down is bounded by one more cell per chain of a trans-dimensional
 * space; both are rounded up to 32 KiB (the buffers hold the full size anyway), so that the captured graphs change
 * rarely.  *h2d_bytes / *d2h_bytes (may be NULL) = bytes copied each way by this call.  All four buffers hold
 * bbb_engine_hosted_ragged_bytes() bytes.  The wait is bbb_engine_step_hosted_wait. */
int64_t bbb_engine_hosted_ragged_bytes(bbb_engine *e);
/* MAPPED hosted step: the same round trip with the chain kernel itself moving the states over the bus.  The caller's
 * chains live in pinned (page-locked, device-mapped) host memory as fixed-stride RECORDS, chain c at word
 * c * bbb_engine_hosted_record_words():
 *   header: k of every space (int64) | sigma (and correlation) per hierarchical target | temperature | iteration (int64);
 *   then per space s (room for kmax cells; only the k in use are read / written, so only they cross the bus):
 *   cell j at (ndim + n_params) * j words into the space's block: site coordinates, then parameter values.
 * bbb_engine_step_hosted_mapped: every chain's block reads its record of host_in straight from host memory (requested
 * when the block starts, compared with the resident state the caches belong to before anything is decided), advances
 * the chain one iteration and writes the new record to host_out -- no staging buffers, no copy-engine transfers, the
 * traffic overlaps the step.  The proposal of the NEXT iteration is drawn behind the call (counter-based generators:
 * a pure function of the resident state), so it overlaps the host's turn-around; any other call on the engine
 * invalidates it and the next hosted call draws it again.  A chain whose record is not bit-identical to the resident
 * state is left untouched and flagged; bbb_engine_step_hosted_mapped_wait (which blocks until the step is done) then
 * writes those records into the engine, re-derives the caches (bbb_engine_refresh), advances exactly those chains and
 * reports their number in *n_redone (0 on the fast path).  bbb_engine_pack_states_records writes the current states as
 * records (image: device or mapped host memory of n_chains records).  bbb_engine_hosted_mapped_bytes: bytes the
 * kernels have read from / written to host memory so far.  Not for tape-driven engines. */
int64_t bbb_engine_hosted_record_words(bbb_engine *e);
int bbb_engine_pack_states_records(bbb_engine *e, void *image, void *stream);
int bbb_engine_step_hosted_mapped(bbb_engine *e, const void *host_in, void *host_out, void *stream);
int bbb_engine_step_hosted_mapped_wait(bbb_engine *e, int32_t *n_redone, void *stream);
int bbb_engine_hosted_mapped_bytes(bbb_engine *e, int64_t *h2d_bytes, int64_t *d2h_bytes);
int bbb_engine_pack_states_ragged(bbb_engine *e, void *image, void *stream);
int bbb_engine_step_hosted_ragged(bbb_engine *e, const void *host_in, void *host_out, void *dev_stage, void *dev_out,
                                  int64_t *h2d_bytes, int64_t *d2h_bytes, void *stream);
/* d_pred of the CURRENT states of one target, [n_data][C] (the "{target}.dpred" cache entry,
 * likelihood/_log_likelihood.py:311-320). */
int bbb_engine_dpred(bbb_engine *e, int target, double *dpred_out, void *stream);
/* Per-chain totals over targets with each chain's own sigma: misfit and log|C_e|
 * (`LogLikelihood._get_misfit_and_det`, likelihood/_log_likelihood.py:308-332), both [C]. */
int bbb_engine_misfit_logdet(bbb_engine *e, double *misfit_out, double *logdet_out, void *stream);

/* The rows this engine contributes to a parallel-tempering swap round: [C][3] (misfit, logdet, T), written by one
 * kernel (what the caller all-gathers over NCCL before bbb_pt_swap). */
int bbb_engine_swap_rows(bbb_engine *e, double *rows_out, void *stream);

/* Parallel-tempering swap round: `ParallelTempering.on_end_advance_chain`
 * (samplers/_samplers.py:334-342).  `gathered` is [n_total][3] (misfit, logdet, T) for ALL chains
 * of the job in global chain order (all-gathered by the caller over NCCL); every rank runs the
 * same n_total sequential attempts from the shared Philox key (seed, round) and writes the new
 * temperatures of its own chains [id0, id0 + n_local) to temperature_out [n_local] (which may be the engine's own
 * temperature buffer); the temperature column of `gathered` is updated in place.  Up to 8192 chains the attempts are
 * replayed in parallel waves of pairwise disjoint attempts (bit-identical to the sequential loop).
 * If `tape` is not NULL the uniforms come from it (3 per attempt). */
int bbb_pt_swap(double *gathered, int64_t n_total, uint64_t seed, int64_t round,
                int64_t id0, int64_t n_local, double *temperature_out, int32_t *n_accepted_out,
                const double *tape, void *stream);

/* ---- stand-alone kernels (parity entry points; the engine uses the same device code) ---- */
/* F1: nearest Voronoi site of every query point for every chain.
 * Voronoi2D.interpolate_tessellation (discretization/_voronoi.py:1429-1454) = KDTree.query.
 * sites [2][K][C], k [C], points [G] -> idx [G][C] (uint8 if K <= 256 else uint16). */
int bbb_raster2d(const double *sites, const int32_t *k, int32_t K, int64_t C, const double *points_x,
                 const double *points_y, int32_t G, void *idx_out, void *stream);
/* F1/F4: 1-D nearest-nuclei lookup, interpolate_nearest_1d (_utils_1d.pyx:87-114,171-192).
 * sites [K][C] sorted, values [K][C], x [R] -> idx [R][C] int32 and field [R][C]. */
int bbb_raster1d(const double *sites, const double *values, const int32_t *k, int32_t K, int64_t C,
                 const double *x, int32_t R, int32_t *idx_out, double *field_out, void *stream);
/* F2: straight-ray forward, d_pred[r][c] = sum_j data[j] * f(values[idx[indices[j]][c]][c])
 * (31_sw_tomography.ipynb cell 12: jacobian @ (1 / vel[nearest])).  Writes d_pred [R][C] if not
 * NULL and phi[c] = sum_r w_r (d_pred - dobs)^2 if not NULL (w = NULL means 1). */
int bbb_ray_forward(const void *idx, const double *values, const int32_t *k, int32_t K, int64_t C,
                    int32_t G, int32_t R, const int32_t *indptr, const int32_t *indices,
                    const double *data, int32_t transform, const double *dobs, const double *w,
                    double *dpred_out, double *phi_out, double *partial_scratch, void *stream);
/* F3: dense linear forward d_pred = G @ m, m [M][C] -> d_pred [R][C] (02_hierarchical_polyfit.ipynb
 * cell 12). */
int bbb_dense_forward(const double *G, const double *m, int32_t R, int32_t M, int64_t C,
                      double *dpred_out, void *stream);
/* L1/L2: Gaussian misfit, log-determinant and (tempered) log-likelihood ratio
 * (likelihood/_log_likelihood.py:204-251,308-332; likelihood/_target.py:136-208), uncorrelated
 * hierarchical noise: dpred [R][C], sigma [C] -> misfit [C], logdet [C]. */
int bbb_loglike(const double *dpred, const double *dobs, const double *sigma, int32_t R, int64_t C,
                double *misfit_out, double *logdet_out, void *stream);
/* L3: ParameterSpace.log_prior (parameterization/_parameter_space.py:189-195): values [P][K][C],
 * k [C] -> log_prior [C]. */
int bbb_logprior(const bbb_space_spec *space, const double *values, const int32_t *k, int64_t C,
                 double *log_prior_out, void *stream);
/* Philox4x32-10 known-answer check: fills out[4] (host) for counter[4], key[2]. */
int bbb_philox_kat(const uint32_t *counter, const uint32_t *key, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* BBB200_H */
