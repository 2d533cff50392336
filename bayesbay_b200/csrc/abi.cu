// extern "C" boundary of libbbb200.so (see include/bbb200.h).  Host-side only: validates the spec,
// keeps a device copy of (spec, buffers) and enqueues kernels on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "hosted.cuh"
#include "kernels.h"

using namespace bbb;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(int err, const char *what) {
    return fail(BBB_ERR_CUDA, "%s: CUDA error %d (%s)", what, err, cudaGetErrorString((cudaError_t)err));
}
#define CU(expr)                                               \
    do {                                                       \
        int _e = (int)(expr);                                  \
        if (_e != 0) return cuda_fail(_e, #expr);              \
    } while (0)

struct bbb_engine {
    EngineParams host;
    EngineParams *dev;
    int device;
    bool bound;
    bool has_ray;
    int n_sm;
    int k_bound[BBB_MAX_SPACES];  // host-side upper bound of max_c k[s][c] (grows by one per step, tightened by the caller)
    bool brute_force;  // BBB_BRUTE_FORCE_RASTER=1: re-rasterise every proposal from scratch (reference behaviour)
    bool full_forward; // BBB_FULL_FORWARD=1: never use the chain-local step (full SpMM every proposal, reference behaviour)
    int ray_t;         // chain-local engines: the one ray target
    long long resync_every, steps_since_resync;
    // chain-local engines, multi-iteration calls: the chains are split into parts (default 4, BBB_PIPELINE=n, 1 = off)
    // that run on internal streams, so that one part's proposal kernel (a latency-bound kernel of one warp per chain)
    // and the tail of its chain kernel overlap the other parts' chain kernels
    int pipeline;  // number of parts (1 = off)
    // programmatic dependent launch of the chain kernel behind the proposal kernel: 1 = single launches (default),
    // 2 = also the parts of pipelined multi-iteration calls, 0 = off (BBB_DEPENDENT_LAUNCH)
    int dependent_launch;
    int order_parity[4];  // per part: which set of block-order class counts the next proposal kernel counts into
    bool order_pending[4];  // a proposal kernel has filed its chains and the chain kernel has not consumed them yet
    cudaStream_t half_stream[4];
    cudaEvent_t ev_fork, ev_join[4];
    // multi-iteration calls are replayed from CUDA graphs of STEP_GRAPH_ITERS pipelined iterations (every launch of a
    // step kernel carries the engine parameters by value, a 4 KB argument: enqueued call by call the host side takes
    // longer than the device, 0.129 against 0.095 ms per iteration); one graph per block-order parity of the parts
    struct StepGraph {
        bool used;
        int parity;
        cudaGraph_t graph;
        cudaGraphExec_t exec;
    } step_graph[2][16];  // [16 iterations | 4 iterations][parity mask]
    int step_graphs;  // 1: on (default), 0: BBB_STEP_GRAPH=0 or capture failed
    cudaStream_t step_cap_stream;
    // chain-local engines, BBB_FUSED_RUN=1 (developer switch, off by default): bbb_engine_step runs the fused persistent
    // kernel (k_chain_run: proposal + chain step + decision of all iterations up to the next from-scratch re-evaluation
    // in ONE launch) instead of the per-iteration launches above.  Bit-identical results; measured SLOWER on B200
    // (0.109 against 0.083 ms per iteration back to back: the proposer warps' straight-line code and the workers'
    // loops thrash the SM's 32 KB instruction cache -- profiles/r2_fused_run_ncu.md).  run_ctl: the kernel's control
    // words (ticket counters, per-chain progress, ready queue).
    int fused_run;
    int32_t *run_ctl;
    // hosted step (bbb_engine_step_hosted / _wait): per-part "uploaded states differ" words on the device and their
    // pinned host copy, and what the pending call needs should a part have to be redone
    int32_t *hosted_flag_dev, *hosted_flag_host;
    // shape of one hosted call: the wire format, and where each part's image lives in the four buffers (word offsets)
    // and how many words of it travel each way
    struct HostedShape {
        int ragged;  // 0: rows of kb cells per chain (bbb_engine_step_hosted), 1: only the cells in use (.._ragged)
        int32_t kb_in[BBB_MAX_SPACES], kb_out[BBB_MAX_SPACES];
        long long in_off[4], in_words[4], out_off[4], out_words[4];
        bool same(const HostedShape &o, int n_spaces, int n_parts) const {
            if (ragged != o.ragged) return false;
            for (int s = 0; s < n_spaces && !ragged; ++s)
                if (kb_in[s] != o.kb_in[s] || kb_out[s] != o.kb_out[s]) return false;
            for (int h = 0; h < n_parts; ++h)
                if (in_off[h] != o.in_off[h] || in_words[h] != o.in_words[h] || out_off[h] != o.out_off[h] || out_words[h] != o.out_words[h])
                    return false;
            return true;
        }
    };
    struct {
        bool pending;
        HostedShape shape;
        void *host_out, *dev_stage, *dev_out;
    } hosted;
    // the hosted step is replayed from CUDA graphs (one launch instead of ~40 driver calls per iteration): captured on
    // `cap_stream` the first time a combination of image sizes, buffers and block-order parities is seen
    struct HostedGraph {
        bool used;
        HostedShape shape;
        const void *host_in;
        void *host_out, *dev_stage, *dev_out;
        int parity;
        cudaGraph_t graph;
        cudaGraphExec_t exec;
    } hosted_graph[16];
    int hosted_graph_next;
    int hosted_graphs;  // 1: replay from graphs (default), 0: enqueue every call (BBB_HOSTED_GRAPH=0, or capture failed)
    cudaStream_t cap_stream, copy_stream;
    cudaEvent_t ev_h2d[4], ev_hjoin[4], ev_hfork;
    // parts of the hosted step (BBB_HOSTED_PARTS, default 2) and their streams
    int hosted_np;
    cudaStream_t hosted_stream[4];
    // mapped hosted step (bbb_engine_step_hosted_mapped): per-chain mismatch flags and byte counters on the device, a
    // mapped pinned word the kernel raises on a mismatch, the event the wait blocks on (recorded right behind the chain
    // kernel: the proposal of the NEXT iteration and a periodic re-evaluation are enqueued behind it and overlap the
    // host's turn-around).  spec_valid: the proposal of the next iteration has been drawn from the resident states and
    // nothing has touched the engine since (every entry point clears it through NEED_BOUND).
    // mapped_flag_host words: [0] some chain mismatched, [1] a proposal warp gave up waiting for its chain's completion
    // signal (BBB_MAPPED_PDL=1 only).
    int32_t *mapped_mismatch, *mapped_flag_host, *mapped_chain_done;
    unsigned long long *mapped_moved;
    bool proposer_finished;  // the pending proposal kernel (run_propose) has decided the chains that need no forward
    int mapped_seq;
    int mapped_pdl;  // developer switch BBB_MAPPED_PDL (default 0, see bbb_engine_step_hosted_mapped)
    cudaEvent_t ev_mapped_done;
    bool spec_valid;
    struct {
        bool pending;
        const void *host_in;
        void *host_out;
    } mapped;
};

// the cached graphs hold the engine parameters BY VALUE: whatever changes them drops the graphs
// resources of the mapped hosted step (sized by the bound chain count: dropped when buffers are bound again)
static void mapped_drop(bbb_engine *e) {
    if (e->mapped_mismatch == nullptr) return;
    cudaFree(e->mapped_mismatch);
    cudaFree(e->mapped_moved);
    cudaFreeHost(e->mapped_flag_host);
    cudaFree(e->mapped_chain_done);
    cudaEventDestroy(e->ev_mapped_done);
    e->mapped_mismatch = nullptr;
    e->spec_valid = false;
}

static void hosted_graphs_drop(bbb_engine *e) {
    for (auto &row : e->step_graph)
        for (auto &g : row)
            if (g.used) {
                cudaGraphExecDestroy(g.exec);
                cudaGraphDestroy(g.graph);
                g.used = false;
            }
    for (auto &g : e->hosted_graph)
        if (g.used) {
            cudaGraphExecDestroy(g.exec);
            cudaGraphDestroy(g.graph);
            g.used = false;
        }
}

extern "C" {

int bbb_abi_version(void) { return BBB_ABI_VERSION; }
const char *bbb_last_error(void) { return g_err; }
int64_t bbb_ray_tiles(int64_t n_data) { return (n_data + RAY_TILE - 1) / RAY_TILE; }
int64_t bbb_idx_row_stride(int64_t n_points) { return (n_points + 15) & ~(int64_t)15; }
int64_t bbb_chain_local_tile_rays(int64_t n_data, int32_t kmax, int64_t n_points) {
    return chain_local_tile_rays(n_data, kmax, n_points);
}

static int validate_spec(const bbb_problem_spec *P) {
    if (P->n_spaces < 1 || P->n_spaces > BBB_MAX_SPACES) return fail(BBB_ERR_INVALID, "n_spaces out of range");
    if (P->n_targets < 0 || P->n_targets > BBB_MAX_TARGETS) return fail(BBB_ERR_INVALID, "n_targets out of range");
    for (int s = 0; s < P->n_spaces; ++s) {
        const bbb_space_spec &S = P->spaces[s];
        if (S.kind < BBB_SPACE_PLAIN || S.kind > BBB_SPACE_VORONOI2D) return fail(BBB_ERR_INVALID, "space %d: bad kind", s);
        const int want_ndim = S.kind == BBB_SPACE_PLAIN ? 0 : (S.kind == BBB_SPACE_VORONOI1D ? 1 : 2);
        if (S.ndim != want_ndim) return fail(BBB_ERR_INVALID, "space %d: ndim %d does not match kind", s, S.ndim);
        if (S.kmin < 1 || S.kmax < S.kmin) return fail(BBB_ERR_INVALID, "space %d: need 1 <= kmin <= kmax", s);
        if (S.kmax > 65535) return fail(BBB_ERR_UNSUPPORTED, "space %d: kmax > 65535", s);
        if (!S.trans_d && S.kmin != S.kmax) return fail(BBB_ERR_INVALID, "space %d: fixed dimension needs kmin == kmax", s);
        if (S.trans_d && (S.kinit_max < S.kmin || S.kinit_max > S.kmax))
            return fail(BBB_ERR_INVALID, "space %d: kinit_max outside [kmin, kmax]", s);
        if (S.n_params < 0 || S.n_params > BBB_MAX_PARAMS) return fail(BBB_ERR_INVALID, "space %d: n_params", s);
        if (S.w_birth < 0 || S.w_death < 0 || S.w_values < 0 || S.w_sites < 0 ||
            S.w_birth + S.w_death + S.w_values + S.w_sites <= 0)
            return fail(BBB_ERR_INVALID, "space %d: move weights", s);
        if (!S.trans_d && (S.w_birth > 0 || S.w_death > 0)) return fail(BBB_ERR_INVALID, "space %d: birth/death on a fixed space", s);
        if (S.kind == BBB_SPACE_PLAIN && S.w_sites > 0) return fail(BBB_ERR_INVALID, "space %d: site move on a plain space", s);
        if (S.n_params == 0 && S.w_values > 0) return fail(BBB_ERR_INVALID, "space %d: value move without parameters", s);
        for (int d = 0; d < S.ndim; ++d)
            if (!(S.vmin[d] < S.vmax[d]) || !(S.site_std[d] > 0)) return fail(BBB_ERR_INVALID, "space %d: site domain / perturb_std", s);
        for (int p = 0; p < S.n_params; ++p) {
            if (S.prior_kind[p] < BBB_PRIOR_UNIFORM || S.prior_kind[p] > BBB_PRIOR_LAPLACE)
                return fail(BBB_ERR_INVALID, "space %d param %d: prior kind", s, p);
            if (S.n_pos[p] > 0) {  // table validated by the caller (device memory)
                if (S.ndim != 1) return fail(BBB_ERR_UNSUPPORTED, "space %d param %d: position-dependent priors need a Voronoi1D space", s, p);
                if (S.n_pos[p] < 2 || !S.pos_table[p]) return fail(BBB_ERR_INVALID, "space %d param %d: position table", s, p);
                continue;
            }
            if (S.prior_kind[p] == BBB_PRIOR_UNIFORM ? !(S.prior_a[p] < S.prior_b[p]) : !(S.prior_b[p] > 0))
                return fail(BBB_ERR_INVALID, "space %d param %d: prior hyper-parameters", s, p);
            if (!(S.perturb_std[p] > 0) || !(S.perturb_std_birth[p] > 0))
                return fail(BBB_ERR_INVALID, "space %d param %d: perturb_std", s, p);
        }
    }
    for (int t = 0; t < P->n_targets; ++t) {
        const bbb_target_spec &T = P->targets[t];
        if (T.n_data < 1 || T.dobs == nullptr) return fail(BBB_ERR_INVALID, "target %d: no data", t);
        if (T.space < 0 || T.space >= P->n_spaces) return fail(BBB_ERR_INVALID, "target %d: space index", t);
        const bbb_space_spec &S = P->spaces[T.space];
        if (T.cov_kind == BBB_COV_HIERARCHICAL) {
            if (!(T.std_min >= 0) || !(T.std_max >= T.std_min) || !(T.std_perturb_std > 0))
                return fail(BBB_ERR_INVALID, "target %d: noise std range", t);
            if (T.correlated && (!(T.corr_min >= 0) || !(T.corr_max >= T.corr_min) || !(T.corr_max < 1.0 || T.corr_min < 1.0) ||
                                 !(T.corr_perturb_std > 0) || T.n_data < 3))
                return fail(BBB_ERR_INVALID, "target %d: noise correlation range (needs 0 <= min <= max, at least 3 data)", t);
        } else if (T.correlated) {
            return fail(BBB_ERR_INVALID, "target %d: correlated noise needs a hierarchical target", t);
        } else if (T.cov_kind == BBB_COV_VECTOR || T.cov_kind == BBB_COV_MATRIX) {
            if (T.cov_vec == nullptr) return fail(BBB_ERR_INVALID, "target %d: cov_vec is NULL", t);
        } else if (T.cov_kind != BBB_COV_SCALAR) {
            return fail(BBB_ERR_INVALID, "target %d: cov_kind", t);
        }
        if (T.fwd_kind == BBB_FWD_RAY2D) {
            if (S.kind != BBB_SPACE_VORONOI2D) return fail(BBB_ERR_INVALID, "target %d: ray2d forward needs a Voronoi2D space", t);
            if (T.param < 0 || T.param >= S.n_params) return fail(BBB_ERR_INVALID, "target %d: param index", t);
            if (T.n_points < 1 || !T.indptr || !T.indices || !T.data || !T.points_x || !T.points_y)
                return fail(BBB_ERR_INVALID, "target %d: ray2d arrays", t);
            if (T.csc_tile_rays != 0) {
                if (T.csc_tile_rays != chain_local_tile_rays(T.n_data, S.kmax, T.n_points))
                    return fail(BBB_ERR_INVALID, "target %d: csc_tile_rays must be bbb_chain_local_tile_rays()", t);
                if (T.csc_n_tiles != (T.n_data + T.csc_tile_rays - 1) / T.csc_tile_rays || T.csc_n_levels < 1 ||
                    T.csc_n_levels > BBB_CSC_LEVELS || !(T.csc_bound >= 0))
                    return fail(BBB_ERR_INVALID, "target %d: CSC tiles / levels", t);
                for (int l = 0; l < T.csc_n_levels; ++l)
                    if (!T.csc_colptr[l] || !T.csc_rows[l] || !T.csc_data[l]) return fail(BBB_ERR_INVALID, "target %d: CSC arrays of level %d", t, l);
            }
        } else if (T.fwd_kind == BBB_FWD_LOOKUP1D) {
            if (S.kind != BBB_SPACE_VORONOI1D) return fail(BBB_ERR_INVALID, "target %d: lookup1d forward needs a Voronoi1D space", t);
            if (T.param < 0 || T.param >= S.n_params || !T.x) return fail(BBB_ERR_INVALID, "target %d: lookup1d arrays", t);
        } else if (T.fwd_kind == BBB_FWD_POLYNOMIAL) {
            if (S.kind != BBB_SPACE_PLAIN) return fail(BBB_ERR_INVALID, "target %d: polynomial forward needs a plain ParameterSpace", t);
            if (T.param < 0 || T.param >= S.n_params || !T.x) return fail(BBB_ERR_INVALID, "target %d: polynomial arrays", t);
        } else if (T.fwd_kind == BBB_FWD_LINEAR) {
            if (S.trans_d) return fail(BBB_ERR_INVALID, "target %d: linear forward needs a fixed-dimension space", t);
            if (T.n_lin_params < 1 || T.n_lin_params > BBB_MAX_LINEAR_PARAMS || !T.G)
                return fail(BBB_ERR_INVALID, "target %d: linear arrays", t);
            for (int q = 0; q < T.n_lin_params; ++q)
                if (T.lin_params[q] < 0 || T.lin_params[q] >= S.n_params) return fail(BBB_ERR_INVALID, "target %d: lin_params", t);
        } else {
            return fail(BBB_ERR_INVALID, "target %d: fwd_kind", t);
        }
    }
    return BBB_OK;
}

int bbb_engine_create(const bbb_problem_spec *spec, int device, bbb_engine **out) {
    if (!spec || !out) return fail(BBB_ERR_INVALID, "NULL argument");
    int rc = validate_spec(spec);
    if (rc != BBB_OK) return rc;
    CU(cudaSetDevice(device));
    bbb_engine *e = new (std::nothrow) bbb_engine();
    if (!e) return fail(BBB_ERR_INVALID, "out of host memory");
    memset(&e->host, 0, sizeof(e->host));
    e->host.spec = *spec;
    e->device = device;
    e->bound = false;
    e->has_ray = false;
    {
        const char *bf = getenv("BBB_BRUTE_FORCE_RASTER");
        e->brute_force = bf != nullptr && bf[0] == '1';
        const char *ff = getenv("BBB_FULL_FORWARD");
        e->full_forward = ff != nullptr && ff[0] == '1';
    }
    e->ray_t = -1;
    e->resync_every = 256;
    e->steps_since_resync = 0;
    {
        const char *pl = getenv("BBB_PIPELINE");
        e->pipeline = pl != nullptr ? atoi(pl) : 4;
        if (e->pipeline < 1) e->pipeline = 1;
        if (e->pipeline > 4) e->pipeline = 4;
        const char *hp = getenv("BBB_HOSTED_PARTS");
        e->hosted_np = hp != nullptr ? atoi(hp) : (e->pipeline < 2 ? e->pipeline : 2);
        if (e->hosted_np < 1) e->hosted_np = 1;
        if (e->hosted_np > 4) e->hosted_np = 4;
    }
    {
        const char *dl = getenv("BBB_DEPENDENT_LAUNCH");
        e->dependent_launch = dl != nullptr ? atoi(dl) : 1;
    }
    for (int h = 0; h < 4; ++h) e->half_stream[h] = nullptr;
    {
        const char *sg = getenv("BBB_STEP_GRAPH");
        e->step_graphs = (sg != nullptr && atoi(sg) == 0) ? 0 : 1;
        e->step_cap_stream = nullptr;
    }
    {
        const char *fr = getenv("BBB_FUSED_RUN");
        e->fused_run = (fr != nullptr && atoi(fr) == 1) ? 1 : 0;
        e->run_ctl = nullptr;
    }
    for (int t = 0; t < spec->n_targets; ++t) e->has_ray |= (spec->targets[t].fwd_kind == BBB_FWD_RAY2D);
    for (int s = 0; s < BBB_MAX_SPACES; ++s) e->k_bound[s] = s < spec->n_spaces ? spec->spaces[s].kmax : 0;
    e->n_sm = 148;
    cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device);
    int err = (int)cudaMalloc(&e->dev, sizeof(EngineParams));
    if (err != 0) {
        delete e;
        return cuda_fail(err, "cudaMalloc(EngineParams)");
    }
    *out = e;
    return BBB_OK;
}

int bbb_engine_bind_buffers(bbb_engine *e, const bbb_buffers *b) {
    if (!e || !b) return fail(BBB_ERR_INVALID, "NULL argument");
    const bbb_problem_spec &P = e->host.spec;
    if (b->n_chains < 1) return fail(BBB_ERR_SHAPE, "n_chains < 1");
    for (int s = 0; s < P.n_spaces; ++s) {
        if (!b->k[s] || !b->sel[s]) return fail(BBB_ERR_SHAPE, "space %d: k/sel buffer missing", s);
        if (P.spaces[s].ndim > 0 && !b->sites[s]) return fail(BBB_ERR_SHAPE, "space %d: sites buffer missing", s);
        if (P.spaces[s].n_params > 0 && !b->values[s]) return fail(BBB_ERR_SHAPE, "space %d: values buffer missing", s);
    }
    for (int t = 0; t < P.n_targets; ++t) {
        if (!b->phi[t]) return fail(BBB_ERR_SHAPE, "target %d: phi buffer missing", t);
        if (P.targets[t].cov_kind == BBB_COV_HIERARCHICAL && !b->sigma[t]) return fail(BBB_ERR_SHAPE, "target %d: sigma buffer missing", t);
        if (P.targets[t].correlated && !b->corr[t]) return fail(BBB_ERR_SHAPE, "target %d: corr buffer missing", t);
        if (((P.targets[t].correlated && P.targets[t].fwd_kind == BBB_FWD_RAY2D) || P.targets[t].cov_kind == BBB_COV_MATRIX) &&
            !b->dpred_scratch[t])
            return fail(BBB_ERR_SHAPE, "target %d: dpred_scratch buffer missing", t);
        if (P.targets[t].fwd_kind == BBB_FWD_RAY2D &&
            (!b->cell_idx[t] || !b->idx_sel[t] || !b->partial[t] || !b->rescan_list[t] || !b->rescan_count[t]))
            return fail(BBB_ERR_SHAPE, "target %d: cell_idx/idx_sel/partial/rescan buffer missing", t);
    }
    if (!b->temperature || !b->iteration || !b->move || !b->edit || !b->edit_site || !b->edit_vals || !b->log_prob_ratio || !b->log_like_ratio ||
        !b->accepted || !b->n_proposed || !b->n_accepted || !b->tape_cursor)
        return fail(BBB_ERR_SHAPE, "a per-chain scalar buffer is missing");
    if (b->tape && b->tape_stride < 1) return fail(BBB_ERR_SHAPE, "tape_stride < 1");
    e->host.buf = *b;
    hosted_graphs_drop(e);
    mapped_drop(e);
    for (int t = 0; t < P.n_targets; ++t) {
        e->host.n_partial[t] = 0;
        if (P.targets[t].fwd_kind != BBB_FWD_RAY2D) continue;
        const int K = P.spaces[P.targets[t].space].kmax;
        e->host.n_partial[t] = spmm_ray_groups(P.targets[t].n_data, b->n_chains, K, K <= 256 ? 1 : 2, e->n_sm);
        if (e->host.n_partial[t] == 0)
            return fail(BBB_ERR_UNSUPPORTED, "target %d: n_dimensions_max %d is too large for the ray kernel's shared-memory value table", t, K);
    }
    // chain-local step: exactly one ray target, CSC copy of its matrix and the chain-major caches provided
    e->host.chain_local = 0;
    e->ray_t = -1;
    e->host.ray_t = -1;
    for (int h = 0; h < 4; ++h) { e->order_parity[h] = 0; e->order_pending[h] = false; }
    {
        int n_ray = 0, rt = -1;
        for (int t = 0; t < P.n_targets; ++t)
            if (P.targets[t].fwd_kind == BBB_FWD_RAY2D) { ++n_ray; rt = t; }
        if (n_ray == 1 && !e->full_forward && !e->brute_force && P.targets[rt].csc_tile_rays > 0 && !reduces_dpred(P.targets[rt])) {
            const bool have = b->idx_cm[rt] && b->res[rt] && b->change_list[rt] && b->dpred_scratch[rt];
            if (!have && (b->idx_cm[rt] || b->res[rt] || b->change_list[rt]))
                return fail(BBB_ERR_SHAPE, "target %d: the chain-local step needs idx_cm, res, change_list and dpred_scratch", rt);
            if (have) {
                e->host.chain_local = 1;
                e->ray_t = rt;
                e->host.ray_t = rt;
            }
        }
    }
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpy(e->dev, &e->host, sizeof(EngineParams), cudaMemcpyHostToDevice));
    if (e->run_ctl) { cudaFree(e->run_ctl); e->run_ctl = nullptr; }
    if (e->host.chain_local) {
        CU(cudaMalloc(&e->run_ctl, sizeof(int32_t) * (size_t)chain_run_ctl_words(b->n_chains)));
        CU(cudaMemset(e->run_ctl, 0, sizeof(int32_t) * (size_t)chain_run_ctl_words(b->n_chains)));
    }
    e->bound = true;
    return BBB_OK;
}

void bbb_engine_destroy(bbb_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->half_stream[0] != nullptr) {
        for (int h = 0; h < e->pipeline; ++h) {
            cudaStreamDestroy(e->half_stream[h]);
            cudaEventDestroy(e->ev_join[h]);
        }
        cudaEventDestroy(e->ev_fork);
    }
    hosted_graphs_drop(e);
    if (e->step_cap_stream) cudaStreamDestroy(e->step_cap_stream);
    if (e->cap_stream) {
        cudaStreamDestroy(e->cap_stream);
        cudaStreamDestroy(e->copy_stream);
        for (auto &ev : e->ev_h2d) cudaEventDestroy(ev);
        for (auto &ev : e->ev_hjoin) cudaEventDestroy(ev);
        cudaEventDestroy(e->ev_hfork);
        for (auto &st : e->hosted_stream) cudaStreamDestroy(st);
    }
    if (e->run_ctl) cudaFree(e->run_ctl);
    if (e->hosted_flag_dev) cudaFree(e->hosted_flag_dev);
    if (e->hosted_flag_host) cudaFreeHost(e->hosted_flag_host);
    mapped_drop(e);
    cudaFree(e->dev);
    delete e;
}

static int idx_bytes_of(const bbb_problem_spec &P, int t) { return P.spaces[P.targets[t].space].kmax <= 256 ? 1 : 2; }

static int run_raster(bbb_engine *e, int t, bool proposal, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const bbb_buffers &B = e->host.buf;
    const bbb_target_spec &T = P.targets[t];
    const bbb_space_spec &S = P.spaces[T.space];
    RasterArgs a;
    a.sites = B.sites[T.space];
    a.k = B.k[T.space];
    a.K = S.kmax;
    a.k_rows = proposal ? (e->k_bound[T.space] + 1 < S.kmax ? e->k_bound[T.space] + 1 : S.kmax) : S.kmax;
    a.C = B.n_chains;
    a.sel = B.sel[T.space];
    a.idx_sel = B.idx_sel[t];
    a.slot_xor = proposal ? 1 : 0;
    a.move = proposal ? B.move : nullptr;
    a.lpr = B.log_prob_ratio;
    a.edit = B.edit;
    a.edit_site = B.edit_site;
    a.n_spaces = P.n_spaces;
    a.space = T.space;
    a.px = T.points_x;
    a.py = T.points_y;
    a.G = T.n_points;
    a.idx = B.cell_idx[t];
    a.idx_bytes = idx_bytes_of(P, t);
    a.points_per_block = raster_points_per_block(a.G, a.C);
    a.rescan_list = B.rescan_list[t];
    a.rescan_count = B.rescan_count[t];
    if (!proposal) {
        // current states: from-scratch rasterisation into slot 0, nothing pending afterwards
        int rc = (int)cudaMemsetAsync(B.idx_sel[t], 0, (size_t)B.n_chains, st);
        return rc != 0 ? rc : launch_raster2d(a, st);
    }
    // commit pending accepted geometries, proposed slot := current slot, then update the proposing chains:
    // change-locally, or from scratch when BBB_BRUTE_FORCE_RASTER=1 (the reference's behaviour)
    int rc = launch_idx_commit_copy(B.cell_idx[t], a.idx_bytes, B.idx_sel[t], a.G, a.C, st);
    if (rc != 0) return rc;
    return e->brute_force ? launch_raster2d(a, st) : launch_raster2d_inc(a, st);
}

// proposal + parallel write of the proposed slots
// (decide_no_forward: resident stepping of a chain-local engine -- chains whose proposal needs no forward are decided by
// the proposal kernel and skipped by the chain kernel; never for a proposal drawn ahead of a hosted step)
static int run_propose(bbb_engine *e, cudaStream_t st, bool decide_no_forward = false) {
    const long long C = e->host.buf.n_chains;
    e->order_pending[0] = true;
    e->proposer_finished = decide_no_forward && e->host.chain_local;
    return launch_propose(e->host, 0, C, e->order_parity[0] * 4, nullptr, st, nullptr, 0, nullptr, e->proposer_finished);
}

static int run_spmm(bbb_engine *e, int t, bool proposal, double *dpred, bool write_partial, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const bbb_buffers &B = e->host.buf;
    const bbb_target_spec &T = P.targets[t];
    const bbb_space_spec &S = P.spaces[T.space];
    SpmmArgs a;
    a.idx = B.cell_idx[t];
    a.idx_bytes = idx_bytes_of(P, t);
    a.idx_sel = B.idx_sel[t];
    a.values = B.values[T.space];
    a.k = B.k[T.space];
    a.sel = B.sel[T.space];
    a.slot_xor = proposal ? 1 : 0;
    a.n_params = S.n_params;
    a.param = T.param;
    a.K = S.kmax;
    a.C = B.n_chains;
    a.G = T.n_points;
    a.R = T.n_data;
    a.indptr = T.indptr;
    a.indices = T.indices;
    a.data = T.data;
    a.transform = T.transform;
    a.dobs = T.dobs;
    a.w = (T.cov_kind == BBB_COV_VECTOR) ? T.cov_vec : nullptr;
    a.move = proposal ? B.move : nullptr;
    a.lpr = B.log_prob_ratio;
    a.n_spaces = P.n_spaces;
    a.space = T.space;
    a.dpred = dpred;
    if (reduces_dpred(T) && write_partial && dpred == nullptr) a.dpred = B.dpred_scratch[t];  // reduced by k_corr_sums
    a.partial = write_partial ? B.partial[t] : nullptr;
    // the value table holds k_rows cells per chain; a tighter bound allows more chains per lane
    a.k_rows = proposal ? (e->k_bound[T.space] + 1 < S.kmax ? e->k_bound[T.space] + 1 : S.kmax) : e->k_bound[T.space];
    a.n_groups = spmm_ray_groups(a.R, a.C, a.k_rows, a.idx_bytes, e->n_sm);
    if (a.n_groups == 0) return (int)cudaErrorInvalidConfiguration;
    if (write_partial && a.n_groups != e->host.n_partial[t]) {
        // the accept stage sums partial[0 .. n_partial): keep the device copy in step (stream ordered)
        e->host.n_partial[t] = a.n_groups;
        // (captured launches hold the parameters by value; the kernels of a chain-local engine's graphs -- proposal,
        // chain step, state images -- never read n_partial, so its graphs stay)
        if (!e->host.chain_local) hosted_graphs_drop(e);
        int rc = (int)cudaMemcpyAsync(&e->dev->n_partial[t], &e->host.n_partial[t], sizeof(int), cudaMemcpyHostToDevice, st);
        if (rc != 0) return rc;
    }
    int rc = launch_ray_spmm(a, st);
    if (rc == 0 && reduces_dpred(T) && write_partial) rc = launch_corr_sums(e->dev, a.C, t, proposal ? 1 : 0, st);
    return rc;
}

#define NEED_BOUND(e)                                                                  \
    do {                                                                               \
        if (!(e)) return fail(BBB_ERR_INVALID, "NULL engine");                         \
        if (!(e)->bound) return fail(BBB_ERR_SHAPE, "bbb_engine_bind_buffers not called"); \
        (e)->spec_valid = false;                                                       \
        CU(cudaSetDevice((e)->device));                                                \
    } while (0)

// Chain-local engines: bring the [G][C] index slot 0 in line with the chain-major index (the SpMM and the
// result kernels read the former).
static int sync_index_slot0(bbb_engine *e, cudaStream_t st) {
    const int t = e->ray_t;
    const bbb_buffers &B = e->host.buf;
    const long long G = e->host.spec.targets[t].n_points;
    int rc = launch_transpose_idx(B.idx_cm[t], B.cell_idx[t], idx_bytes_of(e->host.spec, t), B.n_chains, G,
                                  bbb_idx_row_stride(G), B.n_chains, st);
    return rc != 0 ? rc : (int)cudaMemsetAsync(B.idx_sel[t], 0, (size_t)B.n_chains, st);
}

// Chain-local engines: forward and misfit of the CURRENT states from scratch (full SpMM) into res / phi;
// `reraster` also re-derives the nearest-site index from the sites.
static int chain_local_from_scratch(bbb_engine *e, bool reraster, cudaStream_t st) {
    const int t = e->ray_t;
    const bbb_buffers &B = e->host.buf;
    const bbb_target_spec &T = e->host.spec.targets[t];
    const long long G = T.n_points, C = B.n_chains;
    int rc = reraster ? run_raster(e, t, false, st) : sync_index_slot0(e, st);
    if (rc == 0) rc = run_spmm(e, t, false, B.dpred_scratch[t], true, st);
    if (rc == 0) rc = launch_refresh_finish(e->dev, C, st);
    if (rc == 0 && reraster)
        rc = launch_transpose_idx(B.cell_idx[t], B.idx_cm[t], idx_bytes_of(e->host.spec, t), G, C, C, bbb_idx_row_stride(G), st);
    if (rc == 0) rc = launch_residual_transpose(B.dpred_scratch[t], B.res[t], T.n_data, C, T.dobs, st);
    e->steps_since_resync = 0;
    return rc;
}

int bbb_engine_refresh(bbb_engine *e, void *stream) {
    NEED_BOUND(e);
    cudaStream_t st = (cudaStream_t)stream;
    const bbb_problem_spec &P = e->host.spec;
    for (int s = 0; s < P.n_spaces; ++s) e->k_bound[s] = P.spaces[s].kmax;  // caller-written states: no knowledge
    if (e->host.chain_local) {
        CU(chain_local_from_scratch(e, true, st));
        return BBB_OK;
    }
    for (int t = 0; t < P.n_targets; ++t) {
        if (P.targets[t].fwd_kind != BBB_FWD_RAY2D) continue;
        CU(run_raster(e, t, false, st));
        CU(run_spmm(e, t, false, nullptr, true, st));
    }
    CU(launch_refresh_finish(e->dev, e->host.buf.n_chains, st));
    return BBB_OK;
}

int bbb_engine_chain_local(bbb_engine *e) { return (e && e->bound && e->host.chain_local) ? 1 : 0; }

int bbb_engine_set_resync(bbb_engine *e, int64_t every, int64_t steps_since) {
    if (!e) return fail(BBB_ERR_INVALID, "NULL engine");
    if (every < 0 || steps_since < 0) return fail(BBB_ERR_INVALID, "negative resync schedule");
    e->resync_every = every;
    e->steps_since_resync = steps_since;
    return BBB_OK;
}
int64_t bbb_engine_steps_since_resync(bbb_engine *e) { return e ? e->steps_since_resync : 0; }

int bbb_engine_resync(bbb_engine *e, void *stream) {
    NEED_BOUND(e);
    if (e->host.chain_local) CU(chain_local_from_scratch(e, false, (cudaStream_t)stream));
    return BBB_OK;
}

// a step adds at most one cell: keep the host-side upper bound of k valid (BEFORE anything sized by it runs)
static void bump_k_bound(bbb_engine *e, long long n) {
    const bbb_problem_spec &P = e->host.spec;
    for (int s = 0; s < P.n_spaces; ++s) {
        const long long b = (long long)e->k_bound[s] + n;
        e->k_bound[s] = b < P.spaces[s].kmax ? (int)b : P.spaces[s].kmax;
    }
}

static int run_chain_step(bbb_engine *e, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    int rc = launch_chain_step(e->host, 0, e->host.buf.n_chains, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax,
                               idx_bytes_of(P, e->ray_t), e->order_pending[0] ? e->order_parity[0] * 4 : -1, e->dependent_launch >= 1, nullptr, st,
                               nullptr, e->proposer_finished);
    e->proposer_finished = false;
    if (rc != 0) return rc;
    if (e->order_pending[0]) { e->order_parity[0] ^= 1; e->order_pending[0] = false; }
    bump_k_bound(e, 1);
    ++e->steps_since_resync;
    if (e->resync_every > 0 && e->steps_since_resync >= e->resync_every) rc = chain_local_from_scratch(e, false, st);
    return rc;
}

// parts of the chain set (whole proposal blocks) and their internal streams
static long long hosted_part_chains(const bbb_engine *e) {
    const long long C = e->host.buf.n_chains;
    return ((C + e->hosted_np - 1) / e->hosted_np + 3) & ~3LL;
}
static long long part_chains(const bbb_engine *e) {
    const long long C = e->host.buf.n_chains;
    return ((C + e->pipeline - 1) / e->pipeline + 3) & ~3LL;
}
static int part_streams(bbb_engine *e) {
    if (e->half_stream[0] != nullptr) return 0;
    for (int h = 0; h < e->pipeline; ++h) {
        int rc = (int)cudaStreamCreateWithFlags(&e->half_stream[h], cudaStreamNonBlocking);
        if (rc == 0) rc = (int)cudaEventCreateWithFlags(&e->ev_join[h], cudaEventDisableTiming);
        if (rc != 0) return rc;
    }
    return (int)cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
}

// n chain-local iterations of all chains, the parts of the chain set on the engine's internal streams (chains never
// interact inside a step, so the parts need no ordering between them); joined back into `st`.
static int run_chain_steps_direct(bbb_engine *e, long long n, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    const long long C = e->host.buf.n_chains;
    const int NP = e->pipeline;
    const long long part = part_chains(e);
    int rc = part_streams(e);
    if (rc != 0) return rc;
    rc = (int)cudaEventRecord(e->ev_fork, st);
    for (int h = 0; h < NP && rc == 0; ++h) rc = (int)cudaStreamWaitEvent(e->half_stream[h], e->ev_fork, 0);
    for (long long it = 0; it < n && rc == 0; ++it) {
        for (int h = 0; h < NP && rc == 0; ++h) {
            const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
            if (c0 >= c1) continue;
            rc = launch_propose(e->host, c0, c1, e->order_parity[h] * 4 + h, nullptr, e->half_stream[h], nullptr, 0, nullptr, true);
            if (rc == 0)
                rc = launch_chain_step(e->host, c0, c1, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax,
                                       idx_bytes_of(P, e->ray_t), e->order_parity[h] * 4 + h, e->dependent_launch >= 2, nullptr, e->half_stream[h],
                                       nullptr, true);
            e->order_parity[h] ^= 1;
        }
    }
    for (int h = 0; h < NP && rc == 0; ++h) {
        rc = (int)cudaEventRecord(e->ev_join[h], e->half_stream[h]);
        if (rc == 0) rc = (int)cudaStreamWaitEvent(st, e->ev_join[h], 0);
    }
    return rc;
}

// graphs of 16 and of 4 iterations (even counts: the parts' block-order parities are back where they started)
constexpr int STEP_GRAPH_ITERS[2] = {16, 4};

static int run_chain_steps_pipelined(bbb_engine *e, long long n, cudaStream_t st) {
    int rc = part_streams(e);
    if (rc != 0) return rc;
    while (n >= STEP_GRAPH_ITERS[1] && e->step_graphs) {
        const int size = n >= STEP_GRAPH_ITERS[0] ? 0 : 1;
        const int iters = STEP_GRAPH_ITERS[size];
        int parity = 0;
        for (int h = 0; h < 4; ++h) parity |= e->order_parity[h] << h;
        bbb_engine::StepGraph &g = e->step_graph[size][parity];
        if (!g.used) {
            if (e->step_cap_stream == nullptr) {
                rc = (int)cudaStreamCreateWithFlags(&e->step_cap_stream, cudaStreamNonBlocking);
                if (rc != 0) return rc;
            }
            rc = (int)cudaStreamBeginCapture(e->step_cap_stream, cudaStreamCaptureModeRelaxed);
            if (rc == 0) {
                rc = run_chain_steps_direct(e, iters, e->step_cap_stream);
                cudaGraph_t graph = nullptr;
                const int rc_end = (int)cudaStreamEndCapture(e->step_cap_stream, &graph);
                if (rc == 0) rc = rc_end;
                if (rc == 0) rc = (int)cudaGraphInstantiate(&g.exec, graph, 0);
                if (rc == 0) {
                    g.graph = graph;
                    g.parity = parity;
                    g.used = true;
                } else if (graph != nullptr) {
                    cudaGraphDestroy(graph);
                }
            }
            if (!g.used) {  // capture is not available here: enqueue every launch from now on
                (void)cudaGetLastError();
                e->step_graphs = 0;
                break;
            }
        }
        rc = (int)cudaGraphLaunch(g.exec, st);
        if (rc != 0) return rc;
        n -= iters;
    }
    return n > 0 ? run_chain_steps_direct(e, n, st) : 0;
}

int bbb_engine_init_states(bbb_engine *e, void *stream) {
    NEED_BOUND(e);
    CU(launch_init_states(e->dev, e->host.buf.n_chains, (cudaStream_t)stream));
    int rc = bbb_engine_refresh(e, stream);
    for (int s = 0; s < e->host.spec.n_spaces; ++s) {
        const bbb_space_spec &S = e->host.spec.spaces[s];
        e->k_bound[s] = S.trans_d ? S.kinit_max : S.kmin;  // the initialisation kernel draws k <= kinit_max
    }
    return rc;
}

// true maximum of k over the resident chains of every space (device reduction, blocking read-back)
static int measure_k_max(bbb_engine *e, int32_t *out, cudaStream_t st) {
    int32_t *dev = nullptr;
    CU(cudaMalloc(&dev, BBB_MAX_SPACES * sizeof(int32_t)));
    int rc = (int)cudaMemsetAsync(dev, 0, BBB_MAX_SPACES * sizeof(int32_t), st);
    if (rc == 0) rc = launch_max_k(e->dev, e->host.buf.n_chains, dev, st);
    if (rc == 0) rc = (int)cudaMemcpyAsync(out, dev, BBB_MAX_SPACES * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (rc == 0) rc = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    if (rc != 0) return fail(BBB_ERR_CUDA, "measuring the largest k: %s", cudaGetErrorString((cudaError_t)rc));
    return BBB_OK;
}

int bbb_engine_set_k_bound(bbb_engine *e, int space, int32_t bound) {
    if (!e) return fail(BBB_ERR_INVALID, "NULL engine");
    if (space < 0 || space >= e->host.spec.n_spaces) return fail(BBB_ERR_INVALID, "space index");
    const bbb_space_spec &S = e->host.spec.spaces[space];
    if (bound < S.kmin || bound > S.kmax) return fail(BBB_ERR_INVALID, "bound outside [kmin, kmax]");
    if (e->bound) {  // a bound below the true maximum would let the ray kernel index outside its value table: refused
        CU(cudaSetDevice(e->device));
        int32_t kmax_now[BBB_MAX_SPACES] = {0};
        if (int rc = measure_k_max(e, kmax_now, 0)) return rc;
        if (bound < kmax_now[space])
            return fail(BBB_ERR_INVALID, "space %d: bound %d is below the largest k of the resident chains (%d)", space, bound, kmax_now[space]);
    }
    e->k_bound[space] = bound;
    return BBB_OK;
}

int bbb_engine_sync_k_bound(bbb_engine *e, int32_t *bounds_out, void *stream) {
    if (!e) return fail(BBB_ERR_INVALID, "NULL engine");
    if (!e->bound) return fail(BBB_ERR_SHAPE, "bbb_engine_bind_buffers not called");
    CU(cudaSetDevice(e->device));
    int32_t kmax_now[BBB_MAX_SPACES] = {0};
    if (int rc = measure_k_max(e, kmax_now, (cudaStream_t)stream)) return rc;
    for (int s = 0; s < e->host.spec.n_spaces; ++s) {
        const bbb_space_spec &S = e->host.spec.spaces[s];
        int32_t b = kmax_now[s] < S.kmin ? S.kmin : (kmax_now[s] > S.kmax ? S.kmax : kmax_now[s]);
        if (S.trans_d) e->k_bound[s] = b;
        if (bounds_out) bounds_out[s] = e->k_bound[s];
    }
    return BBB_OK;
}

int bbb_engine_step(bbb_engine *e, int64_t n_steps, void *stream) {
    NEED_BOUND(e);
    if (n_steps < 0) return fail(BBB_ERR_INVALID, "n_steps < 0");
    cudaStream_t st = (cudaStream_t)stream;
    const bbb_problem_spec &P = e->host.spec;
    const long long C = e->host.buf.n_chains;
    if (!e->has_ray) {
        if (n_steps > 0) CU(launch_fused_steps(e->host, C, n_steps, st));
        return BBB_OK;
    }
    if (e->host.chain_local && e->fused_run && n_steps > 0) {
        // segments between two from-scratch re-evaluations: ONE persistent launch each
        const bbb_target_spec &T = P.targets[e->ray_t];
        int64_t left = n_steps;
        while (left > 0) {
            int64_t seg = left;
            if (e->resync_every > 0 && e->resync_every - e->steps_since_resync < seg) seg = e->resync_every - e->steps_since_resync;
            if (seg < 1) seg = 1;
            if (seg * C > (1LL << 30)) seg = (1LL << 30) / C > 0 ? (1LL << 30) / C : 1;
            CU(launch_chain_run(e->host, 0, C, seg, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax, idx_bytes_of(P, e->ray_t),
                                e->run_ctl, 2 * e->n_sm, st));
            bump_k_bound(e, seg);
            e->steps_since_resync += seg;
            if (e->resync_every > 0 && e->steps_since_resync >= e->resync_every) CU(chain_local_from_scratch(e, false, st));
            left -= seg;
        }
        return BBB_OK;
    }
    if (e->host.chain_local && e->pipeline > 1 && n_steps > 1 && C >= 2LL * e->n_sm) {
        // segments between two from-scratch re-evaluations run pipelined
        int64_t left = n_steps;
        while (left > 0) {
            int64_t seg = left;
            if (e->resync_every > 0 && e->resync_every - e->steps_since_resync < seg) seg = e->resync_every - e->steps_since_resync;
            if (seg < 1) seg = 1;
            CU(run_chain_steps_pipelined(e, seg, st));
            bump_k_bound(e, seg);
            e->steps_since_resync += seg;
            if (e->resync_every > 0 && e->steps_since_resync >= e->resync_every) CU(chain_local_from_scratch(e, false, st));
            left -= seg;
        }
        return BBB_OK;
    }
    for (int64_t it = 0; it < n_steps; ++it) {
        CU(run_propose(e, st, true));
        if (e->host.chain_local) {
            CU(run_chain_step(e, st));
        } else {
            for (int t = 0; t < P.n_targets; ++t) {
                if (P.targets[t].fwd_kind != BBB_FWD_RAY2D) continue;
                CU(run_raster(e, t, true, st));
                CU(run_spmm(e, t, true, nullptr, true, st));
            }
            CU(launch_finish(e->host, C, st));
            bump_k_bound(e, 1);
        }
    }
    return BBB_OK;
}

int bbb_engine_stage(bbb_engine *e, int stage, void *stream) {
    NEED_BOUND(e);
    cudaStream_t st = (cudaStream_t)stream;
    const bbb_problem_spec &P = e->host.spec;
    const long long C = e->host.buf.n_chains;
    if (!e->has_ray) return fail(BBB_ERR_UNSUPPORTED, "staged stepping needs a ray target (other problems run the fused kernel)");
    switch (stage) {
        case 0: CU(run_propose(e, st, true)); break;
        case 1:
            if (e->host.chain_local) break;
            for (int t = 0; t < P.n_targets; ++t)
                if (P.targets[t].fwd_kind == BBB_FWD_RAY2D) CU(run_raster(e, t, true, st));
            break;
        case 2:
            if (e->host.chain_local) {
                CU(run_chain_step(e, st));
                break;
            }
            for (int t = 0; t < P.n_targets; ++t)
                if (P.targets[t].fwd_kind == BBB_FWD_RAY2D) CU(run_spmm(e, t, true, nullptr, true, st));
            break;
        case 3:
            if (!e->host.chain_local) {
                CU(launch_finish(e->host, C, st));
                bump_k_bound(e, 1);
            }
            break;
        default: return fail(BBB_ERR_INVALID, "stage must be 0..3");
    }
    return BBB_OK;
}

int bbb_engine_set_annealing(bbb_engine *e, double temperature_start, double cooling_rate, int64_t cooling_iterations,
                             int64_t burnin_iterations) {
    NEED_BOUND(e);
    if (temperature_start < 0 || cooling_iterations < 0 || burnin_iterations < 0) return fail(BBB_ERR_INVALID, "negative annealing parameter");
    hosted_graphs_drop(e);
    e->host.anneal.enabled = temperature_start > 0 ? 1 : 0;
    e->host.anneal.t0 = temperature_start;
    e->host.anneal.rate = cooling_rate;
    e->host.anneal.cooling_iterations = cooling_iterations;
    e->host.anneal.burnin_iterations = burnin_iterations;
    CU(cudaMemcpyAsync(&e->dev->anneal, &e->host.anneal, sizeof(Annealing), cudaMemcpyHostToDevice, 0));
    CU(cudaStreamSynchronize(0));
    return BBB_OK;
}

int bbb_engine_gather_space(bbb_engine *e, int space, int32_t *k_out, double *sites_out, double *values_out,
                            void *stream) {
    NEED_BOUND(e);
    if (space < 0 || space >= e->host.spec.n_spaces) return fail(BBB_ERR_INVALID, "space index");
    if (!k_out) return fail(BBB_ERR_INVALID, "k_out is NULL");
    CU(launch_gather_space(e->dev, e->host.buf.n_chains, space, e->host.spec.spaces[space].kmax, k_out, sites_out,
                           values_out, (cudaStream_t)stream));
    return BBB_OK;
}

static int check_kb(bbb_engine *e, const int32_t *kb) {
    if (!kb) return fail(BBB_ERR_INVALID, "kb is NULL");
    for (int s = 0; s < e->host.spec.n_spaces; ++s)
        if (kb[s] < 1 || kb[s] > e->host.spec.spaces[s].kmax) return fail(BBB_ERR_INVALID, "kb[%d] outside [1, kmax]", s);
    return BBB_OK;
}

int64_t bbb_engine_state_image_bytes(bbb_engine *e, const int32_t *kb) {
    if (!e || !e->bound || check_kb(e, kb) != BBB_OK) return -1;
    return 8 * state_image_words(e->host, kb);
}

int bbb_engine_pack_states(bbb_engine *e, const int32_t *kb, void *image, void *stream) {
    NEED_BOUND(e);
    if (int rc = check_kb(e, kb)) return rc;
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    CU(launch_state_image(e->host, 0, kb, image, nullptr, 0, e->host.buf.n_chains, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_engine_compare_states(bbb_engine *e, const int32_t *kb, const void *image, int32_t *differ, void *stream) {
    NEED_BOUND(e);
    if (int rc = check_kb(e, kb)) return rc;
    if (!image || !differ) return fail(BBB_ERR_INVALID, "NULL argument");
    CU(launch_state_image(e->host, 1, kb, const_cast<void *>(image), differ, 0, e->host.buf.n_chains, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_engine_unpack_states(bbb_engine *e, const int32_t *kb, const void *image, void *stream) {
    NEED_BOUND(e);
    if (int rc = check_kb(e, kb)) return rc;
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    CU(launch_state_image(e->host, 2, kb, const_cast<void *>(image), nullptr, 0, e->host.buf.n_chains, (cudaStream_t)stream));
    return bbb_engine_refresh(e, stream);
}

int bbb_engine_pack_states_range(bbb_engine *e, const int32_t *kb, int64_t c_first, int64_t c_end, void *image, void *stream) {
    NEED_BOUND(e);
    if (int rc = check_kb(e, kb)) return rc;
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    if (c_first < 0 || c_end < c_first || c_end > e->host.buf.n_chains) return fail(BBB_ERR_INVALID, "chain range");
    CU(launch_state_image(e->host, 0, kb, image, nullptr, c_first, c_end, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_engine_hosted_parts(bbb_engine *e, int64_t *first_chain) {
    if (!e || !e->bound || !first_chain) return -1;
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e);
    for (int h = 0; h <= e->hosted_np; ++h) first_chain[h] = h * part < C ? h * part : C;
    return e->hosted_np;
}

// row of the image that holds k of space s
static long long image_k_row(const bbb_problem_spec &P, const int32_t *kb, int s) {
    long long row = 0;
    for (int q = 0; q < s; ++q) row += 1 + (long long)(P.spaces[q].ndim + P.spaces[q].n_params) * kb[q];
    return row;
}

// Developer trace (BBB_HOSTED_TRACE=1, implies no graphs): timed events behind every operation of a hosted step,
// printed by the wait as microseconds after the fork.
static cudaEvent_t g_trace_ev[4][8], g_trace_fork;
static bool g_trace_on = false, g_trace_init = false;
static void trace_mark(int h, int i, cudaStream_t s) {
    if (g_trace_on) cudaEventRecord(g_trace_ev[h][i], s);
}

// everything one hosted step puts on the device, from `st` through the parts' streams and back into `st`
// (does not touch the engine's host-side bookkeeping: it is also what a graph capture records)
#define ENQ(expr)                      \
    do {                               \
        int _e = (int)(expr);          \
        if (_e != 0) return _e;        \
    } while (0)
// image kernels of either wire format (mode: 0 pack, 1 compare, 2 unpack)
static int hosted_image(bbb_engine *e, const bbb_engine::HostedShape &sh, int mode, bool out_side, void *image, int32_t *differ,
                        long long c0, long long c1, cudaStream_t st) {
    if (sh.ragged) return launch_state_image_ragged(e->host, mode, image, differ, c0, c1, st);
    return launch_state_image(e->host, mode, out_side ? sh.kb_out : sh.kb_in, image, differ, c0, c1, st);
}

static int hosted_enqueue(bbb_engine *e, const bbb_engine::HostedShape &sh, const void *host_in, void *host_out,
                          void *dev_stage, void *dev_out, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e);
    const int NP = e->hosted_np;
    const long long *in_words = reinterpret_cast<const long long *>(host_in);
    ENQ(cudaMemsetAsync(e->hosted_flag_dev, 0, 4 * sizeof(int32_t), st));
    if (g_trace_on) cudaEventRecord(g_trace_fork, st);
    ENQ(cudaEventRecord(e->ev_hfork, st));
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(dev_stage), *out = reinterpret_cast<unsigned long long *>(dev_out);
    // Uploads: all parts back to back on the copy stream (they share one copy engine anyway).
    ENQ(cudaStreamWaitEvent(e->copy_stream, e->ev_hfork, 0));
    for (int h = 0; h < NP; ++h) {
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        if (c0 >= c1) continue;
        const long long in_off = sh.in_off[h];
        ENQ(cudaMemcpyAsync(stage + in_off, in_words + in_off, (size_t)sh.in_words[h] * 8, cudaMemcpyHostToDevice, e->copy_stream));
        ENQ(cudaEventRecord(e->ev_h2d[h], e->copy_stream));
    }
    // Counter-based generators make the proposal a pure function of (resident state, iteration, chain): it is drawn
    // SPECULATIVELY while the part's upload is in flight; should the uploaded states differ, the chain kernel is a
    // no-op and the wait proposes again from the written states.  Tape-driven engines (tests) consume their tape, so
    // there the proposal waits for the comparison and is guarded like the chain kernel.
    const bool speculative = e->host.buf.tape == nullptr;
    for (int h = 0; h < NP; ++h) {
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        if (c0 >= c1) continue;
        cudaStream_t hs = e->hosted_stream[h];
        const long long in_off = sh.in_off[h], out_off = sh.out_off[h];
        ENQ(cudaStreamWaitEvent(hs, e->ev_hfork, 0));
        trace_mark(h, 0, hs);
        if (speculative) ENQ(launch_propose(e->host, c0, c1, e->order_parity[h] * 4 + h, nullptr, hs));
        trace_mark(h, 1, hs);
        ENQ(cudaStreamWaitEvent(hs, e->ev_h2d[h], 0));
        trace_mark(h, 2, hs);
        ENQ(hosted_image(e, sh, 1, false, stage + in_off, e->hosted_flag_dev + h, c0, c1, hs));
        if (!speculative) ENQ(launch_propose(e->host, c0, c1, e->order_parity[h] * 4 + h, e->hosted_flag_dev + h, hs));
        trace_mark(h, 3, hs);
        ENQ(launch_chain_step(e->host, c0, c1, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax, idx_bytes_of(P, e->ray_t),
                              e->order_parity[h] * 4 + h, !speculative && e->dependent_launch >= 1, e->hosted_flag_dev + h, hs));
        trace_mark(h, 4, hs);
        ENQ(hosted_image(e, sh, 0, true, out + out_off, nullptr, c0, c1, hs));
        trace_mark(h, 5, hs);
        ENQ(cudaMemcpyAsync(reinterpret_cast<unsigned long long *>(host_out) + out_off, out + out_off, (size_t)sh.out_words[h] * 8,
                            cudaMemcpyDeviceToHost, hs));
        trace_mark(h, 6, hs);
        ENQ(cudaEventRecord(e->ev_hjoin[h], hs));
        ENQ(cudaStreamWaitEvent(st, e->ev_hjoin[h], 0));
    }
    ENQ(cudaMemcpyAsync(e->hosted_flag_host, e->hosted_flag_dev, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return 0;
}
#undef ENQ

// First-use resources of the hosted step.
static int hosted_init(bbb_engine *e) {
    if (e->hosted_flag_dev != nullptr) return BBB_OK;
    CU(cudaMalloc(&e->hosted_flag_dev, 4 * sizeof(int32_t)));
    CU(cudaHostAlloc(&e->hosted_flag_host, 4 * sizeof(int32_t), cudaHostAllocDefault));
    CU(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev_h2d) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->ev_hjoin) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->ev_hfork, cudaEventDisableTiming));
    for (auto &hs : e->hosted_stream) CU(cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking));
    const char *hg = getenv("BBB_HOSTED_GRAPH");
    e->hosted_graphs = (hg != nullptr && atoi(hg) == 0) ? 0 : 1;
    if (const char *tr = getenv("BBB_HOSTED_TRACE")) {
        if (atoi(tr) != 0 && !g_trace_init) {
            for (auto &row : g_trace_ev)
                for (auto &ev : row) cudaEventCreate(&ev);
            cudaEventCreate(&g_trace_fork);
            g_trace_init = g_trace_on = true;
        }
        if (g_trace_on) e->hosted_graphs = 0;
    }
    return BBB_OK;
}

// Put one hosted call of shape `sh` on the device (replayed from a captured graph when there is one) and do the
// engine's host-side bookkeeping of an iteration.
static int hosted_launch(bbb_engine *e, const bbb_engine::HostedShape &sh, const void *host_in, void *host_out, void *dev_stage,
                         void *dev_out, cudaStream_t st) {
    const bbb_problem_spec &P = e->host.spec;
    const int NP = e->hosted_np;
    if (int rc = hosted_init(e)) return rc;
    int parity = 0;
    for (int h = 0; h < NP; ++h) parity |= e->order_parity[h] << h;
    bool launched = false;
    if (e->hosted_graphs) {
        bbb_engine::HostedGraph *g = nullptr;
        for (auto &c : e->hosted_graph) {
            if (!c.used || c.host_in != host_in || c.host_out != host_out || c.dev_stage != dev_stage || c.dev_out != dev_out ||
                c.parity != parity)
                continue;
            if (c.shape.same(sh, P.n_spaces, NP)) { g = &c; break; }
        }
        if (g == nullptr) {
            bbb_engine::HostedGraph &c = e->hosted_graph[e->hosted_graph_next];
            e->hosted_graph_next = (e->hosted_graph_next + 1) % 16;
            if (c.used) {
                cudaGraphExecDestroy(c.exec);
                cudaGraphDestroy(c.graph);
                c.used = false;
            }
            int rc = (int)cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeRelaxed);
            if (rc == 0) {
                rc = hosted_enqueue(e, sh, host_in, host_out, dev_stage, dev_out, e->cap_stream);
                cudaGraph_t graph = nullptr;
                const int rc_end = (int)cudaStreamEndCapture(e->cap_stream, &graph);
                if (rc == 0) rc = rc_end;
                if (rc == 0) rc = (int)cudaGraphInstantiate(&c.exec, graph, 0);
                if (rc == 0) {
                    c.graph = graph;
                    c.used = true;
                    c.shape = sh;
                    c.host_in = host_in; c.host_out = host_out; c.dev_stage = dev_stage; c.dev_out = dev_out; c.parity = parity;
                    g = &c;
                } else if (graph != nullptr) {
                    cudaGraphDestroy(graph);
                }
            }
            if (g == nullptr) {  // capture is not available here: enqueue every call from now on
                (void)cudaGetLastError();
                e->hosted_graphs = 0;
            }
        }
        if (g != nullptr) {
            CU(cudaGraphLaunch(g->exec, st));
            launched = true;
        }
    }
    if (!launched) CU(hosted_enqueue(e, sh, host_in, host_out, dev_stage, dev_out, st));
    for (int h = 0; h < NP; ++h) e->order_parity[h] ^= 1;
    e->order_pending[0] = false;
    bump_k_bound(e, 1);
    ++e->steps_since_resync;
    if (e->resync_every > 0 && e->steps_since_resync >= e->resync_every) CU(chain_local_from_scratch(e, false, st));
    e->hosted.pending = true;
    e->hosted.shape = sh;
    e->hosted.host_out = host_out;
    e->hosted.dev_stage = dev_stage;
    e->hosted.dev_out = dev_out;
    return BBB_OK;
}

int bbb_engine_step_hosted(bbb_engine *e, const int32_t *kb_in, const void *host_in, int32_t *kb_out, void *host_out,
                           void *dev_stage, void *dev_out, void *stream) {
    NEED_BOUND(e);
    if (!e->host.chain_local) return fail(BBB_ERR_UNSUPPORTED, "the hosted step needs a chain-local engine (2-D ray target)");
    if (e->hosted.pending) return fail(BBB_ERR_INVALID, "the previous hosted step was not waited for (bbb_engine_step_hosted_wait)");
    if (int rc = check_kb(e, kb_in)) return rc;
    if (!host_in || !host_out || !kb_out || !dev_stage || !dev_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (host_in == host_out || dev_stage == dev_out) return fail(BBB_ERR_INVALID, "in and out images must be different buffers");
    const bbb_problem_spec &P = e->host.spec;
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e);
    const int NP = e->hosted_np;
    const long long rows_in = state_image_rows(e->host, kb_in);
    // rows of the downloaded image: the largest k of the uploaded chains plus the cell one step can add
    const long long *in_words = reinterpret_cast<const long long *>(host_in);
    for (int s = 0; s < P.n_spaces; ++s) {
        const long long krow = image_k_row(P, kb_in, s);
        long long kmax_seen = 1;
        for (int h = 0; h < NP; ++h) {
            const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
            const long long *kr = in_words + rows_in * c0 + krow * (c1 - c0);
            for (long long lc = 0; lc < c1 - c0; ++lc) kmax_seen = kr[lc] > kmax_seen ? kr[lc] : kmax_seen;
        }
        if (kmax_seen > kb_in[s]) return fail(BBB_ERR_INVALID, "space %d: the image holds k = %lld > kb = %d", s, kmax_seen, kb_in[s]);
        // (a multiple of 8, so that the image sizes -- and with them the captured graphs -- change rarely)
        const long long kb = P.spaces[s].trans_d ? ((kmax_seen + 1 + 7) & ~7LL) : kmax_seen;
        kb_out[s] = (int32_t)(kb < P.spaces[s].kmax ? kb : P.spaces[s].kmax);
    }
    bbb_engine::HostedShape sh{};
    sh.ragged = 0;
    for (int s = 0; s < P.n_spaces; ++s) { sh.kb_in[s] = kb_in[s]; sh.kb_out[s] = kb_out[s]; }
    const long long rows_out = state_image_rows(e->host, kb_out);
    for (int h = 0; h < NP; ++h) {
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        sh.in_off[h] = rows_in * c0;
        sh.in_words[h] = rows_in * (c1 - c0);
        sh.out_off[h] = rows_out * c0;
        sh.out_words[h] = rows_out * (c1 - c0);
    }
    return hosted_launch(e, sh, host_in, host_out, dev_stage, dev_out, (cudaStream_t)stream);
}

// ---- ragged hosted image: part h of a hosted image starts at word h * ragged_part_stride(e), whatever it holds ----
static long long ragged_part_words(const bbb_engine *e, long long n_c, const long long *cells /* per space */) {
    const bbb_problem_spec &P = e->host.spec;
    long long n = (long long)ragged_header_rows(e->host) * n_c;
    for (int s = 0; s < P.n_spaces; ++s) n += cells[s] * (P.spaces[s].ndim + P.spaces[s].n_params);
    return n;
}
static long long ragged_part_stride(const bbb_engine *e) {
    const bbb_problem_spec &P = e->host.spec;
    const long long part = hosted_part_chains(e);
    long long cells[BBB_MAX_SPACES];
    for (int s = 0; s < P.n_spaces; ++s) cells[s] = part * P.spaces[s].kmax;
    return (ragged_part_words(e, part, cells) + 511) & ~511LL;
}
constexpr long long RAGGED_COPY_QUANTUM = 4096;  // words: copy sizes are rounded up to 32 KiB, so that the captured graphs change rarely

int64_t bbb_engine_hosted_ragged_bytes(bbb_engine *e) {
    if (!e || !e->bound) return -1;
    return 8 * ragged_part_stride(e) * e->hosted_np;
}

int bbb_engine_pack_states_ragged(bbb_engine *e, void *image, void *stream) {
    NEED_BOUND(e);
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e), stride = ragged_part_stride(e);
    for (int h = 0; h < e->hosted_np; ++h) {
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        CU(launch_state_image_ragged(e->host, 0, reinterpret_cast<unsigned long long *>(image) + stride * h, nullptr, c0, c1, (cudaStream_t)stream));
    }
    return BBB_OK;
}

int bbb_engine_step_hosted_ragged(bbb_engine *e, const void *host_in, void *host_out, void *dev_stage, void *dev_out,
                                  int64_t *h2d_bytes, int64_t *d2h_bytes, void *stream) {
    NEED_BOUND(e);
    if (!e->host.chain_local) return fail(BBB_ERR_UNSUPPORTED, "the hosted step needs a chain-local engine (2-D ray target)");
    if (e->hosted.pending) return fail(BBB_ERR_INVALID, "the previous hosted step was not waited for (bbb_engine_step_hosted_wait)");
    if (!host_in || !host_out || !dev_stage || !dev_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (host_in == host_out || dev_stage == dev_out) return fail(BBB_ERR_INVALID, "in and out images must be different buffers");
    const bbb_problem_spec &P = e->host.spec;
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e), stride = ragged_part_stride(e);
    const int NP = e->hosted_np;
    const long long *in_words = reinterpret_cast<const long long *>(host_in);
    bbb_engine::HostedShape sh{};
    sh.ragged = 1;
    long long up = 0, down = 0;
    for (int h = 0; h < NP; ++h) {
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C, n_c = c1 - c0;
        sh.in_off[h] = sh.out_off[h] = stride * h;
        if (n_c <= 0) continue;
        // the k rows of the uploaded header say how many cells travel; a step adds at most one cell per chain
        long long cells_in[BBB_MAX_SPACES], cells_out[BBB_MAX_SPACES];
        for (int s = 0; s < P.n_spaces; ++s) {
            const long long *kr = in_words + stride * h + (long long)s * n_c;
            long long tot = 0;
            for (long long lc = 0; lc < n_c; ++lc) {
                if (kr[lc] < 0 || kr[lc] > P.spaces[s].kmax)
                    return fail(BBB_ERR_INVALID, "space %d: the image holds k = %lld for chain %lld (kmax %d)", s, kr[lc], c0 + lc, P.spaces[s].kmax);
                tot += kr[lc];
            }
            cells_in[s] = tot;
            cells_out[s] = tot + (P.spaces[s].trans_d ? n_c : 0);
            if (cells_out[s] > n_c * P.spaces[s].kmax) cells_out[s] = n_c * P.spaces[s].kmax;
        }
        auto rounded = [&](long long w) {
            w = (w + RAGGED_COPY_QUANTUM - 1) / RAGGED_COPY_QUANTUM * RAGGED_COPY_QUANTUM;
            return w < stride ? w : stride;
        };
        sh.in_words[h] = rounded(ragged_part_words(e, n_c, cells_in));
        sh.out_words[h] = rounded(ragged_part_words(e, n_c, cells_out));
        up += sh.in_words[h];
        down += sh.out_words[h];
    }
    if (h2d_bytes) *h2d_bytes = 8 * up;
    if (d2h_bytes) *d2h_bytes = 8 * down;
    return hosted_launch(e, sh, host_in, host_out, dev_stage, dev_out, (cudaStream_t)stream);
}

int bbb_engine_step_hosted_wait(bbb_engine *e, int32_t *n_redone, void *stream) {
    NEED_BOUND(e);
    if (!e->hosted.pending) return fail(BBB_ERR_INVALID, "no hosted step pending");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaStreamSynchronize(st));
    e->hosted.pending = false;
    if (g_trace_on) {
        static int n_printed = 0;
        if (++n_printed % 50 == 0) {
            const char *names[7] = {"start", "propose", "uploaded", "compare", "chain", "pack", "d2h"};
            for (int h = 0; h < e->hosted_np; ++h) {
                fprintf(stderr, "part %d:", h);
                for (int i = 0; i < 7; ++i) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, g_trace_fork, g_trace_ev[h][i]);
                    fprintf(stderr, " %s %.1f", names[i], ms * 1e3f);
                }
                fprintf(stderr, "\n");
            }
        }
    }
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    const long long C = e->host.buf.n_chains, part = hosted_part_chains(e);
    const int NP = e->hosted_np;
    int redo = 0;
    for (int h = 0; h < NP; ++h) redo += e->hosted_flag_host[h] != 0 ? 1 : 0;
    if (n_redone) *n_redone = redo;
    if (redo == 0) return BBB_OK;
    // Parts whose uploaded states differ from the resident ones were left untouched by the guarded kernels: write
    // their states, re-derive the caches of all chains (the other parts have already advanced; for them this is the
    // periodic from-scratch re-evaluation), advance the differing parts and download them.
    const bbb_engine::HostedShape &sh = e->hosted.shape;
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(e->hosted.dev_stage), *out = reinterpret_cast<unsigned long long *>(e->hosted.dev_out);
    for (int h = 0; h < NP; ++h) {
        if (e->hosted_flag_host[h] == 0) continue;
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        CU(hosted_image(e, sh, 2, false, stage + sh.in_off[h], nullptr, c0, c1, st));
    }
    if (int rc = bbb_engine_refresh(e, stream)) return rc;
    for (int h = 0; h < NP; ++h) {
        if (e->hosted_flag_host[h] == 0) continue;
        const long long c0 = h * part < C ? h * part : C, c1 = (h + 1) * part < C ? (h + 1) * part : C;
        CU(launch_propose(e->host, c0, c1, -1, nullptr, st));
        CU(launch_chain_step(e->host, c0, c1, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax, idx_bytes_of(P, e->ray_t), -1,
                             false, nullptr, st));
        CU(hosted_image(e, sh, 0, true, out + sh.out_off[h], nullptr, c0, c1, st));
        CU(cudaMemcpyAsync(reinterpret_cast<unsigned long long *>(e->hosted.host_out) + sh.out_off[h], out + sh.out_off[h],
                           (size_t)sh.out_words[h] * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return BBB_OK;
}

// ---- mapped hosted step (hosted.cuh): records in pinned host memory, read and written by the chain kernel itself ----
static HostedIO mapped_layout(const bbb_engine *e) {
    const bbb_problem_spec &P = e->host.spec;
    HostedIO io{};
    io.hrw = ragged_header_rows(e->host);
    long long off = io.hrw;
    for (int s = 0; s < P.n_spaces; ++s) {
        io.base[s] = (int)off;
        off += (long long)P.spaces[s].kmax * (P.spaces[s].ndim + P.spaces[s].n_params);
    }
    io.stride = (off + 15) & ~15LL;  // records start on 128-byte lines
    return io;
}
static int mapped_init(bbb_engine *e) {
    if (e->mapped_mismatch != nullptr) return BBB_OK;
    const long long C = e->host.buf.n_chains;
    CU(cudaMalloc(&e->mapped_mismatch, (size_t)C * sizeof(int32_t)));
    CU(cudaMemset(e->mapped_mismatch, 0, (size_t)C * sizeof(int32_t)));
    CU(cudaMalloc(&e->mapped_moved, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(e->mapped_moved, 0, 2 * sizeof(unsigned long long)));
    CU(cudaHostAlloc(&e->mapped_flag_host, 16 * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
    for (int i = 0; i < 16; ++i) e->mapped_flag_host[i] = 0;
    CU(cudaMalloc(&e->mapped_chain_done, (size_t)C * sizeof(int32_t)));
    CU(cudaMemset(e->mapped_chain_done, 0, (size_t)C * sizeof(int32_t)));
    e->mapped_seq = 0;
    CU(cudaEventCreateWithFlags(&e->ev_mapped_done, cudaEventDisableTiming));
    const char *sw = getenv("BBB_MAPPED_PDL");
    e->mapped_pdl = sw != nullptr ? atoi(sw) : 0;
    return BBB_OK;
}

int64_t bbb_engine_hosted_record_words(bbb_engine *e) {
    if (!e || !e->bound) return -1;
    return mapped_layout(e).stride;
}

int bbb_engine_pack_states_records(bbb_engine *e, void *image, void *stream) {
    NEED_BOUND(e);
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    HostedIO io = mapped_layout(e);
    io.out = reinterpret_cast<unsigned long long *>(image);
    CU(launch_hosted_pack(e->host, io, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_engine_step_hosted_mapped(bbb_engine *e, const void *host_in, void *host_out, void *stream) {
    const bool spec = e != nullptr && e->spec_valid;
    NEED_BOUND(e);
    if (!e->host.chain_local) return fail(BBB_ERR_UNSUPPORTED, "the hosted step needs a chain-local engine (2-D ray target)");
    if (e->hosted.pending || e->mapped.pending) return fail(BBB_ERR_INVALID, "the previous hosted step was not waited for");
    if (e->host.buf.tape != nullptr) return fail(BBB_ERR_UNSUPPORTED, "tape-driven engines use bbb_engine_step_hosted");
    if (!host_in || !host_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (host_in == host_out) return fail(BBB_ERR_INVALID, "in and out images must be different buffers");
    // the kernel reaches the images through their device aliases (the same addresses under unified addressing)
    const void *dev_alias[2] = {nullptr, nullptr};
    {
        const void *ptrs[2] = {host_in, host_out};
        for (int i = 0; i < 2; ++i) {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, ptrs[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) {
                (void)cudaGetLastError();
                return fail(BBB_ERR_INVALID, "hosted images must live in pinned (page-locked, device-mapped) host memory");
            }
            dev_alias[i] = at.devicePointer;
        }
    }
    if (int rc = mapped_init(e)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    const long long C = e->host.buf.n_chains;
    HostedIO io = mapped_layout(e);
    io.in = reinterpret_cast<const unsigned long long *>(dev_alias[0]);
    io.out = reinterpret_cast<unsigned long long *>(const_cast<void *>(dev_alias[1]));
    io.mismatch = e->mapped_mismatch;
    io.host_flag = e->mapped_flag_host;
    io.moved = e->mapped_moved;
    // The proposal of this iteration was drawn behind the previous hosted call (counter-based generators make it a pure
    // function of the resident states), unless something touched the engine in between.  Behind the chain kernel and the
    // event the wait blocks on goes the proposal kernel of the NEXT iteration: it runs during the host's turn-around
    // instead of standing between two calls.
    // Measured alternatives (B200, C3, 1024 chains; 0.131 ms per call as built): the chain set in 2 / 4 parts on internal
    // streams so that a part's proposal overlaps the other parts' chain kernels, 0.140 / 0.144 ms; the proposal kernel as
    // the chain kernel's programmatic dependent, every warp waiting for its own chain's completion signal
    // (BBB_MAPPED_PDL=1), 0.141 ms -- the waiting warps cost the chain kernel's tail more than the overlap gains; the
    // host polling a pinned word written by the kernel's last block behind system-wide fences instead of an event,
    // 0.154 ms -- a system-wide fence per block waits for the block's writes over the bus; comparing the upload with a
    // device-side shadow of the records last written (two contiguous reads instead of scattered reads of the resident
    // arrays), 0.130 ms -- no gain: what the comparison costs is the read over the bus, not the resident side.
    const int seq = ++e->mapped_seq;
    io.chain_done = e->mapped_pdl ? e->mapped_chain_done : nullptr;
    io.seq = seq;
    if (!spec) CU(run_propose(e, st));
    CU(launch_chain_step(e->host, 0, C, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax, idx_bytes_of(P, e->ray_t),
                         e->order_pending[0] ? e->order_parity[0] * 4 : -1, !spec && e->dependent_launch >= 1, nullptr, st, &io));
    if (e->order_pending[0]) { e->order_parity[0] ^= 1; e->order_pending[0] = false; }
    e->order_pending[0] = true;
    CU(cudaEventRecord(e->ev_mapped_done, st));  // what the wait blocks on
    CU(launch_propose(e->host, 0, C, e->order_parity[0] * 4, nullptr, st, e->mapped_pdl ? e->mapped_chain_done : nullptr, seq,
                      e->mapped_flag_host + 1));
    bump_k_bound(e, 1);
    ++e->steps_since_resync;
    if (e->resync_every > 0 && e->steps_since_resync >= e->resync_every) CU(chain_local_from_scratch(e, false, st));
    e->spec_valid = true;
    e->mapped.pending = true;
    e->mapped.host_in = dev_alias[0];
    e->mapped.host_out = const_cast<void *>(dev_alias[1]);
    return BBB_OK;
}

int bbb_engine_step_hosted_mapped_wait(bbb_engine *e, int32_t *n_redone, void *stream) {
    const bool spec = e != nullptr && e->spec_valid;
    NEED_BOUND(e);
    if (!e->mapped.pending) return fail(BBB_ERR_INVALID, "no mapped hosted step pending");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaEventSynchronize(e->ev_mapped_done));
    e->mapped.pending = false;
    if (e->mapped_flag_host[1] != 0) {
        e->mapped_flag_host[1] = 0;
        e->spec_valid = false;
        return fail(BBB_ERR_CUDA, "a proposal warp gave up waiting for its chain's completion signal");
    }
    e->spec_valid = spec;
    if (n_redone) *n_redone = 0;
    if (*reinterpret_cast<volatile int32_t *>(e->mapped_flag_host) == 0) return BBB_OK;
    // Chains whose uploaded record differed from the resident state were left untouched: write their records into the
    // engine, re-derive the caches of all chains (for the others this is the periodic from-scratch re-evaluation),
    // advance exactly those chains (their new records go to the out image) and draw the next proposals again.
    e->mapped_flag_host[0] = 0;
    const bbb_problem_spec &P = e->host.spec;
    const bbb_target_spec &T = P.targets[e->ray_t];
    const long long C = e->host.buf.n_chains;
    CU(cudaStreamSynchronize(st));
    if (n_redone) {
        int32_t *flags = (int32_t *)malloc((size_t)C * sizeof(int32_t));
        if (flags == nullptr) return fail(BBB_ERR_INVALID, "out of host memory");
        const cudaError_t err = cudaMemcpy(flags, e->mapped_mismatch, (size_t)C * sizeof(int32_t), cudaMemcpyDeviceToHost);
        int n = 0;
        for (long long c = 0; c < C && err == cudaSuccess; ++c) n += flags[c] != 0 ? 1 : 0;
        free(flags);
        CU(err);
        *n_redone = n;
    }
    HostedIO io = mapped_layout(e);
    io.in = reinterpret_cast<const unsigned long long *>(e->mapped.host_in);
    io.out = reinterpret_cast<unsigned long long *>(e->mapped.host_out);
    io.mismatch = e->mapped_mismatch;
    io.host_flag = e->mapped_flag_host;
    io.moved = e->mapped_moved;
    io.only = e->mapped_mismatch;
    io.chain_done = nullptr;  // (the redo pass is waited for with a stream synchronisation)
    CU(launch_hosted_unpack(e->host, io, st));
    if (int rc = bbb_engine_refresh(e, stream)) return rc;
    CU(launch_propose(e->host, 0, C, -1, nullptr, st));
    // (the flags select the chains and are only raised by a chain whose record still differs -- which cannot happen now)
    CU(launch_chain_step(e->host, 0, C, e->ray_t, T.csc_tile_rays, T.n_data, P.spaces[T.space].kmax, idx_bytes_of(P, e->ray_t), -1, false,
                         nullptr, st, &io));
    CU(cudaMemsetAsync(e->mapped_mismatch, 0, (size_t)C * sizeof(int32_t), st));
    CU(launch_propose(e->host, 0, C, -1, nullptr, st));  // the next iteration's proposals, now from the redone states
    CU(cudaStreamSynchronize(st));
    e->mapped_flag_host[0] = 0;
    e->spec_valid = true;
    return BBB_OK;
}

int bbb_engine_hosted_mapped_bytes(bbb_engine *e, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!e || !e->bound) return fail(BBB_ERR_INVALID, "NULL / unbound engine");
    unsigned long long v[2] = {0, 0};
    if (e->mapped_moved != nullptr) CU(cudaMemcpy(v, e->mapped_moved, sizeof(v), cudaMemcpyDeviceToHost));
    if (h2d_bytes) *h2d_bytes = (int64_t)v[0];
    if (d2h_bytes) *d2h_bytes = (int64_t)v[1];
    return BBB_OK;
}

int bbb_engine_dpred(bbb_engine *e, int target, double *dpred_out, void *stream) {
    NEED_BOUND(e);
    if (target < 0 || target >= e->host.spec.n_targets || !dpred_out) return fail(BBB_ERR_INVALID, "target index / NULL output");
    if (e->host.spec.targets[target].fwd_kind == BBB_FWD_RAY2D) {
        if (e->host.chain_local) CU(sync_index_slot0(e, (cudaStream_t)stream));
        CU(run_spmm(e, target, false, dpred_out, false, (cudaStream_t)stream));
    }
    else
        CU(launch_small_dpred(e->dev, e->host.buf.n_chains, target, dpred_out, (cudaStream_t)stream));
    return BBB_OK;
}

int64_t bbb_engine_save_point_bytes(bbb_engine *e, const int32_t *kb, int32_t want_dpred) {
    if (!e || !e->bound || check_kb(e, kb) != BBB_OK) return -1;
    long long n = save_point_state_words(e->host, kb);
    if (want_dpred)
        for (int t = 0; t < e->host.spec.n_targets; ++t) n += e->host.buf.n_chains * (long long)e->host.spec.targets[t].n_data;
    return 8 * n;
}

int bbb_engine_save_point(bbb_engine *e, const int32_t *kb, int32_t want_dpred, void *image, double *scratch, void *stream) {
    NEED_BOUND(e);
    if (int rc = check_kb(e, kb)) return rc;
    if (!image) return fail(BBB_ERR_INVALID, "image is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const bbb_problem_spec &P = e->host.spec;
    const bbb_buffers &B = e->host.buf;
    const long long C = B.n_chains;
    CU(launch_save_states(e->host, kb, image, st));
    if (!want_dpred) return BBB_OK;
    double *out = reinterpret_cast<double *>(image) + save_point_state_words(e->host, kb);
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        if (e->host.chain_local && t == e->ray_t) {
            // the residuals of the current states are resident, chain-major: d_pred = res + d_obs
            CU(launch_save_dpred_local(B.res[t], T.dobs, C, T.n_data, out, st));
        } else {
            if (!scratch) return fail(BBB_ERR_INVALID, "target %d: a [n_data][C] scratch buffer is needed for its d_pred", t);
            if (int rc = bbb_engine_dpred(e, t, scratch, stream)) return rc;
            CU(launch_transpose_f64(scratch, out, T.n_data, C, st));
        }
        out += C * (long long)T.n_data;
    }
    return BBB_OK;
}

int bbb_engine_misfit_logdet(bbb_engine *e, double *misfit_out, double *logdet_out, void *stream) {
    NEED_BOUND(e);
    if (!misfit_out || !logdet_out) return fail(BBB_ERR_INVALID, "NULL output");
    CU(launch_misfit_logdet(e->dev, e->host.buf.n_chains, misfit_out, logdet_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_engine_swap_rows(bbb_engine *e, double *rows_out, void *stream) {
    NEED_BOUND(e);
    if (!rows_out) return fail(BBB_ERR_INVALID, "NULL output");
    CU(launch_misfit_logdet(e->dev, e->host.buf.n_chains, rows_out, nullptr, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_pt_swap(double *gathered, int64_t n_total, uint64_t seed, int64_t round, int64_t id0, int64_t n_local,
                double *temperature_out, int32_t *n_accepted_out, const double *tape, void *stream) {
    if (!gathered || !temperature_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (n_total < 2) return fail(BBB_ERR_INVALID, "parallel tempering needs at least 2 chains");
    if (id0 < 0 || n_local < 0 || id0 + n_local > n_total) return fail(BBB_ERR_INVALID, "local chain range");
    CU(launch_pt_swap(gathered, n_total, seed, round, id0, n_local, temperature_out, n_accepted_out, tape,
                      (cudaStream_t)stream));
    return BBB_OK;
}

// ---- stand-alone kernels ------------------------------------------------------------------------
int bbb_raster2d(const double *sites, const int32_t *k, int32_t K, int64_t C, const double *points_x,
                 const double *points_y, int32_t G, void *idx_out, void *stream) {
    if (!sites || !k || !points_x || !points_y || !idx_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (K < 1 || K > 65535 || C < 1 || G < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    RasterArgs a;
    memset(&a, 0, sizeof(a));
    a.sites = sites;
    a.k = k;
    a.K = K;
    a.k_rows = K;
    a.C = C;
    a.px = points_x;
    a.py = points_y;
    a.G = G;
    a.idx = idx_out;
    a.idx_bytes = K <= 256 ? 1 : 2;
    a.points_per_block = raster_points_per_block(G, C);
    CU(launch_raster2d(a, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_raster1d(const double *sites, const double *values, const int32_t *k, int32_t K, int64_t C, const double *x,
                 int32_t R, int32_t *idx_out, double *field_out, void *stream) {
    if (!sites || !k || !x) return fail(BBB_ERR_INVALID, "NULL argument");
    if (field_out && !values) return fail(BBB_ERR_INVALID, "field requested without values");
    if (K < 1 || C < 1 || R < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    CU(launch_raster1d(sites, values, k, K, C, x, R, idx_out, field_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_ray_forward(const void *idx, const double *values, const int32_t *k, int32_t K, int64_t C, int32_t G,
                    int32_t R, const int32_t *indptr, const int32_t *indices, const double *data, int32_t transform,
                    const double *dobs, const double *w, double *dpred_out, double *phi_out, double *partial_scratch,
                    void *stream) {
    if (!idx || !values || !k || !indptr || !indices || !data || !dobs) return fail(BBB_ERR_INVALID, "NULL argument");
    if (phi_out && !partial_scratch) return fail(BBB_ERR_INVALID, "phi requested without partial scratch");
    if (K < 1 || K > 65535 || C < 1 || G < 1 || R < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    SpmmArgs a;
    memset(&a, 0, sizeof(a));
    a.idx = idx;
    a.idx_bytes = K <= 256 ? 1 : 2;
    a.values = values;
    a.k = k;
    a.n_params = 1;
    a.param = 0;
    a.K = K;
    a.k_rows = K;
    a.C = C;
    a.G = G;
    a.R = R;
    a.indptr = indptr;
    a.indices = indices;
    a.data = data;
    a.transform = transform;
    a.dobs = dobs;
    a.w = w;
    a.dpred = dpred_out;
    a.partial = phi_out ? partial_scratch : nullptr;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    a.n_groups = spmm_ray_groups(R, C, K, a.idx_bytes, n_sm);
    if (a.n_groups == 0) return fail(BBB_ERR_UNSUPPORTED, "K = %d is too large for the ray kernel's shared-memory value table", K);
    CU(launch_ray_spmm(a, (cudaStream_t)stream));
    if (phi_out) CU(launch_phi_from_partial(partial_scratch, a.n_groups, C, phi_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_dense_forward(const double *G, const double *m, int32_t R, int32_t M, int64_t C, double *dpred_out,
                      void *stream) {
    if (!G || !m || !dpred_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (R < 1 || M < 1 || C < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    CU(launch_dense_forward(G, m, R, M, C, dpred_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_loglike(const double *dpred, const double *dobs, const double *sigma, int32_t R, int64_t C, double *misfit_out,
                double *logdet_out, void *stream) {
    if (!dpred || !dobs || !sigma || !misfit_out || !logdet_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (R < 1 || C < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    CU(launch_loglike(dpred, dobs, sigma, R, C, misfit_out, logdet_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_logprior(const bbb_space_spec *space, const double *values, const int32_t *k, int64_t C, double *log_prior_out,
                 void *stream) {
    if (!space || !values || !k || !log_prior_out) return fail(BBB_ERR_INVALID, "NULL argument");
    if (space->n_params < 1 || space->n_params > BBB_MAX_PARAMS || C < 1) return fail(BBB_ERR_INVALID, "bad sizes");
    CU(launch_logprior(*space, values, k, C, log_prior_out, (cudaStream_t)stream));
    return BBB_OK;
}

int bbb_philox_kat(const uint32_t *counter, const uint32_t *key, uint32_t *out) {
    if (!counter || !key || !out) return fail(BBB_ERR_INVALID, "NULL argument");
    philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1], out);
    return BBB_OK;
}

}  // extern "C"
