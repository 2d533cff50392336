"""Device engine: owns the HBM buffers (torch tensors, buffer management only) and drives the
CUDA kernels of ``libbbb200.so`` through its C ABI.

State layout: every per-chain array is structure-of-arrays with
the chain index fastest and cells padded to ``n_dimensions_max`` --
``sites [2][ndim][K][C]``, ``values [2][P][K][C]``, ``k [2][C]`` per parameter space (two slots:
current / proposed, selected per chain by ``sel[C]``), ``cell_idx [2][G][C]`` (uint8 when
K <= 256) per ray target, and ``[C]`` scalars (sigma, phi = sum of squared residuals,
temperature, iteration, counters).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L

_PRIOR_KIND = {"uniform": L.PRIOR_UNIFORM, "gaussian": L.PRIOR_GAUSSIAN, "laplace": L.PRIOR_LAPLACE}
_SPACE_KIND = {"plain": L.SPACE_PLAIN, "voronoi1d": L.SPACE_VORONOI1D, "voronoi2d": L.SPACE_VORONOI2D}
_FWD_KIND = {"ray2d": L.FWD_RAY2D, "lookup1d": L.FWD_LOOKUP1D, "linear": L.FWD_LINEAR, "polynomial": L.FWD_POLYNOMIAL}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def ray_matrix_csc_tiles(indptr, indices, data, n_points: int, tile_rays: int):
    """Ray-tile-blocked CSC copy of a CSR ray matrix (one level of ``bbb_target_spec.csc_*``): entries sorted by
    (ray tile, column, ray); returns ``(colptr [n_tiles][G+1] int32, rows [nnz] uint16 -- ray index inside its
    tile --, data [nnz] float64, bound = max over rays of sum |data|)``."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    data = np.asarray(data, dtype=np.float64)
    n_rays = indptr.size - 1
    n_tiles = max(1, -(-n_rays // tile_rays))
    lens = np.diff(indptr)
    rows = np.repeat(np.arange(n_rays, dtype=np.int64), lens)
    tile = rows // tile_rays
    order = np.lexsort((rows, indices, tile))
    counts = np.bincount(tile * n_points + indices, minlength=n_tiles * n_points)
    cum = np.concatenate(([0], np.cumsum(counts)))
    colptr = np.empty((n_tiles, n_points + 1), dtype=np.int64)
    for q in range(n_tiles):
        colptr[q] = cum[q * n_points: (q + 1) * n_points + 1]
    if cum[-1] >= 2 ** 30:
        raise ValueError("ray matrix too large for the CSC offsets")
    bound = float(np.bincount(rows, weights=np.abs(data), minlength=n_rays).max(initial=0.0))
    return (colptr.astype(np.int32), (rows[order] - tile[order] * tile_rays).astype(np.uint16),
            np.ascontiguousarray(data[order]), bound)


def ray_matrix_csc_levels(indptr, indices, data, n_points: int, tile_rays: int, n_levels: int = L.CSC_LEVELS):
    """All levels of the CSC copy: level l sums every aligned group of 4**l consecutive columns into one."""
    import scipy.sparse as sp

    n_rays = np.asarray(indptr).size - 1
    A = sp.csr_matrix((np.asarray(data, dtype=np.float64), np.asarray(indices), np.asarray(indptr)), shape=(n_rays, n_points))
    out = []
    for lvl in range(n_levels):
        s = 4 ** lvl
        if lvl == 0:
            M = A
        else:
            n_groups = -(-n_points // s)
            S = sp.csr_matrix((np.ones(n_points), (np.arange(n_points), np.arange(n_points) // s)), shape=(n_points, n_groups))
            M = (A @ S).tocsr()
            M.sum_duplicates()
        out.append(ray_matrix_csc_tiles(M.indptr, M.indices, M.data, M.shape[1], tile_rays))
    return out


def morton_order(points):
    """Permutation that lists 2-D points along a Morton (Z-order) curve of their coordinate RANKS, so that aligned
    groups of 4 / 16 consecutive points are 2x2 / 4x4 blocks of a regular grid (and compact clusters of any other
    point set).  ``perm[new] = old``."""
    pts = np.asarray(points, dtype=np.float64)
    ix = np.unique(pts[:, 0], return_inverse=True)[1].astype(np.uint64)
    iy = np.unique(pts[:, 1], return_inverse=True)[1].astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0xFFFFFFFF)
        for sh, mask in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                         (2, 0x3333333333333333), (1, 0x5555555555555555)):
            v = (v | (v << np.uint64(sh))) & np.uint64(mask)
        return v

    return np.argsort(spread(ix) | (spread(iy) << np.uint64(1)), kind="stable")


class Engine:
    """One engine per (process, device).  ``spec`` is the plain-data problem description emitted by
    :func:`bayesbay_b200._compile.compile_problem`."""

    def __init__(self, spec: dict, n_chains: int, device="cuda:0", seed: int = 0, chain_id0: int = 0,
                 tape: Optional[np.ndarray] = None):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("bayesbay_b200 needs a CUDA device (no CPU fallback)")
        self.spec = spec
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.C = int(n_chains)
        self.seed = int(seed)
        self.chain_id0 = int(chain_id0)
        self._keep: List[torch.Tensor] = []  # device constants borrowed by the C side
        self.point_perm = {}  # ray target -> (perm, inv) of its query points (device order = points[perm])
        self._handle = C.c_void_p()
        self._cspec = self._build_spec()
        L.check(self.lib.bbb_engine_create(C.byref(self._cspec), self.dev_index, C.byref(self._handle)))
        self._alloc(tape)
        L.check(self.lib.bbb_engine_bind_buffers(self._handle, C.byref(self._cbuf)))
        self._steps_since_sync = 0
        self._has_ray = any(t["forward"]["kind"] == "ray2d" for t in spec["targets"])
        self.chain_local = bool(self.lib.bbb_engine_chain_local(self._handle))

    # ------------------------------------------------------------------ spec / buffers
    def _const(self, arr, dtype):
        t = torch.as_tensor(np.ascontiguousarray(arr, dtype=dtype)).to(self.device)
        self._keep.append(t)
        return t

    def _build_spec(self) -> L.ProblemSpec:
        spec = self.spec
        ps = L.ProblemSpec()
        if len(spec["spaces"]) > L.MAX_SPACES or len(spec["targets"]) > L.MAX_TARGETS:
            raise ValueError("too many parameter spaces / targets for the device engine")
        ps.n_spaces = len(spec["spaces"])
        ps.n_targets = len(spec["targets"])
        ps.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        ps.w_noise = -1.0 if spec.get("w_noise") is None else float(spec["w_noise"])
        for i, sp in enumerate(spec["spaces"]):
            s = ps.spaces[i]
            s.kind = _SPACE_KIND[sp["kind"]]
            s.ndim = sp["ndim"]
            s.trans_d = int(bool(sp["trans_d"]))
            s.kmin, s.kmax, s.kinit_max = sp["kmin"], sp["kmax"], sp["kinit_max"]
            s.birth_from = L.BIRTH_FROM_NEIGHBOUR if sp["birth_from"] == "neighbour" else L.BIRTH_FROM_PRIOR
            if len(sp["params"]) > L.MAX_PARAMS:
                raise ValueError(f"parameter space {sp['name']!r}: more than {L.MAX_PARAMS} parameters")
            s.n_params = len(sp["params"])
            for d in range(sp["ndim"]):
                s.vmin[d], s.vmax[d], s.site_std[d] = sp["vmin"][d], sp["vmax"][d], sp["site_std"][d]
            w = sp["weights"]
            s.w_birth, s.w_death, s.w_values, s.w_sites = w["birth"], w["death"], w["values"], w["sites"]
            s.w_outer = float(sp.get("w_outer") or 0.0)
            for j, p in enumerate(sp["params"]):
                s.prior_kind[j] = _PRIOR_KIND[p["kind"]]
                tab = p.get("table")
                if tab is None:
                    s.prior_a[j], s.prior_b[j] = p["a"], p["b"]
                    s.perturb_std[j], s.perturb_std_birth[j] = p["perturb_std"], p["perturb_std_birth"]
                else:  # position-dependent hyper-parameters: [5][n] rows position, a, b, perturb_std, perturb_std_birth
                    rows = np.stack([np.asarray(tab[k], dtype=np.float64) for k in
                                     ("position", "a", "b", "perturb_std", "perturb_std_birth")])
                    if np.any(np.diff(rows[0]) <= 0):
                        raise ValueError(f"prior {p['name']!r}: `position` must be strictly ascending")
                    s.n_pos[j] = rows.shape[1]
                    s.pos_table[j] = self._const(rows, np.float64).data_ptr()
        for i, tg in enumerate(spec["targets"]):
            t = ps.targets[i]
            dobs = self._const(tg["dobs"], np.float64)
            t.n_data = dobs.numel()
            t.dobs = dobs.data_ptr()
            if tg["hierarchical"]:
                t.cov_kind = L.COV_HIERARCHICAL
                t.std_min, t.std_max, t.std_perturb_std = tg["std_min"], tg["std_max"], tg["std_perturb_std"]
                if tg.get("correlated"):
                    t.correlated = 1
                    t.corr_min, t.corr_max, t.corr_perturb_std = tg["corr_min"], tg["corr_max"], tg["corr_perturb_std"]
            elif np.isscalar(tg["cov_inv"]):
                t.cov_kind = L.COV_SCALAR
                t.cov_scalar = float(tg["cov_inv"])
            else:
                cov = np.asarray(tg["cov_inv"], dtype=np.float64)
                if cov.ndim == 2 and cov.shape == (t.n_data, t.n_data):
                    # `covariance_mat_inv @ vector` (likelihood/_target.py:150-151): row-major [n][n]
                    t.cov_kind = L.COV_MATRIX
                elif cov.ndim == 1 and cov.size == t.n_data:
                    t.cov_kind = L.COV_VECTOR
                else:
                    raise ValueError("an inverse covariance is a scalar, a per-datum vector [n] or a matrix [n][n]")
                t.cov_vec = self._const(np.ascontiguousarray(cov), np.float64).data_ptr()
            f = tg["forward"]
            t.fwd_kind = _FWD_KIND[f["kind"]]
            t.space = f["space"]
            if f["kind"] == "ray2d":
                t.param = f["param"]
                t.transform = L.TRANSFORM_RECIPROCAL if f.get("transform", "reciprocal") == "reciprocal" else L.TRANSFORM_IDENTITY
                pts = np.asarray(f["points"], dtype=np.float64)
                t.n_points = pts.shape[0]
                # the device sees the query points in Morton order (internal: current_cell_idx() undoes it)
                K = spec["spaces"][f["space"]]["kmax"]
                tile = int(self.lib.bbb_chain_local_tile_rays(t.n_data, K, t.n_points))
                local = tile > 0 and not tg.get("correlated") and not (not tg["hierarchical"] and np.ndim(tg["cov_inv"]) == 2)
                perm = morton_order(pts) if local else np.arange(t.n_points)
                inv = np.empty_like(perm)
                inv[perm] = np.arange(t.n_points)
                self.point_perm[i] = (torch.as_tensor(perm).to(self.device), torch.as_tensor(inv).to(self.device))
                pts = pts[perm]
                f = dict(f, indices=inv[np.asarray(f["indices"])])
                t.nnz = int(np.asarray(f["indices"]).size)
                if np.asarray(f["indptr"]).size != t.n_data + 1:
                    raise ValueError("ray matrix rows do not match the number of data")
                if t.nnz and int(np.max(f["indices"])) >= t.n_points:
                    raise ValueError("ray matrix column index outside the query grid")
                t.indptr = self._const(f["indptr"], np.int32).data_ptr()
                t.indices = self._const(f["indices"], np.int32).data_ptr()
                t.data = self._const(f["data"], np.float64).data_ptr()
                t.points_x = self._const(pts[:, 0], np.float64).data_ptr()
                t.points_y = self._const(pts[:, 1], np.float64).data_ptr()
                # CSC copy for the chain-local step (tile == 0: this shape cannot run it, the engine does full SpMMs)
                if local:
                    levels = ray_matrix_csc_levels(f["indptr"], f["indices"], f["data"], t.n_points, tile)
                    t.csc_tile_rays, t.csc_n_tiles, t.csc_n_levels = tile, levels[0][0].shape[0], len(levels)
                    for lvl, (colptr, rows, cdat, bound) in enumerate(levels):
                        t.csc_colptr[lvl] = self._const(colptr, np.int32).data_ptr()
                        t.csc_rows[lvl] = self._const(rows.view(np.int16), np.int16).data_ptr()
                        t.csc_data[lvl] = self._const(cdat, np.float64).data_ptr()
                    t.csc_bound = levels[0][3]
                    # bounding boxes of the 16-point groups (Morton order: compact blocks) for the rasteriser's pruning
                    n_grp = (t.n_points + 15) // 16
                    padded = np.concatenate([pts, np.repeat(pts[-1:], n_grp * 16 - t.n_points, axis=0)]).reshape(n_grp, 16, 2)
                    box = np.stack([padded[:, :, 0].min(1), padded[:, :, 0].max(1), padded[:, :, 1].min(1), padded[:, :, 1].max(1)], 1)
                    if not os.environ.get("BBB_NO_GROUP_BOX"):  # (developer switch: A/B of the pruning)
                        t.group_box = self._const(box, np.float64).data_ptr()
            elif f["kind"] in ("lookup1d", "polynomial"):
                t.param = f["param"]
                x = np.asarray(f["x"], dtype=np.float64)
                if x.size != t.n_data:
                    raise ValueError("data positions do not match the number of data")
                t.x = self._const(x, np.float64).data_ptr()
            else:
                G = np.asarray(f["G"], dtype=np.float64)
                K = spec["spaces"][f["space"]]["kmax"]
                if G.shape != (t.n_data, len(f["params"]) * K):
                    raise ValueError(f"linear operator shape {G.shape} != ({t.n_data}, {len(f['params']) * K})")
                t.n_lin_params = len(f["params"])
                for q, p in enumerate(f["params"]):
                    t.lin_params[q] = p
                t.G = self._const(G, np.float64).data_ptr()
        return ps

    def _alloc(self, tape):
        Cn, dev = self.C, self.device
        f64, i32, i64, u8 = torch.float64, torch.int32, torch.int64, torch.uint8
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        b = L.Buffers()
        b.n_chains, b.chain_id0 = Cn, self.chain_id0
        self.sites, self.values, self.k, self.sel = [], [], [], []
        for i, sp in enumerate(self.spec["spaces"]):
            K, nd, P = sp["kmax"], sp["ndim"], len(sp["params"])
            self.sites.append(z((2, nd, K, Cn), f64) if nd else None)
            self.values.append(z((2, P, K, Cn), f64) if P else None)
            self.k.append(z((2, Cn), i32))
            self.sel.append(z((Cn,), u8))
            b.sites[i], b.values[i] = _ptr(self.sites[i]), _ptr(self.values[i])
            b.k[i], b.sel[i] = _ptr(self.k[i]), _ptr(self.sel[i])
        self.sigma, self.phi, self.cell_idx, self.idx_sel, self.partial = [], [], [], [], []
        self.rescan_list, self.rescan_count = [], []
        self.idx_cm, self.res, self.change_list = [], [], []
        self.corr, self.dpred_scratch = [], []
        n_ray = sum(tg["forward"]["kind"] == "ray2d" for tg in self.spec["targets"])
        for i, tg in enumerate(self.spec["targets"]):
            self.sigma.append(torch.ones((2, Cn), dtype=f64, device=dev) if tg["hierarchical"] else None)
            self.phi.append(z((2, 3, Cn), f64))
            corr_t = bool(tg.get("correlated"))
            self.corr.append(z((2, Cn), f64) if corr_t else None)
            # chain-local step: one ray target whose CSC copy was built (see _build_spec)
            local_t = tg["forward"]["kind"] == "ray2d" and n_ray == 1 and self._cspec.targets[i].csc_tile_rays > 0
            matrix_t = (not tg["hierarchical"]) and np.ndim(tg["cov_inv"]) == 2  # residuals kept for res^T M res
            self.dpred_scratch.append(z((int(np.asarray(tg["dobs"]).size), Cn), f64)
                                      if ((corr_t or local_t) and tg["forward"]["kind"] == "ray2d") or matrix_t else None)
            b.corr[i], b.dpred_scratch[i] = _ptr(self.corr[i]), _ptr(self.dpred_scratch[i])
            if local_t:
                G = int(np.asarray(tg["forward"]["points"]).shape[0])
                K = self.spec["spaces"][tg["forward"]["space"]]["kmax"]
                self.idx_cm.append(z((Cn, int(self.lib.bbb_idx_row_stride(G))), u8 if K <= 256 else torch.int16))
                self.res.append(z((Cn, int(np.asarray(tg["dobs"]).size)), f64))
                self.change_list.append(torch.empty((Cn, G + (G + 15) // 16), dtype=i32, device=dev))
            else:
                self.idx_cm.append(None)
                self.res.append(None)
                self.change_list.append(None)
            b.idx_cm[i], b.res[i], b.change_list[i] = _ptr(self.idx_cm[i]), _ptr(self.res[i]), _ptr(self.change_list[i])
            if tg["forward"]["kind"] == "ray2d":
                K = self.spec["spaces"][tg["forward"]["space"]]["kmax"]
                G = int(np.asarray(tg["forward"]["points"]).shape[0])
                tiles = int(self.lib.bbb_ray_tiles(int(np.asarray(tg["dobs"]).size)))
                self.cell_idx.append(z((2, G, Cn), u8 if K <= 256 else torch.int16))
                self.idx_sel.append(z((Cn,), u8))
                self.partial.append(z((tiles, Cn), f64))
                self.rescan_list.append(torch.empty((Cn, G), dtype=i32, device=dev))
                self.rescan_count.append(z((Cn,), i32))
            else:
                self.rescan_list.append(None)
                self.rescan_count.append(None)
                self.cell_idx.append(None)
                self.idx_sel.append(None)
                self.partial.append(None)
            b.sigma[i], b.phi[i] = _ptr(self.sigma[i]), _ptr(self.phi[i])
            b.cell_idx[i], b.idx_sel[i], b.partial[i] = _ptr(self.cell_idx[i]), _ptr(self.idx_sel[i]), _ptr(self.partial[i])
            b.rescan_list[i], b.rescan_count[i] = _ptr(self.rescan_list[i]), _ptr(self.rescan_count[i])
        nm = 4 * len(self.spec["spaces"]) + 1
        self.temperature = torch.ones((Cn,), dtype=f64, device=dev)
        self.iteration = z((Cn,), i64)
        self.move = z((Cn,), i32)
        self.edit = z((2, Cn), i32)
        self.edit_site = z((2, Cn), f64)
        self.edit_vals = z((L.MAX_PARAMS, Cn), f64)
        self.log_prob_ratio = z((Cn,), f64)
        self.log_like_ratio = z((Cn,), f64)
        self.accepted = z((Cn,), i32)
        self.n_proposed = z((nm, Cn), i64)
        self.n_accepted = z((nm, Cn), i64)
        self.draw_cursor = z((Cn,), i64)
        self.n_exceptions = z((2, Cn), i64)  # non-finite likelihood ratios, resampling loops that gave up
        b.temperature, b.iteration = _ptr(self.temperature), _ptr(self.iteration)
        b.move, b.edit = _ptr(self.move), _ptr(self.edit)
        b.edit_site, b.edit_vals = _ptr(self.edit_site), _ptr(self.edit_vals)
        b.log_prob_ratio, b.log_like_ratio, b.accepted = _ptr(self.log_prob_ratio), _ptr(self.log_like_ratio), _ptr(self.accepted)
        b.n_proposed, b.n_accepted = _ptr(self.n_proposed), _ptr(self.n_accepted)
        b.tape_cursor = _ptr(self.draw_cursor)
        b.n_exceptions = _ptr(self.n_exceptions)
        # longest-first block order of the chain kernel (scratch of the chain-local step; must start zeroed)
        self.block_order = (z((L.ORDER_BUCKETS * Cn + 128,), i32)
                            if any(x is not None for x in self.idx_cm) and not os.environ.get("BBB_NO_BLOCK_ORDER") else None)
        b.block_order = _ptr(self.block_order)
        self.tape = None
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.float64)
            if tape.ndim != 2 or tape.shape[0] != Cn:
                raise ValueError("tape must be [n_chains][length]")
            self.tape = torch.as_tensor(tape).to(dev)
            b.tape, b.tape_stride = _ptr(self.tape), tape.shape[1]
        self._cbuf = b

    # ------------------------------------------------------------------ driving
    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def init_states(self):
        L.check(self.lib.bbb_engine_init_states(self._handle, self._stream))

    def refresh(self):
        L.check(self.lib.bbb_engine_refresh(self._handle, self._stream))
        self._steps_since_sync = 1 << 30  # the engine forgot its bound: tighten before the next step

    RESYNC_EVERY = 256   # chain-local engines: steps between two from-scratch evaluations of residuals and misfit
    K_BOUND_RESYNC = 64  # steps between two tightenings of the engine's bound on max k (one tiny D2H sync each)
    K_BOUND_RESYNC_LOCAL = 1024  # chain-local engines only size their rare from-scratch evaluations by the bound

    def step(self, n_steps: int = 1):
        n_steps = int(n_steps)
        every = self.K_BOUND_RESYNC_LOCAL if self.chain_local else self.K_BOUND_RESYNC
        while n_steps > 0:
            if self._steps_since_sync >= every:
                self.sync_k_bound()
            n = min(n_steps, every - self._steps_since_sync)
            L.check(self.lib.bbb_engine_step(self._handle, n, self._stream))
            self._steps_since_sync += n
            n_steps -= n

    def sync_k_bound(self):
        """Tell the engine the true maximum of k over the resident chains (it only knows an upper bound that
        grows by one per step); shared-memory staging of the tomography kernels is sized by it."""
        self._steps_since_sync = 0
        if not self._has_ray:
            return
        L.check(self.lib.bbb_engine_sync_k_bound(self._handle, None, self._stream))  # (one reduction kernel + a 16-byte read-back)

    def stage(self, stage: int):
        """One stage of one step (0 propose, 1 raster, 2 ray forward + misfit, 3 accept)."""
        if stage == 0:
            if self._steps_since_sync >= (self.K_BOUND_RESYNC_LOCAL if self.chain_local else self.K_BOUND_RESYNC):
                self.sync_k_bound()
            self._steps_since_sync += 1
        L.check(self.lib.bbb_engine_stage(self._handle, int(stage), self._stream))

    def set_annealing(self, temperature_start, cooling_rate, cooling_iterations, burnin_iterations):
        torch.cuda.synchronize(self.device)
        L.check(self.lib.bbb_engine_set_annealing(self._handle, float(temperature_start), float(cooling_rate),
                                                  int(cooling_iterations), int(burnin_iterations)))

    def gather_space(self, s: int):
        sp = self.spec["spaces"][s]
        K, nd, P = sp["kmax"], sp["ndim"], len(sp["params"])
        k = torch.empty((self.C,), dtype=torch.int32, device=self.device)
        sites = torch.empty((nd, K, self.C), dtype=torch.float64, device=self.device) if nd else None
        vals = torch.empty((P, K, self.C), dtype=torch.float64, device=self.device) if P else None
        L.check(self.lib.bbb_engine_gather_space(self._handle, s, _ptr(k), _ptr(sites), _ptr(vals), self._stream))
        return k, sites, vals

    def save_point_bytes(self, kb, want_dpred: bool) -> int:
        return int(self.lib.bbb_engine_save_point_bytes(self._handle, self._kb(kb), int(bool(want_dpred))))

    def save_point(self, kb, want_dpred: bool, image: torch.Tensor, scratch: Optional[torch.Tensor] = None):
        """The current state of every chain (and d_pred) as one chain-major device image (include/bbb200.h)."""
        L.check(self.lib.bbb_engine_save_point(self._handle, self._kb(kb), int(bool(want_dpred)), _ptr(image), _ptr(scratch),
                                               self._stream))

    def dpred(self, t: int) -> torch.Tensor:
        R = int(np.asarray(self.spec["targets"][t]["dobs"]).size)
        out = torch.empty((R, self.C), dtype=torch.float64, device=self.device)
        L.check(self.lib.bbb_engine_dpred(self._handle, t, _ptr(out), self._stream))
        return out

    def dpred_chain_major(self, t: int) -> torch.Tensor:
        """[C][n_data] d_pred of the current states.  Chain-local engines hold the residual ``d_pred - d_obs`` of
        the current states resident (re-based on a full evaluation every ``RESYNC_EVERY`` steps), so this is one
        addition; other engines evaluate the forward."""
        if self.chain_local and self.res[t] is not None:
            dobs = torch.as_tensor(np.asarray(self.spec["targets"][t]["dobs"], dtype=np.float64)).to(self.device)
            return self.res[t] + dobs[None, :]
        return self.dpred(t).t().contiguous()

    def misfit_logdet(self):
        m = torch.empty((self.C,), dtype=torch.float64, device=self.device)
        ld = torch.empty((self.C,), dtype=torch.float64, device=self.device)
        L.check(self.lib.bbb_engine_misfit_logdet(self._handle, _ptr(m), _ptr(ld), self._stream))
        return m, ld

    def misfit_logdet_T(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[C][3] rows (misfit, logdet, T) of the current states, written by one kernel: what this rank contributes
        to a parallel-tempering swap round."""
        if out is None:
            out = torch.empty((self.C, 3), dtype=torch.float64, device=self.device)
        L.check(self.lib.bbb_engine_swap_rows(self._handle, _ptr(out), self._stream))
        return out

    def ray_tiles(self, t: int) -> int:
        """Ray tiles of the chain-local step of ray target ``t`` (1: all rays' accumulators fit one block)."""
        return max(1, int(self._cspec.targets[t].csc_n_tiles))

    def resync(self):
        """Chain-local engines: re-evaluate forward and misfit of the current states from scratch (the engine also
        does this on its own every ``bbb_engine_set_resync`` steps)."""
        L.check(self.lib.bbb_engine_resync(self._handle, self._stream))

    def current_cell_idx(self, t: int) -> torch.Tensor:
        """[G][C] nearest-site index of the current states of ray target ``t``."""
        inv = self.point_perm[t][1]
        if self.chain_local:
            G = self.cell_idx[t].shape[1]
            return self.idx_cm[t][:, :G].t().index_select(0, inv).contiguous()
        sel = self.idx_sel[t].long()
        idx = self.cell_idx[t]
        return torch.where(sel[None, :] == 0, idx[0], idx[1]).index_select(0, inv)

    # ------------------------------------------------------------------ host <-> device state
    def load_states(self, states: List[dict], temperatures=None):
        """Write starting states.  ``states[c]`` = ``{"spaces": [{"k", "sites", "values": [..]}],
        "sigma": [..]}`` (numpy).  Resets slots and re-evaluates the forward of every chain."""
        if len(states) != self.C:
            raise ValueError(f"expected {self.C} starting states, got {len(states)}")
        for s, sp in enumerate(self.spec["spaces"]):
            K, nd, P = sp["kmax"], sp["ndim"], len(sp["params"])
            k = np.array([st["spaces"][s]["k"] for st in states], dtype=np.int32)
            if (k < sp["kmin"]).any() or (k > sp["kmax"]).any():
                raise ValueError(f"starting state: n_dimensions outside [{sp['kmin']}, {sp['kmax']}]")
            if nd:
                sites = np.zeros((nd, K, self.C))
                for c, st in enumerate(states):
                    sites[:, : k[c], c] = np.asarray(st["spaces"][s]["sites"], dtype=np.float64).reshape(k[c], nd).T
                self.sites[s][0].copy_(torch.as_tensor(sites))
            if P:
                vals = np.zeros((P, K, self.C))
                for c, st in enumerate(states):
                    for p in range(P):
                        vals[p, : k[c], c] = st["spaces"][s]["values"][p]
                self.values[s][0].copy_(torch.as_tensor(vals))
            self.k[s].copy_(torch.as_tensor(np.stack((k, k))))
            self.sel[s].zero_()
        for t, tg in enumerate(self.spec["targets"]):
            if tg["hierarchical"]:
                sg = np.array([st["sigma"][t] for st in states], dtype=np.float64)
                if (sg < tg["std_min"]).any() or (sg > tg["std_max"]).any():
                    raise ValueError(f"starting state: noise std of {tg['name']!r} outside [std_min, std_max]")
                self.sigma[t].copy_(torch.as_tensor(np.stack((sg, sg))))
            if tg.get("correlated"):
                rr = np.array([st["corr"][t] for st in states], dtype=np.float64)
                self.corr[t].copy_(torch.as_tensor(np.stack((rr, rr))))
            if self.idx_sel[t] is not None:
                self.idx_sel[t].zero_()
        if temperatures is not None:
            self.temperature.copy_(torch.as_tensor(np.asarray(temperatures, dtype=np.float64)))
        self.refresh()

    def fetch_states(self) -> List[dict]:
        """Current state of every chain as host dicts (inverse of :meth:`load_states`)."""
        out = [dict(spaces=[], sigma=[], corr=[]) for _ in range(self.C)]
        for s, sp in enumerate(self.spec["spaces"]):
            k, sites, vals = self.gather_space(s)
            k = k.cpu().numpy()
            sites = None if sites is None else sites.cpu().numpy()
            vals = None if vals is None else vals.cpu().numpy()
            for c in range(self.C):
                kc = int(k[c])
                st = dict(k=kc, sites=None, values=[])
                if sites is not None:
                    a = sites[:, :kc, c].T.copy()
                    st["sites"] = a[:, 0] if sp["ndim"] == 1 else a
                if vals is not None:
                    st["values"] = [vals[p, :kc, c].copy() for p in range(vals.shape[0])]
                out[c]["spaces"].append(st)
        for t, tg in enumerate(self.spec["targets"]):
            sg = self.sigma[t][0].cpu().numpy() if tg["hierarchical"] else None
            rr = self.corr[t][0].cpu().numpy() if tg.get("correlated") else None
            for c in range(self.C):
                out[c]["sigma"].append(None if sg is None else float(sg[c]))
                out[c]["corr"].append(None if rr is None else float(rr[c]))
        return out

    # ------------------------------------------------------------------ packed chain-state image
    def _kb(self, kb):
        n = len(self.spec["spaces"])
        kb = [int(kb)] * n if np.isscalar(kb) else [int(v) for v in kb]
        return (C.c_int32 * n)(*[min(max(v, 1), sp["kmax"]) for v, sp in zip(kb, self.spec["spaces"])])

    def state_image_bytes(self, kb) -> int:
        """Bytes of the packed image of all chains' current states when ``kb`` cell rows per space travel."""
        return int(self.lib.bbb_engine_state_image_bytes(self._handle, self._kb(kb)))

    def pack_states(self, kb, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Current state of every chain as ONE device buffer (int64 words; see include/bbb200.h)."""
        n = self.state_image_bytes(kb) // 8
        if out is None:
            out = torch.empty((n,), dtype=torch.int64, device=self.device)
        L.check(self.lib.bbb_engine_pack_states(self._handle, self._kb(kb), _ptr(out[:n]), self._stream))
        return out[:n]

    def states_differ(self, kb, image: torch.Tensor, flag: torch.Tensor) -> torch.Tensor:
        """Sets ``flag`` (device int32, zeroed by the caller) if some chain's state is not bit-identical to ``image``."""
        L.check(self.lib.bbb_engine_compare_states(self._handle, self._kb(kb), _ptr(image), _ptr(flag), self._stream))
        return flag

    def unpack_states(self, kb, image: torch.Tensor):
        """Write a packed image into the engine and re-derive rasterisation, forward and misfit."""
        L.check(self.lib.bbb_engine_unpack_states(self._handle, self._kb(kb), _ptr(image), self._stream))
        self._steps_since_sync = 1 << 30

    # ------------------------------------------------------------------ hosted step (chains held by the caller)
    def hosted_parts(self):
        """First chains of the parts of a hosted image (``len - 1`` parts; see include/bbb200.h)."""
        first = (C.c_int64 * 5)()
        n = int(self.lib.bbb_engine_hosted_parts(self._handle, first))
        if n < 1:
            raise RuntimeError("bbb_engine_hosted_parts failed")
        return [int(first[h]) for h in range(n + 1)]

    def pack_states_hosted(self, kb, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Current states as a HOSTED image: the parts' images back to back (device buffer of int64 words)."""
        n = self.state_image_bytes(kb) // 8
        rows = n // self.C
        if out is None:
            out = torch.empty((n,), dtype=torch.int64, device=self.device)
        first = self.hosted_parts()
        for c0, c1 in zip(first[:-1], first[1:]):
            if c1 > c0:
                L.check(self.lib.bbb_engine_pack_states_range(self._handle, self._kb(kb), c0, c1, _ptr(out[rows * c0:]), self._stream))
        return out[:n]

    def step_hosted(self, kb_in, host_in: torch.Tensor, host_out: torch.Tensor, stage: torch.Tensor, out: torch.Tensor):
        """Enqueue ONE iteration on the chains of the pinned hosted image ``host_in`` (``kb_in`` cell rows); the new
        states arrive in ``host_out`` with the returned number of cell rows once :meth:`step_hosted_wait` returned."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise ValueError("hosted images must live in pinned host memory")
        if self._steps_since_sync >= self.K_BOUND_RESYNC_LOCAL:
            self.sync_k_bound()
        key = (tuple(kb_in) if not np.isscalar(kb_in) else int(kb_in), host_in.numel(), host_out.numel(), stage.numel(), out.numel())
        cached = self._hosted_args.get(key) if hasattr(self, "_hosted_args") else None
        if cached is None:
            if not hasattr(self, "_hosted_args"):
                self._hosted_args = {}
            kb_c = self._kb(kb_in)
            need_in = self.state_image_bytes(kb_c) // 8
            kb_cap = self._kb([min(int(v) + 8, sp["kmax"]) for v, sp in zip(kb_c, self.spec["spaces"])])
            need_out = self.state_image_bytes(kb_cap) // 8
            if host_in.numel() < need_in or stage.numel() < need_in or host_out.numel() < need_out or out.numel() < need_out:
                raise ValueError("hosted image buffers too small")
            cached = self._hosted_args[key] = (kb_c, (C.c_int32 * len(self.spec["spaces"]))())
        kb_in, kb_out = cached
        L.check(self.lib.bbb_engine_step_hosted(self._handle, kb_in, C.c_void_p(host_in.data_ptr()), kb_out,
                                                C.c_void_p(host_out.data_ptr()), _ptr(stage), _ptr(out), self._stream))
        self._steps_since_sync += 1
        return [int(v) for v in kb_out]

    # ---- ragged hosted image: only the cells in use travel (include/bbb200.h)
    def hosted_ragged_words(self) -> int:
        """Capacity of a ragged hosted image in 8-byte words (all four buffers of :meth:`step_hosted_ragged`)."""
        n = int(self.lib.bbb_engine_hosted_ragged_bytes(self._handle))
        if n < 0:
            raise RuntimeError("bbb_engine_hosted_ragged_bytes failed")
        return n // 8

    def pack_states_ragged(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Current states of all chains as a ragged hosted image (device buffer of int64 words; words behind a part's
        content are left as they are)."""
        if out is None:
            out = torch.zeros((self.hosted_ragged_words(),), dtype=torch.int64, device=self.device)
        L.check(self.lib.bbb_engine_pack_states_ragged(self._handle, _ptr(out), self._stream))
        return out

    def ragged_parts(self, image: torch.Tensor):
        """The content of a ragged hosted image (host or device tensor of int64 words): per part ``(c0, c1, words)``
        with ``words`` the slice the part's header and cells occupy."""
        first = self.hosted_parts()
        stride = self.hosted_ragged_words() // (len(first) - 1)
        n_noise = sum((2 if tg.get("correlated") else 1) for tg in self.spec["targets"] if tg["hierarchical"])
        hr = len(self.spec["spaces"]) + n_noise + 2
        out = []
        for h, (c0, c1) in enumerate(zip(first[:-1], first[1:])):
            n_c = c1 - c0
            words = hr * n_c
            for s, sp in enumerate(self.spec["spaces"]):
                k = image[stride * h + s * n_c: stride * h + (s + 1) * n_c]
                words += int(k.sum().item()) * (sp["ndim"] + len(sp["params"]))
            out.append((c0, c1, image[stride * h: stride * h + words]))
        return out

    def step_hosted_ragged(self, host_in: torch.Tensor, host_out: torch.Tensor, stage: torch.Tensor, out: torch.Tensor):
        """Enqueue ONE iteration on the chains of the pinned ragged hosted image ``host_in``; the new states arrive in
        ``host_out`` once :meth:`step_hosted_wait` returned.  Returns the bytes copied (host -> device, device -> host)."""
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise ValueError("hosted images must live in pinned host memory")
        need = self.hosted_ragged_words()
        if min(host_in.numel(), host_out.numel(), stage.numel(), out.numel()) < need:
            raise ValueError("hosted image buffers too small")
        if self._steps_since_sync >= self.K_BOUND_RESYNC_LOCAL:
            self.sync_k_bound()
        up, down = C.c_int64(0), C.c_int64(0)
        L.check(self.lib.bbb_engine_step_hosted_ragged(self._handle, C.c_void_p(host_in.data_ptr()), C.c_void_p(host_out.data_ptr()),
                                                       _ptr(stage), _ptr(out), C.byref(up), C.byref(down), self._stream))
        self._steps_since_sync += 1
        return int(up.value), int(down.value)

    # ---- mapped hosted step: fixed-stride records in pinned host memory, moved by the chain kernel itself
    def hosted_record_words(self) -> int:
        n = getattr(self, "_record_words", None)
        if n is None:
            n = int(self.lib.bbb_engine_hosted_record_words(self._handle))
            if n < 0:
                raise RuntimeError("bbb_engine_hosted_record_words failed")
            self._record_words = n
        return n

    def pack_states_records(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Current states of all chains as records ``[C][hosted_record_words()]`` of int64 words (device tensor, or a
        pinned host tensor the kernel writes directly; words behind a chain's cells in use are left as they are)."""
        if out is None:
            out = torch.zeros((self.C, self.hosted_record_words()), dtype=torch.int64, device=self.device)
        L.check(self.lib.bbb_engine_pack_states_records(self._handle, C.c_void_p(out.data_ptr()), self._stream))
        return out

    def decode_records(self, image: torch.Tensor):
        """Per chain ``dict(k=[..], cells=[array [k][ndim + n_params] per space], header=int64 words)`` from a record
        image on the host (what a caller holding the chains reads)."""
        rw = self.hosted_record_words()
        img = image.view(self.C, rw)
        n_sp = len(self.spec["spaces"])
        n_noise = sum((2 if tg.get("correlated") else 1) for tg in self.spec["targets"] if tg["hierarchical"])
        hrw = n_sp + n_noise + 2
        out, base = [], hrw
        bases = []
        for sp in self.spec["spaces"]:
            bases.append(base)
            base += sp["kmax"] * (sp["ndim"] + len(sp["params"]))
        for c in range(self.C):
            rec = img[c]
            ks = [int(rec[s]) for s in range(n_sp)]
            cells = []
            for s, sp in enumerate(self.spec["spaces"]):
                w = sp["ndim"] + len(sp["params"])
                cells.append(rec[bases[s]: bases[s] + ks[s] * w].view(torch.float64).numpy().reshape(ks[s], w).copy())
            out.append(dict(k=ks, cells=cells, header=rec[:hrw].clone()))
        return out

    def step_hosted_mapped(self, host_in: torch.Tensor, host_out: torch.Tensor):
        """Enqueue ONE iteration on the chains whose records are in the pinned host tensor ``host_in``; the chain kernel
        reads them from and writes the new records to host memory (``host_out``) itself."""
        a, b = host_in.data_ptr(), host_out.data_ptr()
        checked = self.__dict__.setdefault("_mapped_checked", set())
        if (a, b) not in checked:  # (buffers seen before were checked then; the library checks the pointers on every call)
            if not (host_in.is_pinned() and host_out.is_pinned()):
                raise ValueError("hosted images must live in pinned host memory")
            if min(host_in.numel(), host_out.numel()) < self.C * self.hosted_record_words():
                raise ValueError("hosted image buffers too small")
            checked.add((a, b))
        if self._steps_since_sync >= self.K_BOUND_RESYNC_LOCAL:
            self.sync_k_bound()
        L.check(self.lib.bbb_engine_step_hosted_mapped(self._handle, a, b, self._stream))
        self._steps_since_sync += 1

    def step_hosted_mapped_wait(self) -> int:
        """Block until the mapped hosted step is done; returns the number of chains that had to be redone (0: fast path)."""
        n = C.c_int32(0)
        L.check(self.lib.bbb_engine_step_hosted_mapped_wait(self._handle, C.byref(n), self._stream))
        if n.value:
            self._steps_since_sync = 1 << 30
        return int(n.value)

    def hosted_mapped_bytes(self):
        """Bytes the mapped hosted steps have read from / written to host memory so far."""
        up, down = C.c_int64(0), C.c_int64(0)
        L.check(self.lib.bbb_engine_hosted_mapped_bytes(self._handle, C.byref(up), C.byref(down)))
        return int(up.value), int(down.value)

    def step_hosted_wait(self) -> int:
        """Wait for the hosted step; returns the number of parts whose caches had to be re-derived (0: fast path)."""
        n = C.c_int32(0)
        L.check(self.lib.bbb_engine_step_hosted_wait(self._handle, C.byref(n), self._stream))
        if n.value:
            self._steps_since_sync = 1 << 30
        return int(n.value)

    # ------------------------------------------------------------------ checkpoint / resume
    def _mutable(self):
        named = {"temperature": self.temperature, "iteration": self.iteration, "n_proposed": self.n_proposed,
                 "n_accepted": self.n_accepted, "n_exceptions": self.n_exceptions, "draw_cursor": self.draw_cursor, "move": self.move, "edit": self.edit,
                 "log_prob_ratio": self.log_prob_ratio, "log_like_ratio": self.log_like_ratio, "accepted": self.accepted}
        for i in range(len(self.spec["spaces"])):
            named.update({f"sites{i}": self.sites[i], f"values{i}": self.values[i], f"k{i}": self.k[i], f"sel{i}": self.sel[i]})
        for i in range(len(self.spec["targets"])):
            named.update({f"sigma{i}": self.sigma[i], f"corr{i}": self.corr[i], f"phi{i}": self.phi[i],
                          f"cell_idx{i}": self.cell_idx[i], f"idx_sel{i}": self.idx_sel[i],
                          f"idx_cm{i}": self.idx_cm[i], f"res{i}": self.res[i]})
        return {k: v for k, v in named.items() if v is not None}

    def state_dict(self) -> dict:
        """Everything needed to continue the chains bit-identically (the Philox counters are the per-chain
        iteration numbers, so no generator state has to be saved)."""
        torch.cuda.synchronize(self.device)
        out = {k: v.cpu() for k, v in self._mutable().items()}
        out["_meta"] = dict(n_chains=self.C, seed=self.seed, chain_id0=self.chain_id0,
                            steps_since_resync=int(self.lib.bbb_engine_steps_since_resync(self._handle)))
        return out

    def load_state_dict(self, state: dict):
        meta = state["_meta"]
        if meta["n_chains"] != self.C:
            raise ValueError(f"checkpoint holds {meta['n_chains']} chains, this engine {self.C}")
        if meta["seed"] != self.seed or meta["chain_id0"] != self.chain_id0:
            raise ValueError("checkpoint was written with a different seed / chain shard")
        for k, v in self._mutable().items():
            if tuple(state[k].shape) != tuple(v.shape) or state[k].dtype != v.dtype:
                raise ValueError(f"checkpoint buffer {k!r} does not match this problem")
            v.copy_(state[k])
        for s, sp in enumerate(self.spec["spaces"]):
            L.check(self.lib.bbb_engine_set_k_bound(self._handle, s, sp["kmax"]))
        L.check(self.lib.bbb_engine_set_resync(self._handle, self.RESYNC_EVERY, int(meta.get("steps_since_resync", 0))))
        self._steps_since_sync = 1 << 30
        torch.cuda.synchronize(self.device)

    def close(self):
        if self._handle:
            self.lib.bbb_engine_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pt_swap(gathered: torch.Tensor, seed: int, round_: int, id0: int, n_local: int,
            tape: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """Run one parallel-tempering swap round on ``gathered`` [n_total][3] (misfit, logdet, T).
    Returns ``(local temperatures [n_local], number of accepted swaps)``; ``out`` (e.g. the engine's temperature
    buffer) receives the local temperatures in place of a fresh tensor."""
    lib = L.load()
    n_total = gathered.shape[0]
    assert gathered.is_contiguous() and gathered.dtype == torch.float64
    t_out = torch.empty((n_local,), dtype=torch.float64, device=gathered.device) if out is None else out
    assert t_out.is_contiguous() and t_out.dtype == torch.float64 and t_out.numel() >= n_local
    n_acc = torch.zeros((1,), dtype=torch.int32, device=gathered.device)
    stream = C.c_void_p(torch.cuda.current_stream(gathered.device).cuda_stream)
    L.check(lib.bbb_pt_swap(_ptr(gathered), n_total, seed & 0xFFFFFFFFFFFFFFFF, int(round_), int(id0), int(n_local),
                            _ptr(t_out), _ptr(n_acc), _ptr(tape), stream))
    return t_out, n_acc
