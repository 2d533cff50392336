"""ctypes binding of ``libbbb200.so`` (the C ABI declared in ``include/bbb200.h``).

The CUDA library is the product; there is no CPU fallback.  Importing this module fails
loudly when the shared library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C bayesbay_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

MAX_SPACES = 4
MAX_PARAMS = 8
MAX_TARGETS = 4
MAX_LINEAR_PARAMS = 16
CSC_LEVELS = 3
ABI_VERSION = 6
ORDER_BUCKETS = 10

SPACE_PLAIN, SPACE_VORONOI1D, SPACE_VORONOI2D = 0, 1, 2
PRIOR_UNIFORM, PRIOR_GAUSSIAN, PRIOR_LAPLACE = 0, 1, 2
BIRTH_FROM_PRIOR, BIRTH_FROM_NEIGHBOUR = 0, 1
FWD_RAY2D, FWD_LOOKUP1D, FWD_LINEAR, FWD_POLYNOMIAL = 0, 1, 2, 3
TRANSFORM_IDENTITY, TRANSFORM_RECIPROCAL = 0, 1
COV_HIERARCHICAL, COV_SCALAR, COV_VECTOR, COV_MATRIX = 0, 1, 2, 3
MOVE_BIRTH, MOVE_DEATH, MOVE_VALUES, MOVE_SITES = 0, 1, 2, 3

_vp = C.c_void_p


class SpaceSpec(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("ndim", C.c_int32), ("trans_d", C.c_int32),
        ("kmin", C.c_int32), ("kmax", C.c_int32), ("kinit_max", C.c_int32),
        ("birth_from", C.c_int32), ("n_params", C.c_int32),
        ("vmin", C.c_double * 2), ("vmax", C.c_double * 2), ("site_std", C.c_double * 2),
        ("w_birth", C.c_double), ("w_death", C.c_double), ("w_values", C.c_double), ("w_sites", C.c_double),
        ("w_outer", C.c_double),
        ("prior_kind", C.c_int32 * MAX_PARAMS),
        ("prior_a", C.c_double * MAX_PARAMS), ("prior_b", C.c_double * MAX_PARAMS),
        ("perturb_std", C.c_double * MAX_PARAMS), ("perturb_std_birth", C.c_double * MAX_PARAMS),
        ("n_pos", C.c_int32 * MAX_PARAMS), ("pos_table", _vp * MAX_PARAMS),
    ]


class TargetSpec(C.Structure):
    _fields_ = [
        ("n_data", C.c_int32), ("cov_kind", C.c_int32),
        ("std_min", C.c_double), ("std_max", C.c_double), ("std_perturb_std", C.c_double),
        ("correlated", C.c_int32),
        ("corr_min", C.c_double), ("corr_max", C.c_double), ("corr_perturb_std", C.c_double),
        ("cov_scalar", C.c_double),
        ("dobs", _vp), ("cov_vec", _vp),
        ("fwd_kind", C.c_int32), ("space", C.c_int32), ("param", C.c_int32), ("transform", C.c_int32),
        ("n_points", C.c_int32), ("nnz", C.c_int64),
        ("indptr", _vp), ("indices", _vp), ("data", _vp), ("points_x", _vp), ("points_y", _vp),
        ("x", _vp),
        ("n_lin_params", C.c_int32), ("lin_params", C.c_int32 * MAX_LINEAR_PARAMS),
        ("G", _vp),
        ("csc_tile_rays", C.c_int32), ("csc_n_tiles", C.c_int32), ("csc_n_levels", C.c_int32),
        ("csc_colptr", _vp * CSC_LEVELS), ("csc_rows", _vp * CSC_LEVELS), ("csc_data", _vp * CSC_LEVELS),
        ("csc_bound", C.c_double),
        ("group_box", _vp),
    ]


class ProblemSpec(C.Structure):
    _fields_ = [
        ("n_spaces", C.c_int32), ("n_targets", C.c_int32),
        ("spaces", SpaceSpec * MAX_SPACES), ("targets", TargetSpec * MAX_TARGETS),
        ("seed", C.c_uint64),
        ("w_noise", C.c_double),
    ]


class Buffers(C.Structure):
    _fields_ = [
        ("n_chains", C.c_int64), ("chain_id0", C.c_int64),
        ("sites", _vp * MAX_SPACES), ("values", _vp * MAX_SPACES), ("k", _vp * MAX_SPACES), ("sel", _vp * MAX_SPACES),
        ("sigma", _vp * MAX_TARGETS), ("corr", _vp * MAX_TARGETS), ("phi", _vp * MAX_TARGETS),
        ("dpred_scratch", _vp * MAX_TARGETS), ("cell_idx", _vp * MAX_TARGETS),
        ("idx_sel", _vp * MAX_TARGETS), ("partial", _vp * MAX_TARGETS),
        ("rescan_list", _vp * MAX_TARGETS), ("rescan_count", _vp * MAX_TARGETS),
        ("idx_cm", _vp * MAX_TARGETS), ("res", _vp * MAX_TARGETS), ("change_list", _vp * MAX_TARGETS),
        ("temperature", _vp), ("iteration", _vp),
        ("move", _vp), ("edit", _vp), ("edit_site", _vp), ("edit_vals", _vp), ("log_prob_ratio", _vp), ("log_like_ratio", _vp), ("accepted", _vp),
        ("n_proposed", _vp), ("n_accepted", _vp),
        ("n_exceptions", _vp),
        ("tape", _vp), ("tape_stride", C.c_int64), ("tape_cursor", _vp),
        ("block_order", _vp),
    ]


# every symbol include/bbb200.h declares: name -> (restype, argtypes)
_i32, _i64, _u64 = C.c_int32, C.c_int64, C.c_uint64
SYMBOLS = {
    "bbb_abi_version": (C.c_int, []),
    "bbb_last_error": (C.c_char_p, []),
    "bbb_ray_tiles": (_i64, [_i64]),
    "bbb_idx_row_stride": (_i64, [_i64]),
    "bbb_chain_local_tile_rays": (_i64, [_i64, _i32, _i64]),
    "bbb_engine_chain_local": (C.c_int, [_vp]),
    "bbb_engine_set_resync": (C.c_int, [_vp, _i64, _i64]),
    "bbb_engine_steps_since_resync": (_i64, [_vp]),
    "bbb_engine_resync": (C.c_int, [_vp, _vp]),
    "bbb_engine_create": (C.c_int, [C.POINTER(ProblemSpec), C.c_int, C.POINTER(_vp)]),
    "bbb_engine_bind_buffers": (C.c_int, [_vp, C.POINTER(Buffers)]),
    "bbb_engine_destroy": (None, [_vp]),
    "bbb_engine_init_states": (C.c_int, [_vp, _vp]),
    "bbb_engine_refresh": (C.c_int, [_vp, _vp]),
    "bbb_engine_set_k_bound": (C.c_int, [_vp, C.c_int, _i32]),
    "bbb_engine_step": (C.c_int, [_vp, _i64, _vp]),
    "bbb_engine_stage": (C.c_int, [_vp, C.c_int, _vp]),
    "bbb_engine_set_annealing": (C.c_int, [_vp, C.c_double, C.c_double, _i64, _i64]),
    "bbb_engine_gather_space": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "bbb_engine_save_point_bytes": (_i64, [_vp, C.POINTER(_i32), _i32]),
    "bbb_engine_save_point": (C.c_int, [_vp, C.POINTER(_i32), _i32, _vp, _vp, _vp]),
    "bbb_engine_state_image_bytes": (_i64, [_vp, C.POINTER(_i32)]),
    "bbb_engine_pack_states": (C.c_int, [_vp, C.POINTER(_i32), _vp, _vp]),
    "bbb_engine_compare_states": (C.c_int, [_vp, C.POINTER(_i32), _vp, _vp, _vp]),
    "bbb_engine_unpack_states": (C.c_int, [_vp, C.POINTER(_i32), _vp, _vp]),
    "bbb_engine_pack_states_range": (C.c_int, [_vp, C.POINTER(_i32), _i64, _i64, _vp, _vp]),
    "bbb_engine_hosted_parts": (C.c_int, [_vp, C.POINTER(_i64)]),
    "bbb_engine_step_hosted": (C.c_int, [_vp, C.POINTER(_i32), _vp, C.POINTER(_i32), _vp, _vp, _vp, _vp]),
    "bbb_engine_step_hosted_wait": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "bbb_engine_sync_k_bound": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "bbb_engine_hosted_ragged_bytes": (_i64, [_vp]),
    "bbb_engine_hosted_record_words": (_i64, [_vp]),
    "bbb_engine_pack_states_records": (C.c_int, [_vp, _vp, _vp]),
    "bbb_engine_step_hosted_mapped": (C.c_int, [_vp, _vp, _vp, _vp]),
    "bbb_engine_step_hosted_mapped_wait": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "bbb_engine_hosted_mapped_bytes": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "bbb_engine_pack_states_ragged": (C.c_int, [_vp, _vp, _vp]),
    "bbb_engine_step_hosted_ragged": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "bbb_engine_dpred": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "bbb_engine_misfit_logdet": (C.c_int, [_vp, _vp, _vp, _vp]),
    "bbb_engine_swap_rows": (C.c_int, [_vp, _vp, _vp]),
    "bbb_pt_swap": (C.c_int, [_vp, _i64, _u64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "bbb_raster2d": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp]),
    "bbb_raster1d": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _vp, _vp]),
    "bbb_ray_forward": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp,
                                  _vp, _vp]),
    "bbb_dense_forward": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, _vp]),
    "bbb_loglike": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "bbb_logprior": (C.c_int, [C.POINTER(SpaceSpec), _vp, _vp, _i64, _vp, _vp]),
    "bbb_philox_kat": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

# (BBB_LIB_PATH: developer override, e.g. an instrumented build of the same sources)
LIB_PATH = os.environ.get("BBB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbbb200.so")
_lib = None


class BBBError(RuntimeError):
    pass


def load():
    """Load libbbb200.so (once) and set the prototypes.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (make -C bayesbay_b200/csrc, or "
            "__graft_entry__.build()).  bayesbay_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bbb_abi_version() != ABI_VERSION:
        raise ImportError(f"libbbb200.so ABI {lib.bbb_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    """Turn a negative bbb_status into the Python exception the reference API would raise."""
    if rc == 0:
        return
    msg = load().bbb_last_error().decode("utf-8", "replace")
    if rc in (-1, -2, -4):
        raise ValueError(f"libbbb200: {msg}")
    raise BBBError(f"libbbb200 (status {rc}): {msg}")
