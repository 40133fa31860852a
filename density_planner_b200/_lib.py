"""ctypes binding of libdensity_planner_b200.so (the C ABI declared in include/density_planner_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a DensityPlannerError is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdensity_planner_b200.so")
DIAG_LIB_PATH = os.path.join(_HERE, "libdensity_planner_b200_diag.so")
CSRC = os.path.join(_HERE, "csrc")

# flags (include/density_planner_b200.h)
DP_LOG_DENSITY, DP_DENSITY, DP_CUT, DP_ABS_STATE, DP_XY_ONLY = 1, 2, 4, 8, 16
DP_CONTROLLER_BLOB_FLOATS = 43592


class DensityPlannerError(RuntimeError):
    pass


class GridGeom(ctypes.Structure):
    _fields_ = [("x_min", ctypes.c_float), ("x_max", ctypes.c_float), ("y_min", ctypes.c_float),
                ("y_max", ctypes.c_float), ("gx", ctypes.c_int), ("gy", ctypes.c_int)]


class BinSpec(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("geom", GridGeom), ("lo_x", ctypes.c_double), ("hi_x", ctypes.c_double),
                ("lo_y", ctypes.c_double), ("hi_y", ctypes.c_double), ("bx", ctypes.c_int), ("by", ctypes.c_int),
                ("mass_scale_log2", ctypes.c_int)]


_vp, _i64, _int, _uint, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint, ctypes.c_float

_SIGNATURES = {
    "dp_abi_version": (_int, []),
    "dp_last_error": (ctypes.c_char_p, []),
    "dp_device_ok": (_int, [_int]),
    "dp_controller_create": (_int, [_vp, ctypes.c_size_t, _int, ctypes.POINTER(_vp)]),
    "dp_controller_destroy": (None, [_vp]),
    "dp_rollout": (_int, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _int, _uint, _int, _vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_rollout_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i64, _int, _uint, _int, _vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_rollout_host": (_int, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _int, _uint, _int, _vp, _vp, _vp, _vp]),
    "dp_host_chunk_schedule": (_int, [_i64, _i64, _vp, _int]),
    "dp_step": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dp_pos2gridpos": (_int, [ctypes.POINTER(GridGeom), _vp, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "dp_pos2gridpos_f64": (_int, [ctypes.POINTER(GridGeom), _vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_gridpos2pos": (_int, [ctypes.POINTER(GridGeom), _vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_env_create": (_int, [ctypes.POINTER(GridGeom), _int, _vp, _vp, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "dp_env_destroy": (None, [_vp]),
    "dp_cost_coll": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _i64,
                            _i64, _vp, _vp]),
    "dp_bin2d": (_int, [ctypes.POINTER(BinSpec), _vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]),
    "dp_bin2d_occ": (_int, [ctypes.POINTER(BinSpec), _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "dp_xref_traj": (_int, [_vp, _int, _vp, _f32, _int, _int, _vp, _vp]),
    "dp_xref_traj_bwd": (_int, [_vp, _vp, _f32, _int, _int, _vp, _vp, _vp]),
    "dp_cost_state": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "dp_cost_state_bwd": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _f32, _f32, _f32, _vp,
                                 _i64, _i64, _i64, _vp, _i64, _i64, _vp]),
    "dp_normalize_stats": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp]),
    "dp_normalize_apply": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp, _vp]),
    "dp_pipeline_create": (_int, [_vp, _vp, ctypes.POINTER(BinSpec), _int, _int, _i64, _int, ctypes.POINTER(_vp)]),
    "dp_pipeline_destroy": (None, [_vp]),
    "dp_pipeline_words": (_int, [_vp, ctypes.POINTER(_i64)]),
    "dp_pipeline_rollout": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _f32, _i64, _vp, _vp]),
    "dp_pipeline_bin": (_int, [_vp, _vp, _vp, _vp]),
    "dp_pipeline_finalize": (_int, [_vp, _vp, _vp, _vp]),
    "dp_pipeline_copy_slices": (_int, [_vp, _vp, _vp, _vp]),
    "dp_rollout_fused": (_int, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _vp, _vp]),
    "dp_dnn_create": (_int, [_vp, _vp, _vp, _int, _int, _int, ctypes.POINTER(_vp)]),
    "dp_dnn_destroy": (None, [_vp]),
    "dp_dnn_mask_words": (_int, [_vp, ctypes.POINTER(_int)]),
    "dp_dnn_forward": (_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_dnn_backward": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
}

# diagnostics library (include/density_planner_b200_diag.h): fp32 cross-check rollout + micro-probes, not product ABI
_DIAG_SIGNATURES = {
    "dp_diag_last_error": (ctypes.c_char_p, []),
    "dp_diag_rollout_ffma": (_int, [_vp, _vp, _vp, _vp, _vp, _f32, _i64, _int, _uint, _int, _vp, _vp, _i64, _vp, _vp, _vp]),
    "dp_probe_ffma_peak": (_int, [_int, ctypes.POINTER(ctypes.c_double)]),
    "dp_tc_mma_bench": (_int, [_int, _int, _int, _int, ctypes.POINTER(ctypes.c_double)]),
    "dp_tc_tmem_bench": (_int, [_int, _int, _int, ctypes.POINTER(ctypes.c_double)]),
    "dp_tc_selftest": (_int, [_int, _vp, _vp, _int, _int, _int, _f32, _vp]),
}

EXPORTS = tuple(_SIGNATURES)
DIAG_EXPORTS = tuple(_DIAG_SIGNATURES)
_lib = None
_diag = None


def build(verbose=False):
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j", "8"], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DensityPlannerError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C density_planner_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.dp_abi_version() != 1:
            raise DensityPlannerError("ABI version mismatch")
        _lib = L
    return _lib


def diag_lib():
    """The diagnostics library (fp32 cross-check rollout, tcgen05 / TMEM / FFMA probes)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise DensityPlannerError("%s not found: build it with `make -C density_planner_b200/csrc`" % DIAG_LIB_PATH)
        L = ctypes.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in _DIAG_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _diag = L
    return _diag


def check_diag(rc, what=""):
    if rc != 0:
        msg = diag_lib().dp_diag_last_error()
        raise DensityPlannerError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def check(rc, what=""):
    if rc != 0:
        msg = lib().dp_last_error()
        raise DensityPlannerError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
