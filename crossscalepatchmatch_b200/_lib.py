"""ctypes binding of libcspm_b200.so (the C ABI declared in include/cspm_c.h).

The library is CUDA-only.  If the shared object is missing this module raises
at import; if there is no GPU every compute call raises CspmError(ENODEV).
There is deliberately no CPU fallback here: the CPU restatement lives in
oracle/ and is test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSPM_LIB_PATH: load an alternative build of the same ABI (kernel-variant experiments)
LIB_PATH = os.environ.get("CSPM_LIB_PATH") or os.path.join(_HERE, "libcspm_b200.so")

OK, EINVAL, ENODEV, ENOMEM, EUNSUPPORTED, ECUDA = 0, -1, -2, -3, -4, -5
SCHEDULE_RASTER, SCHEDULE_REDBLACK, SCHEDULE_REDBLACK8 = 0, 1, 2
STAGE_INIT, STAGE_SPATIAL, STAGE_VIEW, STAGE_REFINE, STAGE_PLANE2DISP, STAGE_POSTPROC = range(6)


class CspmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cspm error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("plane_evals", C.c_uint64), ("window_taps", C.c_uint64),
                ("spatial_tasks", C.c_uint64), ("spatial_candidates", C.c_uint64), ("spatial_adopted", C.c_uint64),
                ("spatial_exact_tasks", C.c_uint64)]


_dp, _ip, _u8p, _fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.POINTER(C.c_float)

# name -> (restype, argtypes); exactly the symbols include/cspm_c.h declares
SIGNATURES = {
    "cspm_last_error": (C.c_char_p, []),
    "cspm_last_error_code": (C.c_int, []),
    "cspm_debug_taps_executed": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "cspm_device_count": (C.c_int, []),
    "cspm_version": (C.c_char_p, []),
    "cspm_set_thread_stream": (None, [C.c_void_p]),
    "cspm_pc_taps_per_image_eval": (C.c_uint64, [C.c_void_p]),
    "cspm_pm_create_dev": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]),
    "cspm_cc_create": (C.c_void_p, [C.c_char_p]),
    "cspm_cc_destroy": (None, [C.c_void_p]),
    "cspm_cc_name": (C.c_char_p, [C.c_void_p]),
    "cspm_cc_build_cv": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "cspm_pc_create_ss": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "cspm_pc_create_cs": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_int]),
    "cspm_pc_create_ss_dev": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "cspm_pc_create_cs_dev": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_int]),
    "cspm_pc_create_grdpc": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cspm_pc_create_cspc": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
    "cspm_pc_destroy": (None, [C.c_void_p]),
    "cspm_pc_get_plane_cost": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _ip, _dp]),
    "cspm_pc_info": (C.c_int, [C.c_void_p, _dp, _dp, _ip]),
    "cspm_pc_scale_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _u8p]),
    "cspm_pm_create": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]),
    "cspm_pm_destroy": (None, [C.c_void_p]),
    "cspm_pm_patchmatch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "cspm_pm_patchmatch_batch": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "cspm_pm_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cspm_pm_get_dis": (C.c_int, [C.c_void_p, C.c_int, _u8p, C.c_int]),
    "cspm_pm_set_dis": (C.c_int, [C.c_void_p, C.c_int, _u8p, C.c_int]),
    "cspm_pm_get_planes": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "cspm_pm_set_planes": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "cspm_pm_n_ref": (C.c_int, [C.c_void_p]),
    "cspm_pc_stream": (C.c_void_p, [C.c_void_p]),
    "cspm_pm_stream": (C.c_void_p, [C.c_void_p]),
    "cspm_pm_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "cspm_pc_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "cspm_pm_stage_ms": (C.c_int, [C.c_void_p, _fp]),
    "cspm_debug_counters": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "cspm_debug_wave_samples": (C.c_int, [_dp, C.POINTER(C.c_uint32)]),
    "cspm_debug_lane_hist": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  crossscalepatchmatch_b200 has no CPU fallback."
    )

dll = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(dll, _name)
    _fn.restype, _fn.argtypes = _res, _args


def last_error() -> str:
    return (dll.cspm_last_error() or b"").decode()


def raise_last():
    """A creator returned NULL: raise with the code the library recorded for this thread."""
    raise CspmError(dll.cspm_last_error_code() or EINVAL, last_error())


def check(rc: int) -> int:
    if rc < 0:
        raise CspmError(rc, last_error())
    return rc
