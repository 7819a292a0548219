"""ctypes bindings for the test oracles (TEST INFRASTRUCTURE):
  oracle/liboracle.so         CPU restatement   (prefix orc_)
  oracle/_ref/libcspm_ref.so  the reference's own sources + cvshim (prefix ref_)
Both expose the same entry points, so one wrapper class serves both.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class Lib:
    def __init__(self, path: str, prefix: str):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        f = self._f
        f("cc_build_cv", C.c_int, [C.c_char_p, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp])
        f("pc_create", C.c_void_p, [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_char_p])
        f("pc_destroy", None, [C.c_void_p])
        f("pc_get_plane_cost", None, [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _ip, _dp, C.c_int])
        f("pc_info", C.c_int, [C.c_void_p, _dp, _dp, _ip])
        f("pc_scale_image", C.c_int, [C.c_void_p, C.c_int, C.c_int, _u8p])
        f("pm_create", C.c_void_p, [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64])
        f("pm_destroy", None, [C.c_void_p])
        f("pm_patchmatch", None, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int])
        f("pm_stage", None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int])
        f("pm_get_dis", None, [C.c_void_p, C.c_int, _u8p])
        f("pm_set_dis", None, [C.c_void_p, C.c_int, _u8p])
        f("pm_get_planes", None, [C.c_void_p, C.c_int, _dp, _dp, _dp])
        f("pm_set_planes", None, [C.c_void_p, C.c_int, _dp, _dp, _dp])
        if prefix == "orc_":
            f("gray_f32", None, [_dp, C.c_int, C.POINTER(C.c_float), C.c_int])
            f("sobel_x", None, [C.POINTER(C.c_float), C.c_int, C.c_int, _dp])
            f("pyrdown", None, [_u8p, C.c_int, C.c_int, _u8p])
            f("invert", C.c_int, [_dp, C.c_int, _dp])
            f("round2int", C.c_int, [C.c_double])
            f("handle_border", C.c_int, [C.c_int, C.c_int])
            f("rng_draw", C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32])
            f("rng_uniform", C.c_double, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_double])
            f("rng_normal3", None, [C.c_uint64, C.c_uint64, C.c_uint32, _dp])
            f("plane_param", None, [_dp, _dp, _dp])
            f("pm_n_ref", C.c_int, [C.c_void_p])
            f("pm_set_schedule", None, [C.c_void_p, C.c_int])
        else:
            f("set_seed", None, [C.c_uint64])
            f("set_time_mode", None, [C.c_int])

    def _f(self, name, res, args):
        fn = getattr(self.dll, self.prefix + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, name, fn)

    # -- helpers ------------------------------------------------------------
    def build_cv(self, name, l_rgb, r_rgb, max_dis, right_view):
        h, w, _ = l_rgb.shape
        l = np.ascontiguousarray(l_rgb, np.float64)
        r = np.ascontiguousarray(r_rgb, np.float64)
        vol = np.zeros((max_dis, h, w), np.float64)
        rc = self.cc_build_cv(name.encode(), _ptr(l, _dp), _ptr(r, _dp), w, h, max_dis, int(right_view), _ptr(vol, _dp))
        if rc != 0:
            raise ValueError(f"unknown cc_name {name}")
        return vol


class PlaneCost:
    """PreSSPC (scale_num=0) / PreCSPC of either oracle library."""

    def __init__(self, lib: Lib, l, r, max_disp, wnd_size=35, scale_num=0, reg_lambda=0.0, cc_name="GRD"):
        self.lib = lib
        self.l = np.ascontiguousarray(l, np.uint8)
        self.r = np.ascontiguousarray(r, np.uint8)
        h, w, _ = self.l.shape
        self.w, self.h, self.scale_num = w, h, scale_num
        self.h_ = lib.pc_create(_ptr(self.l, _u8p), _ptr(self.r, _u8p), w, h, self.l.strides[0], max_disp, wnd_size,
                                scale_num, reg_lambda, cc_name.encode())
        if not self.h_:
            raise ValueError("pc_create failed")

    def close(self):
        if self.h_:
            self.lib.pc_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        self.close()

    def get_plane_cost(self, x, y, normal, point, view, nthreads=0):
        x = np.ascontiguousarray(x, np.int32)
        y = np.ascontiguousarray(y, np.int32)
        view = np.ascontiguousarray(view, np.int32)
        normal = np.ascontiguousarray(normal, np.float64)
        point = np.ascontiguousarray(point, np.float64)
        out = np.zeros(len(x), np.float64)
        self.lib.pc_get_plane_cost(self.h_, len(x), _ptr(x, _ip), _ptr(y, _ip), _ptr(normal, _dp), _ptr(point, _dp),
                                   _ptr(view, _ip), _ptr(out, _dp), nthreads)
        return out

    def info(self):
        S = max(self.scale_num, 1)
        wgt = np.zeros(S)
        mc = np.zeros((2, S))
        dims = np.zeros((S, 3), np.int32)
        n = self.lib.pc_info(self.h_, _ptr(wgt, _dp), _ptr(mc, _dp), _ptr(dims, _ip))
        assert n == S
        return wgt, mc, dims

    def scale_image(self, s, view):
        _, _, dims = self.info()
        out = np.zeros((dims[s][1], dims[s][0], 3), np.uint8)
        self.lib.pc_scale_image(self.h_, s, view, _ptr(out, _u8p))
        return out


class PatchMatch:
    INIT, SPATIAL, VIEW, REFINE, PLANE2DISP, POSTPROC = range(6)

    def __init__(self, lib: Lib, l, r, max_dis, dis_scale, seed=0x5EED, schedule=0):
        self.lib = lib
        self.l = np.ascontiguousarray(l, np.uint8)
        self.r = np.ascontiguousarray(r, np.uint8)
        self.h, self.w, _ = self.l.shape
        self.h_ = lib.pm_create(_ptr(self.l, _u8p), _ptr(self.r, _u8p), self.w, self.h, self.l.strides[0], max_dis,
                                dis_scale, seed)
        if schedule:
            lib.pm_set_schedule(self.h_, schedule)  # oracle only: the reference has no such schedule

    def close(self):
        if self.h_:
            self.lib.pm_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        self.close()

    def patchmatch(self, iter_num, pc: PlaneCost, use_pp=False, nthreads=0):
        self.lib.pm_patchmatch(self.h_, iter_num, pc.h_, int(use_pp), nthreads)

    def stage(self, stage, it=0, pc: PlaneCost | None = None, nthreads=0):
        self.lib.pm_stage(self.h_, stage, it, pc.h_ if pc is not None else None, nthreads)

    def dis(self, view):
        out = np.zeros((self.h, self.w), np.uint8)
        self.lib.pm_get_dis(self.h_, view, _ptr(out, _u8p))
        return out

    def set_dis(self, view, d):
        d = np.ascontiguousarray(d, np.uint8)
        self.lib.pm_set_dis(self.h_, view, _ptr(d, _u8p))

    def planes(self, view):
        n = np.zeros((self.h, self.w, 3))
        p = np.zeros((self.h, self.w, 3))
        c = np.zeros((self.h, self.w))
        self.lib.pm_get_planes(self.h_, view, _ptr(n, _dp), _ptr(p, _dp), _ptr(c, _dp))
        return n, p, c

    def set_planes(self, view, n, p, c):
        n = np.ascontiguousarray(n, np.float64)
        p = np.ascontiguousarray(p, np.float64)
        c = np.ascontiguousarray(c, np.float64)
        self.lib.pm_set_planes(self.h_, view, _ptr(n, _dp), _ptr(p, _dp), _ptr(c, _dp))


_libs: dict[str, Lib] = {}


def oracle() -> Lib:
    if "orc" not in _libs:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _libs["orc"] = Lib(path, "orc_")
    return _libs["orc"]


def ref_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libcspm_ref.so"))


def ref() -> Lib:
    if "ref" not in _libs:
        _libs["ref"] = Lib(os.path.join(ORACLE_DIR, "_ref", "libcspm_ref.so"), "ref_")
    return _libs["ref"]


def ref_fast_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libcspm_ref_fast.so"))


def ref_fast() -> Lib:
    """The same unmodified reference sources built -O3 -march=x86-64-v3 (timing build, oracle/Makefile)."""
    if "ref_fast" not in _libs:
        _libs["ref_fast"] = Lib(os.path.join(ORACLE_DIR, "_ref", "libcspm_ref_fast.so"), "ref_")
    return _libs["ref_fast"]
