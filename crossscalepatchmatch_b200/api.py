"""Host-side mirror of the reference's plugin surface over the C ABI.

Same class names, argument order and meaning as the reference so parity tests
read like the reference's call sites (CSPM/main.cc:89-120):

    cc = GetCCType("GRD")
    pc = PreCSPC(l_img, r_img, max_dis, 35, 5, cc, reg_lambda)   # or PreSSPC(...)
    pm = CSPatchMatch(l_img, r_img, max_dis, dis_scale)
    pm.PatchMatch(3, pc, use_pp)
    l_dis = pm.dis(kLeft)

Images are numpy uint8 HxWx3 BGR (what cv::imread returns).  Everything
numerical happens in libcspm_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CspmError, check, dll

kLeft, kRight = 0, 1  # enum RefView, commfunc.h:29
_dp, _ip, _u8p, _fp = _lib._dp, _lib._ip, _lib._u8p, _lib._fp


def _img(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        # CV_Assert(l_img.type() == CV_8UC3), pre_ss_pc.cc:24 / cs_patchmatch.cc:8
        raise CspmError(_lib.EINVAL, "image must be uint8 HxWx3 (CV_8UC3)")
    if a.strides[2] != 1 or a.strides[1] != 3:
        a = np.ascontiguousarray(a)
    return a


def set_thread_stream(cuda_stream: int | None):
    """Handles created by this thread afterwards launch on `cuda_stream` (a raw cudaStream_t)."""
    dll.cspm_set_thread_stream(C.c_void_p(cuda_stream) if cuda_stream else None)


class Plane:
    """plane.h:12-48 — unit normal + point, derived (a, b, c)."""

    def __init__(self, norm=(0.0, 0.0, 0.0), point=(0.0, 0.0, 0.0)):
        self._norm = np.asarray(norm, np.float64).copy()
        self._point = np.asarray(point, np.float64).copy()

    def set_point(self, p):
        self._point = np.asarray(p, np.float64).copy()

    def set_norm(self, n):
        self._norm = np.asarray(n, np.float64).copy()

    def norm(self):
        return self._norm

    def point(self):
        return self._point


class CCMethod:
    """cc_method.h:15-33."""

    def __init__(self, name: str):
        self._h = dll.cspm_cc_create(name.encode())
        if not self._h:
            raise CspmError(_lib.EINVAL, _lib.last_error())
        self.name = name

    def __del__(self):
        if getattr(self, "_h", None) and dll is not None:  # (module globals are gone at interpreter shutdown)
            dll.cspm_cc_destroy(self._h)
            self._h = None

    def _build(self, lImg, rImg, maxDis, right, device):
        l = np.ascontiguousarray(lImg, np.float64)
        r = np.ascontiguousarray(rImg, np.float64)
        if l.ndim != 3 or l.shape[2] != 3 or l.shape != r.shape:
            raise CspmError(_lib.EINVAL, "images must be HxWx3 float64 (CV_64FC3)")  # CV_Assert grd_cc.cpp:63
        h, w, _ = l.shape
        vol = np.empty((maxDis, h, w), np.float64)
        check(dll.cspm_cc_build_cv(self._h, l.ctypes.data_as(_dp), r.ctypes.data_as(_dp), w, h, maxDis, right,
                                   vol.ctypes.data_as(_dp), device))
        return vol

    def buildCV(self, lImg, rImg, maxDis, device=0):
        return self._build(lImg, rImg, maxDis, 0, device)

    def buildRightCV(self, lImg, rImg, maxDis, device=0):
        return self._build(lImg, rImg, maxDis, 1, device)


class GrdCC(CCMethod):
    def __init__(self):
        super().__init__("GRD")


class CenCC(CCMethod):
    def __init__(self):
        super().__init__("CEN")


def GetCCType(name: str):
    """main.cc:39-55 — returns None for unknown names, like the reference."""
    if name == "GRD":
        return GrdCC()
    if name == "CEN":
        return CenCC()
    return None


class IPlaneCost:
    """plane_cost/i_plane_cost.h:13-34."""

    _h = None

    def __del__(self):
        if getattr(self, "_h", None) and dll is not None:  # (module globals are gone at interpreter shutdown)
            dll.cspm_pc_destroy(self._h)
            self._h = None

    def GetPlaneCost(self, ref_x, ref_y, plane: Plane, view) -> float:
        return float(self.GetPlaneCostBatch([ref_x], [ref_y], [plane.norm()], [plane.point()], [view])[0])

    def GetPlaneCostBatch(self, x, y, normal, point, view) -> np.ndarray:
        x = np.ascontiguousarray(x, np.int32)
        y = np.ascontiguousarray(y, np.int32)
        view = np.ascontiguousarray(view, np.int32)
        normal = np.ascontiguousarray(normal, np.float64).reshape(-1, 3)
        point = np.ascontiguousarray(point, np.float64).reshape(-1, 3)
        n = len(x)
        if not (len(y) == len(view) == len(normal) == len(point) == n):
            raise CspmError(_lib.EINVAL, "query arrays differ in length")
        out = np.zeros(n, np.float64)
        check(dll.cspm_pc_get_plane_cost(self._h, n, x.ctypes.data_as(_ip), y.ctypes.data_as(_ip),
                                         normal.ctypes.data_as(_dp), point.ctypes.data_as(_dp),
                                         view.ctypes.data_as(_ip), out.ctypes.data_as(_dp)))
        return out

    def info(self):
        S = check(dll.cspm_pc_info(self._h, None, None, None))
        wgt, mc, dims = np.zeros(S), np.zeros((2, S)), np.zeros((S, 3), np.int32)
        dll.cspm_pc_info(self._h, wgt.ctypes.data_as(_dp), mc.ctypes.data_as(_dp), dims.ctypes.data_as(_ip))
        return wgt, mc, dims

    def scale_image(self, s, view):
        _, _, dims = self.info()
        out = np.zeros((dims[s][1], dims[s][0], 3), np.uint8)
        check(dll.cspm_pc_scale_image(self._h, s, view, out.ctypes.data_as(_u8p)))
        return out

    def stream(self) -> int:
        return dll.cspm_pc_stream(self._h) or 0

    def taps_per_image_eval(self) -> int:
        return int(dll.cspm_pc_taps_per_image_eval(self._h))

    def stats(self):
        s = _lib.Stats()
        check(dll.cspm_pc_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in _lib.Stats._fields_}


def _ptr_and_geom(img, device_ptr):
    if device_ptr is not None:
        ptr, w, h, stride = device_ptr
        return C.c_void_p(ptr), w, h, stride, None
    a = _img(img)
    return C.c_void_p(a.ctypes.data), a.shape[1], a.shape[0], a.strides[0], a


class PreSSPC(IPlaneCost):
    """plane_cost/pre_ss_pc.h:18-52 — PreSSPC(l_img, r_img, max_disp, wnd_size, cc_method)."""

    def __init__(self, l_img, r_img, max_disp, wnd_size, cc_method: CCMethod, device=0, device_ptrs=None):
        if cc_method is None:
            raise CspmError(_lib.EINVAL, "cc_method is NULL")
        lp, w, h, stride, self._kl = _ptr_and_geom(l_img, device_ptrs[0] if device_ptrs else None)
        rp, w2, h2, stride2, self._kr = _ptr_and_geom(r_img, device_ptrs[1] if device_ptrs else None)
        if (w, h, stride) != (w2, h2, stride2):
            raise CspmError(_lib.EINVAL, "left/right geometry differs")
        fn = dll.cspm_pc_create_ss_dev if device_ptrs else dll.cspm_pc_create_ss
        self._h = fn(lp, rp, w, h, stride, max_disp, wnd_size, cc_method._h, device)
        if not self._h:
            _lib.raise_last()


class PreCSPC(IPlaneCost):
    """plane_cost/pre_cs_pc.h:19-62 — PreCSPC(l, r, max_disp, wnd_size, scale_num, cc_method, reg_lambda)."""

    def __init__(self, l_img, r_img, max_disp, wnd_size, scale_num, cc_method: CCMethod, reg_lambda, device=0,
                 device_ptrs=None):
        if cc_method is None:
            raise CspmError(_lib.EINVAL, "cc_method is NULL")
        lp, w, h, stride, self._kl = _ptr_and_geom(l_img, device_ptrs[0] if device_ptrs else None)
        rp, w2, h2, stride2, self._kr = _ptr_and_geom(r_img, device_ptrs[1] if device_ptrs else None)
        if (w, h, stride) != (w2, h2, stride2):
            raise CspmError(_lib.EINVAL, "left/right geometry differs")
        fn = dll.cspm_pc_create_cs_dev if device_ptrs else dll.cspm_pc_create_cs
        self._h = fn(lp, rp, w, h, stride, max_disp, wnd_size, scale_num, cc_method._h, float(reg_lambda), device)
        if not self._h:
            _lib.raise_last()


class GrdPC(IPlaneCost):
    """plane_cost/grd_pc.h:19-71 — GrdPC(l_img, r_img, max_disp, wnd_size): image-interpolating plane cost."""

    def __init__(self, l_img, r_img, max_disp, wnd_size, device=0):
        l, r = _img(l_img), _img(r_img)
        if l.shape != r.shape or l.strides != r.strides:
            raise CspmError(_lib.EINVAL, "left/right geometry differs")
        self._h = dll.cspm_pc_create_grdpc(C.c_void_p(l.ctypes.data), C.c_void_p(r.ctypes.data), l.shape[1], l.shape[0],
                                           l.strides[0], max_disp, wnd_size, device)
        if not self._h:
            _lib.raise_last()


class CSPC(IPlaneCost):
    """plane_cost/cspc.h:17-62 — CSPC(l_img, r_img, max_disp, wnd_size, scale_num, reg_lambda)."""

    def __init__(self, l_img, r_img, max_disp, wnd_size, scale_num, reg_lambda, device=0):
        l, r = _img(l_img), _img(r_img)
        if l.shape != r.shape or l.strides != r.strides:
            raise CspmError(_lib.EINVAL, "left/right geometry differs")
        self._h = dll.cspm_pc_create_cspc(C.c_void_p(l.ctypes.data), C.c_void_p(r.ctypes.data), l.shape[1], l.shape[0],
                                          l.strides[0], max_disp, wnd_size, scale_num, float(reg_lambda), device)
        if not self._h:
            _lib.raise_last()


class CSPatchMatch:
    """cs_patchmatch.h:16-147 — CSPatchMatch(l_img, r_img, max_dis, dis_scale)."""

    def __init__(self, l_img, r_img, max_dis, dis_scale, seed=0x5EED, schedule=_lib.SCHEDULE_RASTER, device=0,
                 device_ptrs=None):
        lp, w, h, stride, self._kl = _ptr_and_geom(l_img, device_ptrs[0] if device_ptrs else None)
        rp, w2, h2, stride2, self._kr = _ptr_and_geom(r_img, device_ptrs[1] if device_ptrs else None)
        if (w, h, stride) != (w2, h2, stride2):
            raise CspmError(_lib.EINVAL, "left/right geometry differs")
        self.h, self.w = h, w
        fn = dll.cspm_pm_create_dev if device_ptrs else dll.cspm_pm_create
        self._h = fn(lp, rp, w, h, stride, max_dis, dis_scale, seed, schedule, device)
        if not self._h:
            _lib.raise_last()

    def __del__(self):
        if getattr(self, "_h", None) and dll is not None:  # (module globals are gone at interpreter shutdown)
            dll.cspm_pm_destroy(self._h)
            self._h = None

    def PatchMatch(self, iter_num, plane_cost: IPlaneCost, use_pp=False):
        check(dll.cspm_pm_patchmatch(self._h, iter_num, plane_cost._h, int(bool(use_pp))))

    def dis(self, view) -> np.ndarray:
        out = np.zeros((self.h, self.w), np.uint8)
        check(dll.cspm_pm_get_dis(self._h, view, out.ctypes.data_as(_u8p), self.w))
        return out

    # ---- parity / measurement hooks (not in the reference's public surface) --
    def stage(self, stage, it=0, plane_cost: IPlaneCost | None = None):
        check(dll.cspm_pm_stage(self._h, stage, it, plane_cost._h if plane_cost is not None else None))

    def set_dis(self, view, d):
        d = np.ascontiguousarray(d, np.uint8)
        check(dll.cspm_pm_set_dis(self._h, view, d.ctypes.data_as(_u8p), self.w))

    def planes(self, view):
        n, p, c = np.zeros((self.h, self.w, 3)), np.zeros((self.h, self.w, 3)), np.zeros((self.h, self.w))
        check(dll.cspm_pm_get_planes(self._h, view, n.ctypes.data_as(_dp), p.ctypes.data_as(_dp), c.ctypes.data_as(_dp)))
        return n, p, c

    def set_planes(self, view, n, p, c):
        n = np.ascontiguousarray(n, np.float64)
        p = np.ascontiguousarray(p, np.float64)
        c = np.ascontiguousarray(c, np.float64)
        check(dll.cspm_pm_set_planes(self._h, view, n.ctypes.data_as(_dp), p.ctypes.data_as(_dp), c.ctypes.data_as(_dp)))

    def n_ref(self) -> int:
        return check(dll.cspm_pm_n_ref(self._h))

    def stream(self) -> int:
        return dll.cspm_pm_stream(self._h) or 0

    def stats(self):
        s = _lib.Stats()
        check(dll.cspm_pm_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in _lib.Stats._fields_}

    def stage_ms(self):
        ms = (C.c_float * 6)()
        check(dll.cspm_pm_stage_ms(self._h, ms))
        return list(ms)


def PatchMatchBatch(patch_matches, plane_costs, iter_num, use_pp=False):
    """cspm_pm_patchmatch_batch: CSPatchMatch.PatchMatch for several independent pairs of identical geometry at
    once (throughput mode: the pairs' spatial sweeps share one launch).  Same results as calling
    pm.PatchMatch(iter_num, pc, use_pp) pair by pair."""
    n = len(patch_matches)
    if n != len(plane_costs) or n == 0:
        raise CspmError(_lib.EINVAL, "need as many plane costs as patch-match handles, at least one")
    pms = (C.c_void_p * n)(*[pm._h for pm in patch_matches])
    pcs = (C.c_void_p * n)(*[pc._h for pc in plane_costs])
    check(dll.cspm_pm_patchmatch_batch(pms, pcs, n, iter_num, int(bool(use_pp))))
