"""Synthetic rectified stereo pairs: band-limited texture,
piecewise-planar ground-truth disparity, right view rendered by warping.

The generator itself is the dependency-free C function of include/cspm_synth.h
(splitmix64); this module binds its host-only build libcspm_synth.so, so the
Python harness, the tests and the C++ driver (apps/cspm_main --synth=...) see
byte-identical images.  Importing this module does NOT load libcspm_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SEED_BASE = 0xC5B20000
_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB_PATH = os.path.join(_HERE, "libcspm_synth.so")
_dll = None


def _lib():
    global _dll
    if _dll is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: build it with `make -C crossscalepatchmatch_b200/csrc`")
        _dll = C.CDLL(SYNTH_LIB_PATH)
        _dll.cspm_synth_make_pair.restype = C.c_int
        _dll.cspm_synth_make_pair.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _dll.cspm_synth_make_pair_variant.restype = C.c_int
        _dll.cspm_synth_make_pair_variant.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    return _dll


def make_pair(w: int, h: int, max_dis: int, seed: int = SEED_BASE, variant: int = 0):
    """Return (left_bgr, right_bgr, gt_left_disparity): uint8 HxWx3 (BGR), f64 HxW.
    variant > 0: the same scene (ground truth) as variant 0 under a different, statistically equivalent texture."""
    left = np.empty((h, w, 3), np.uint8)
    right = np.empty((h, w, 3), np.uint8)
    gt = np.empty((h, w), np.float64)
    rc = _lib().cspm_synth_make_pair_variant(w, h, max_dis, seed, variant, left.ctypes.data, right.ctypes.data, gt.ctypes.data)
    if rc != 0:
        raise ValueError(f"cspm_synth_pair({w}, {h}, {max_dis}) failed")
    return left, right, gt
