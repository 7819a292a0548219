"""Generate tests/golden/cv2_prims.npz from cv2-python (OpenCV 4.13) in the
authoring container.  These vectors pin oracle/prims.h — the restatement of the
OpenCV calls on the reference hot path (grd_cc.cpp:70-77, pre_cs_pc.cc:45,105).
Run: python tests/golden/make_cv2_prims.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20141001)
out = {}
# float gray + Sobel(ksize=1) on an RGB 0..255 f64 image, as GrdCC::buildCV does
rgb = rng.integers(0, 256, size=(37, 53, 3)).astype(np.float64)
gray = cv2.cvtColor(rgb.astype(np.float32), cv2.COLOR_RGB2GRAY)
out["gray_rgb"] = rgb
out["gray_cv2"] = gray
out["sobel_cv2"] = cv2.Sobel(gray, cv2.CV_64F, 1, 0, ksize=1)
# pyrDown on 8-bit BGR, odd and even sizes, three levels
for i, (h, w) in enumerate([(41, 67), (30, 44), (9, 5), (1, 7)]):
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    out[f"pyr{i}_src"] = img
    out[f"pyr{i}_dst"] = cv2.pyrDown(img)
# Mat::inv() of the cross-scale regulariser (pre_cs_pc.cc:89-105)
k = 0
for S in (2, 3, 4, 5, 6, 8):
    for lam in (0.0, 0.3, 1.0, 2.5):
        A = np.zeros((S, S))
        for s in range(S):
            if s == 0:
                A[s, s] = 1 + lam; A[s, s + 1] = -lam
            elif s == S - 1:
                A[s, s] = 1 + lam; A[s, s - 1] = -lam
            else:
                A[s, s] = 1 + 2 * lam; A[s, s - 1] = -lam; A[s, s + 1] = -lam
        ok, inv = cv2.invert(A, flags=cv2.DECOMP_LU)
        out[f"inv{k}_A"] = A
        out[f"inv{k}_inv"] = inv
        k += 1
out["inv_count"] = np.array(k)
# 8-bit RGB2GRAY (CenCC path); OpenCV 4.x uses 15-bit coefficients, 2.4.8 used 14-bit ones
img = rng.integers(0, 256, size=(16, 16, 3), dtype=np.uint8)
out["gray8_src"] = img
out["gray8_cv2"] = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "cv2_prims.npz"), **out)
print("wrote cv2_prims.npz with", len(out), "arrays; cv2", cv2.__version__)
