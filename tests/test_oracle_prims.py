"""Oracle primitives against cv2-python golden vectors and analytic known
answers.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_prims.npz"))
_dp = C.POINTER(C.c_double)


def test_gray_matches_cv2(orc_lib):
    rgb = np.ascontiguousarray(G["gray_rgb"])
    n = rgb.shape[0] * rgb.shape[1]
    out = np.zeros(n, np.float32)
    # the FMA variant is what cv2 4.13 computes: must be bit-exact
    orc_lib.gray_f32(rgb.ctypes.data_as(_dp), n, out.ctypes.data_as(C.POINTER(C.c_float)), 1)
    assert np.array_equal(out.reshape(rgb.shape[:2]), G["gray_cv2"])
    # the normative OpenCV 2.4.8 form (no FMA, three separately rounded products) stays within 2 ulp
    orc_lib.gray_f32(rgb.ctypes.data_as(_dp), n, out.ctypes.data_as(C.POINTER(C.c_float)), 0)
    ref = G["gray_cv2"].ravel()
    ulp = np.abs(out.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
    assert (ulp > 0).mean() < 0.5


def test_sobel_matches_cv2(orc_lib):
    gray = np.ascontiguousarray(G["gray_cv2"])
    h, w = gray.shape
    out = np.zeros((h, w))
    orc_lib.sobel_x(gray.ctypes.data_as(C.POINTER(C.c_float)), w, h, out.ctypes.data_as(_dp))
    assert np.array_equal(out, G["sobel_cv2"])
    assert np.all(out[:, 0] == 0) and np.all(out[:, -1] == 0)


@pytest.mark.parametrize("i", range(4))
def test_pyrdown_matches_cv2(orc_lib, i):
    src = np.ascontiguousarray(G[f"pyr{i}_src"])
    h, w, _ = src.shape
    dst = np.zeros(((h + 1) // 2, (w + 1) // 2, 3), np.uint8)
    orc_lib.pyrdown(src.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, dst.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(dst, G[f"pyr{i}_dst"])


def test_invert_matches_cv2(orc_lib):
    worst = 0.0
    for k in range(int(G["inv_count"])):
        A = np.ascontiguousarray(G[f"inv{k}_A"])
        n = A.shape[0]
        out = np.zeros((n, n))
        assert orc_lib.invert(A.ctypes.data_as(_dp), n, out.ctypes.data_as(_dp)) == 1
        ref = G[f"inv{k}_inv"]
        worst = max(worst, np.abs(out - ref).max() / np.abs(ref).max())
        assert np.allclose(out, ref, rtol=1e-14, atol=1e-16)  # SURVEY §8c tolerance for Mat::inv
    print("worst relative deviation from cv2.invert:", worst)


def test_scale_weight_known_answers(orc_lib):
    """pre_cs_pc.cc:89-109 known answers."""
    l = np.zeros((8, 8, 3), np.uint8)
    for S, lam, want in (
        (3, 0.3, [0.805668016194332, 0.157894736842105, 0.036437246963563]),
        (5, 0.3, [0.805399880759731, 0.156732816625500, 0.030508474576271, 0.005979047781279, 0.001379780257218]),
        (4, 0.0, [1.0, 0.0, 0.0, 0.0]),
    ):
        pc = orc.PlaneCost(orc_lib, l, l, 4, 35, S, lam)
        wgt, _, dims = pc.info()
        assert np.allclose(wgt, want, rtol=0, atol=1e-14)
        assert abs(wgt.sum() - 1.0) < 1e-14 or lam == 0.0
        assert [tuple(d) for d in dims][:3] == [(8, 8, 4), (4, 4, 2), (2, 2, 1)]


def test_round2int_and_border(orc_lib):
    # commfunc.h:117-121 round-half-even; :129-145 single wrap
    for v, want in ((2.5, 2), (3.5, 4), (-0.5, 0), (-1.5, -2), (1.4999999, 1), (254.5, 254), (255.5, 256)):
        assert orc_lib.round2int(v) == want
    assert orc_lib.handle_border(-1, 10) == 9
    assert orc_lib.handle_border(10, 10) == 0
    assert orc_lib.handle_border(13, 10) == 3
    assert orc_lib.handle_border(5, 10) == 5


def test_n_ref(orc_lib):
    l = np.zeros((4, 4, 3), np.uint8)
    for d, want in ((60, 9), (128, 10), (256, 11)):
        pm = orc.PatchMatch(orc_lib, l, l, d, 1)
        assert orc_lib.pm_n_ref(pm.h_) == want


def test_plane_param(orc_lib):
    # plane.h:25-34: a=-nx/nz, b=-ny/nz, c=n.p/nz; |nz|<1e-8 guarded with the sign of nz
    def prm(n, p):
        n, p, o = np.array(n, float), np.array(p, float), np.zeros(3)
        orc_lib.plane_param(n.ctypes.data_as(_dp), p.ctypes.data_as(_dp), o.ctypes.data_as(_dp))
        return o
    assert np.array_equal(prm([0, 0, 1], [3, 4, 5]), [0, 0, 5])
    assert np.array_equal(prm([0, 0, -1], [3, 4, 5]), [0, 0, 5])
    assert np.allclose(prm([0.6, 0, 0.8], [1, 2, 3]), [-0.75, 0, (0.6 + 2.4) / 0.8])
    assert np.array_equal(prm([1, 0, 0], [1, 0, 0]), [-1e8, 0, 1e8])
    assert np.array_equal(prm([1, 0, -1e-12], [1, 0, 0])[:2], [1e8, 0])


def test_rng_contract(orc_lib):
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter/key all zero and all ones
    assert orc_lib.rng_draw(0, 0, 0, 0) == 0x6627E8D5 | (0xE169C58D << 32)
    ones = 0xFFFFFFFF
    assert orc_lib.rng_draw(ones | (ones << 32), ones | (ones << 32), ones, ones) == 0x408F276D | (0x41C83B0E << 32)
    u = [orc_lib.rng_uniform(1, 2, s, 0, -3.0, 5.0) for s in range(2000)]
    assert min(u) >= -3.0 and max(u) < 5.0 and abs(np.mean(u) - 1.0) < 0.2
    v = np.zeros(3)
    vs = []
    for s in range(2000):
        orc_lib.rng_normal3(7, 9, s, v.ctypes.data_as(_dp))
        vs.append(v.copy())
    vs = np.array(vs)
    assert np.allclose(np.linalg.norm(vs, axis=1), 1.0, atol=1e-12)
    assert np.all(np.abs(vs.mean(0)) < 0.06)  # uniform on the sphere


def test_grd_cost_known_answers(orc_lib):
    # constant-colour pair: raw cost 0 wherever the partner is inside the image; max raw cost 2.8
    img = np.full((6, 10, 3), 77.0)
    vol = orc_lib.build_cv("GRD", img, img, 4, 0)
    for d in range(4):
        assert np.all(vol[d][:, d:] == 0.0)
        assert np.all(vol[d][:, :d] == 0.1 * 10 + (1 - 0.1) * 2) or d == 0
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (12, 40, 3)).astype(np.float64)
    b = rng.integers(0, 256, (12, 40, 3)).astype(np.float64)
    vol = orc_lib.build_cv("GRD", a, b, 8, 1)
    assert vol.max() == 2.8 and vol.min() >= 0.0
    assert orc_lib.cc_build_cv(b"BSM", None, None, 1, 1, 1, 0, None) == -1  # GetCCType -> NULL


def test_plane_cost_known_answers(orc_lib):
    from crossscalepatchmatch_b200.synth import make_pair

    w, h, d = 60, 50, 12
    l, r, _ = make_pair(w, h, d, 3)
    pc = orc.PlaneCost(orc_lib, l, r, d, 35, 0, 0.0)
    lrgb, rrgb = l[..., ::-1].astype(np.float64), r[..., ::-1].astype(np.float64)
    vol = orc_lib.build_cv("GRD", lrgb, rrgb, d + 1, 0)
    lut = np.exp(-np.arange(1000) / 10.0)

    def weights(x, y):
        ys, xs = slice(max(y - 17, 0), min(y + 18, h)), slice(max(x - 17, 0), min(x + 18, w))
        sad = np.abs(l[ys, xs].astype(int) - l[y, x].astype(int)).sum(-1)
        return lut[sad], ys, xs

    # fronto-parallel plane at integer disparity d0: cost = sum_q w(p,q) * C[d0](q) exactly (fw == 1)
    for (x, y, d0) in ((30, 25, 5), (0, 0, 3), (w - 1, h - 1, 11), (10, 40, 1)):
        wq, ys, xs = weights(x, y)
        want = 0.0
        for row_w, row_c in zip(wq, vol[d0][ys, xs]):
            for a, b in zip(row_w, row_c):
                want += a * (1.0 * b + (1 - 1.0) * 0.0) if False else a * b
        got = pc.get_plane_cost([x], [y], [[0, 0, 1.0]], [[x, y, float(d0)]], [0])[0]
        assert abs(got - want) <= 1e-12 * max(1.0, want)
    # plane whose disparity floor is <= 0 everywhere: cost = max_cost * sum_q w
    wq, _, _ = weights(20, 20)
    got = pc.get_plane_cost([20], [20], [[0, 0, 1.0]], [[20, 20, 0.5]], [0])[0]
    assert abs(got - 2.8 * wq.sum()) <= 1e-12 * got
    # window clipping: 18 x 18 taps at the corner
    assert weights(0, 0)[0].shape == (18, 18)
