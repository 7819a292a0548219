"""Pin the oracle restatement (oracle/oracle.cc) against the REFERENCE'S OWN
CODE: (a) the committed golden vectors generated from oracle/_ref, always;
(b) oracle/_ref/libcspm_ref.so live, when it is present.  Bit-exact throughout.
CPU only."""
import os

import numpy as np
import pytest

import orc

V = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
TAGS = (("ss", 0, 0.0, "GRD"), ("cs3", 3, 0.3, "GRD"), ("cs5", 5, 0.3, "GRD"),
        ("cen", 0, 0.0, "CEN"),                                # PreSSPC over CenCC
        ("grdpc", 0, 0.0, "IMG"), ("cspc3", 3, 0.3, "IMG"))   # GrdPC, CSPC


def test_cost_volume_vs_golden(orc_lib):
    l, r = V["l"], V["r"]
    d = int(V["whd"][2])
    lrgb, rrgb = l[..., ::-1].astype(np.float64), r[..., ::-1].astype(np.float64)
    assert np.array_equal(orc_lib.build_cv("GRD", lrgb, rrgb, d + 1, 0), V["cv_left"])
    assert np.array_equal(orc_lib.build_cv("GRD", lrgb, rrgb, d + 1, 1), V["cv_right"])
    assert np.array_equal(orc_lib.build_cv("CEN", lrgb, rrgb, d + 1, 0), V["cen_left"])
    assert np.array_equal(orc_lib.build_cv("CEN", lrgb, rrgb, d + 1, 1), V["cen_right"])


@pytest.mark.parametrize("tag,S,lam,cc", TAGS)
def test_plane_cost_and_patchmatch_vs_golden(orc_lib, tag, S, lam, cc):
    l, r = V["l"], V["r"]
    d = int(V["whd"][2])
    pc = orc.PlaneCost(orc_lib, l, r, d, 35, S, lam, cc_name=cc)
    wgt, mc, dims = pc.info()
    assert np.array_equal(wgt, V[f"{tag}_wgt"]) and np.array_equal(mc, V[f"{tag}_maxcost"]) and np.array_equal(dims, V[f"{tag}_dims"])
    cost = pc.get_plane_cost(V["q_x"], V["q_y"], V["q_n"], V["q_p"], V["q_v"])
    assert np.array_equal(cost, V[f"{tag}_cost"])
    pm = orc.PatchMatch(orc_lib, l, r, d, 4, seed=0x5EED)
    pm.patchmatch(2, pc, False, nthreads=2)  # thread count must not matter
    for view in (0, 1):
        n, p, c = pm.planes(view)
        assert np.array_equal(n, V[f"{tag}_normal{view}"])
        assert np.array_equal(p, V[f"{tag}_param{view}"])
        assert np.array_equal(c, V[f"{tag}_mincost{view}"])
        assert np.array_equal(pm.dis(view), V[f"{tag}_dis{view}"])
    pm.stage(orc.PatchMatch.POSTPROC)
    for view in (0, 1):
        assert np.array_equal(pm.dis(view), V[f"{tag}_dispp{view}"])


@pytest.mark.parametrize("w,h,d,S,lam,iters,cc", [(40, 30, 10, 0, 0.0, 3, "GRD"), (33, 27, 20, 3, 0.5, 2, "GRD"),
                                                  (70, 12, 30, 2, 0.3, 1, "GRD"), (36, 28, 9, 3, 0.3, 2, "CEN"),
                                                  (40, 30, 10, 0, 0.0, 2, "IMG"), (33, 27, 20, 3, 0.5, 2, "IMG")])
def test_live_against_reference_build(orc_lib, ref_lib, w, h, d, S, lam, iters, cc):
    from crossscalepatchmatch_b200.synth import make_pair

    ref_lib.set_time_mode(0)
    l, r, _ = make_pair(w, h, d, 100 + w)
    po, pr = orc.PlaneCost(orc_lib, l, r, d, 35, S, lam, cc_name=cc), orc.PlaneCost(ref_lib, l, r, d, 35, S, lam, cc_name=cc)
    for a, b in zip(po.info(), pr.info()):
        assert np.array_equal(a, b)
    mo, mr = orc.PatchMatch(orc_lib, l, r, d, 2, seed=11), orc.PatchMatch(ref_lib, l, r, d, 2, seed=11)
    mo.patchmatch(iters, po, True)
    mr.patchmatch(iters, pr, True)
    for view in (0, 1):
        for a, b in zip(mo.planes(view), mr.planes(view)):
            assert np.array_equal(a, b)
        assert np.array_equal(mo.dis(view), mr.dis(view))


def test_ref_parallel_rng_mode(ref_lib):
    """oracle/cvshim time() mode 2: the reference's OpenMP row loops run on several threads and every row still
    receives the RNG stream id of the single-threaded run (mode 0) — the full-size fixtures in
    tests/golden/cfg_*.npz are generated this way.  Also: the -O3 -march=x86-64-v3 timing build computes the
    same bits as the -O2 parity build."""
    from crossscalepatchmatch_b200.synth import make_pair

    w, h, d = 51, 37, 14
    l, r, _ = make_pair(w, h, d, 77)

    def run(lib, mode, nthreads, S):
        lib.set_time_mode(mode)
        pc = orc.PlaneCost(lib, l, r, d, 35, S, 0.3 if S else 0.0)
        pm = orc.PatchMatch(lib, l, r, d, 4, seed=0x5EED)
        pm.patchmatch(2, pc, False, nthreads)
        out = [pm.dis(0), pm.dis(1)] + list(pm.planes(0)) + list(pm.planes(1))
        lib.set_time_mode(0)
        return out

    libs = [ref_lib] + ([orc.ref_fast()] if orc.ref_fast_available() else [])
    for S in (0, 3):
        base = run(ref_lib, 0, 1, S)
        for lib in libs:
            for nt in (1, 3, 5):
                got = run(lib, 2, nt, S)
                assert all(np.array_equal(a, b) for a, b in zip(base, got)), (lib.path, nt, S)


@pytest.mark.parametrize("schedule", [1, 2])
def test_oracle_redblack_schedule_is_order_free(orc_lib, schedule):
    """The oracle's red-black schedules (restated for the kernel's throughput mode; not in the reference) do not
    depend on the processing order of the active pixels (thread count), and differ from the raster scan."""
    from crossscalepatchmatch_b200.synth import make_pair

    w, h, d = 40, 30, 10
    l, r, _ = make_pair(w, h, d, 5)
    pc = orc.PlaneCost(orc_lib, l, r, d, 35, 0, 0.0)
    outs = []
    for nt, sch in ((1, schedule), (4, schedule), (1, 0)):
        pm = orc.PatchMatch(orc_lib, l, r, d, 4, seed=3, schedule=sch)
        pm.patchmatch(2, pc, False, nt)
        outs.append(pm.planes(0))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert not np.array_equal(outs[0][1], outs[2][1])
    # costs stay GetPlaneCost of the stored planes
    n, p, c = outs[0]
    ys, xs = np.mgrid[0:h, 0:w]
    pt = np.stack([xs.ravel(), ys.ravel(), (p[..., 0] * xs + p[..., 1] * ys + p[..., 2]).ravel()], 1).astype(np.float64)
    got = pc.get_plane_cost(xs.ravel(), ys.ravel(), n.reshape(-1, 3), pt, np.zeros(w * h, np.int32))
    assert np.allclose(got, c.ravel(), rtol=1e-9)
