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
