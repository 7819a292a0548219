"""GPU against the REFERENCE'S OWN CODE in one step: libcspm_b200.so (through the C ABI) must reproduce, bit for bit,
the golden vectors that tests/golden/make_ref_vectors.py generated from oracle/_ref (the unmodified reference sources
compiled against the OpenCV shim) — plane-cost info, batched GetPlaneCost, the PatchMatch state after two iterations,
the disparity maps and the post-processed maps, for every plane-cost plugin.  (test_gpu_parity.py closes the same
loop through the oracle; this one needs neither the oracle nor /root/reference at run time.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

V = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
TAGS = (("ss", 0, 0.0, "GRD"), ("cs3", 3, 0.3, "GRD"), ("cs5", 5, 0.3, "GRD"), ("cen", 0, 0.0, "CEN"),
        ("grdpc", 0, 0.0, "IMG"), ("cspc3", 3, 0.3, "IMG"))


@pytest.mark.parametrize("tag,S,lam,cc", TAGS)
def test_gpu_reproduces_reference_golden_vectors(tag, S, lam, cc):
    import crossscalepatchmatch_b200 as c

    assert c.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    l, r = V["l"], V["r"]
    d = int(V["whd"][2])
    if cc == "IMG":
        pc = c.GrdPC(l, r, d, 35) if S == 0 else c.CSPC(l, r, d, 35, S, lam)
    else:
        ccm = c.GetCCType(cc)
        pc = c.PreSSPC(l, r, d, 35, ccm) if S == 0 else c.PreCSPC(l, r, d, 35, S, ccm, lam)
    wgt, mc, dims = pc.info()
    assert np.array_equal(wgt, V[f"{tag}_wgt"]) and np.array_equal(mc, V[f"{tag}_maxcost"]) and np.array_equal(dims, V[f"{tag}_dims"])
    cost = pc.GetPlaneCostBatch(V["q_x"], V["q_y"], V["q_n"], V["q_p"], V["q_v"])
    assert np.array_equal(cost, V[f"{tag}_cost"])
    pm = c.CSPatchMatch(l, r, d, 4, seed=0x5EED)
    pm.PatchMatch(2, pc, False)
    for view in (0, 1):
        n, p, cst = pm.planes(view)
        assert np.array_equal(n, V[f"{tag}_normal{view}"])
        assert np.array_equal(p, V[f"{tag}_param{view}"])
        assert np.array_equal(cst, V[f"{tag}_mincost{view}"])
        assert np.array_equal(pm.dis(view), V[f"{tag}_dis{view}"])
    pm.stage(c.STAGE_POSTPROC, 0, pc)
    for view in (0, 1):
        assert np.array_equal(pm.dis(view), V[f"{tag}_dispp{view}"])
