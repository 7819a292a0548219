"""Small PatchMatch runs for compute-sanitizer (memcheck / racecheck): every sweep implementation, single- and
cross-scale, a batch, post-processing; results checked against the oracle so that the run is also a parity test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import orc
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import make_pair

w, h, d = 40, 30, 10
O = orc.oracle()
cc = c.GetCCType("GRD")
for S, lam in ((0, 0.0), (2, 0.3)):
    pairs = [make_pair(w, h, d, 5 + k)[:2] for k in range(2)]
    pcs = [c.PreSSPC(l, r, d, 35, cc) if S == 0 else c.PreCSPC(l, r, d, 35, S, cc, lam) for l, r in pairs]
    pms = [c.CSPatchMatch(l, r, d, 4, seed=9 + k) for k, (l, r) in enumerate(pairs)]
    c.PatchMatchBatch(pms, pcs, 1, True)
    for k, (l, r) in enumerate(pairs):
        po = orc.PlaneCost(O, l, r, d, 35, S, lam)
        mo = orc.PatchMatch(O, l, r, d, 4, seed=9 + k)
        mo.patchmatch(1, po, True)
        for v in (0, 1):
            assert np.array_equal(mo.dis(v), pms[k].dis(v))
            for a, b in zip(mo.planes(v), pms[k].planes(v)):
                assert np.array_equal(a, b)
    rb = c.CSPatchMatch(pairs[0][0], pairs[0][1], d, 4, seed=3, schedule=1)
    rb.PatchMatch(1, pcs[0], False)
print("sanitize_small OK", os.environ.get("CSPM_SPATIAL"), os.environ.get("CSPM_SP_CTAS"), os.environ.get("CSPM_WAVE_EXACT"))
