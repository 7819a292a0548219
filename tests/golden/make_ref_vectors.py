"""Generate tests/golden/ref_vectors.npz: outputs of the REFERENCE'S OWN CODE
(oracle/_ref/libcspm_ref.so = unmodified /root/reference sources + oracle/cvshim)
on small seeded inputs.  The inputs are stored too, so the fixture is
self-contained on machines without /root/reference.
Run (authoring container only): python tests/golden/make_ref_vectors.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc
from crossscalepatchmatch_b200.synth import make_pair

R = orc.ref()
R.set_time_mode(0)
out = {}
w, h, d = 48, 36, 12
l, r, _ = make_pair(w, h, d, 0xC5B20000)
out["l"], out["r"], out["whd"] = l, r, np.array([w, h, d])
lrgb, rrgb = l[..., ::-1].astype(np.float64), r[..., ::-1].astype(np.float64)
out["cv_left"] = R.build_cv("GRD", lrgb, rrgb, d + 1, 0)
out["cv_right"] = R.build_cv("GRD", lrgb, rrgb, d + 1, 1)
out["cen_left"] = R.build_cv("CEN", lrgb, rrgb, d + 1, 0)
out["cen_right"] = R.build_cv("CEN", lrgb, rrgb, d + 1, 1)
rng = np.random.default_rng(5)
n = 600
x, y, v = rng.integers(0, w, n), rng.integers(0, h, n), rng.integers(0, 2, n)
nm = rng.normal(size=(n, 3)); nm /= np.linalg.norm(nm, axis=1)[:, None]
nm[:20] = [0, 0, 1]
pt = np.stack([x, y, rng.uniform(-1, d + 1, n)], 1).astype(np.float64)
pt[:20, 2] = rng.integers(1, d, 20)
out["q_x"], out["q_y"], out["q_v"], out["q_n"], out["q_p"] = x, y, v, nm, pt
for tag, S, lam, cc in (("ss", 0, 0.0, "GRD"), ("cs3", 3, 0.3, "GRD"), ("cs5", 5, 0.3, "GRD"),
                        ("cen", 0, 0.0, "CEN"), ("grdpc", 0, 0.0, "IMG"), ("cspc3", 3, 0.3, "IMG")):
    pc = orc.PlaneCost(R, l, r, d, 35, S, lam, cc_name=cc)
    wgt, mc, dims = pc.info()
    out[f"{tag}_wgt"], out[f"{tag}_maxcost"], out[f"{tag}_dims"] = wgt, mc, dims
    out[f"{tag}_cost"] = pc.get_plane_cost(x, y, nm, pt, v, 1)
    pm = orc.PatchMatch(R, l, r, d, 4, seed=0x5EED)
    pm.patchmatch(2, pc, False)
    for view in (0, 1):
        nn, pp, cc = pm.planes(view)
        out[f"{tag}_normal{view}"], out[f"{tag}_param{view}"], out[f"{tag}_mincost{view}"] = nn, pp, cc
        out[f"{tag}_dis{view}"] = pm.dis(view)
    pm.stage(orc.PatchMatch.POSTPROC)
    for view in (0, 1):
        out[f"{tag}_dispp{view}"] = pm.dis(view)
np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
print("wrote ref_vectors.npz:", len(out), "arrays")
