"""Diagnostic build only (-DCSPM_PROFILE_LANES=1): active lanes per executed window row, per stage of the HD run."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
w, h, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x256").split("x"))
_cache = f"/tmp/pair_{w}x{h}x{d}.npz"
if os.path.exists(_cache):
    _z = np.load(_cache); l, r = _z["l"], _z["r"]
else:
    l, r, _ = make_pair(w, h, d, SEED_BASE + 3); np.savez(_cache, l=l, r=r)
pc = c.PreSSPC(l, r, d, 35, c.GetCCType("GRD"))
pm = c.CSPatchMatch(l, r, d, 1)
out = (C.c_uint64 * 33)()
dll = c._lib.dll
def report(name):
    dll.cspm_debug_lane_hist(out, 1)
    h_ = np.array(list(out), dtype=np.float64)
    tot = h_.sum()
    if tot == 0: return
    cum = np.cumsum(h_) / tot
    lanes = (h_ * np.arange(33)).sum() / tot
    print(f"{name}: warp-rows {tot:.3g}, mean active lanes {lanes:.2f}; share of warp-rows with <=4 / <=8 / <=16 / <=24 / 32 lanes: "
          f"{cum[4]:.3f} {cum[8]:.3f} {cum[16]:.3f} {cum[24]:.3f} {h_[32]/tot:.3f}")
dll.cspm_debug_lane_hist(out, 1)
pm.stage(c.STAGE_INIT, 0, pc); report("init")
for it in range(3):
    pm.stage(c.STAGE_SPATIAL, it, pc); dll.cspm_debug_lane_hist(out, 1)
    pm.stage(c.STAGE_VIEW, it, pc); report(f"iter {it} view")
    pm.stage(c.STAGE_REFINE, it, pc); report(f"iter {it} refine")
