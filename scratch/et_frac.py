import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
w, h, d = 960, 540, 128
l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
pc = c.PreSSPC(l, r, d, 35, c.GetCCType("GRD"))
pm = c.CSPatchMatch(l, r, d, 1)
out = (C.c_uint64 * 8)()
pm.stage(c.STAGE_INIT, 0, pc)
for it in range(3):
    for st, name in ((c.STAGE_SPATIAL, "spatial"), (c.STAGE_VIEW, "view"), (c.STAGE_REFINE, "refine")):
        c._lib.dll.cspm_debug_counters(out, 1)
        pm.stage(st, it, pc)
        c._lib.dll.cspm_debug_counters(out, 1)
        if out[7]: print(it, name, "rows executed / nominal = %.3f" % (out[6] / out[7]))
