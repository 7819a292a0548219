import sys, time
sys.path.insert(0, "/root/repo")
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
w, h, d = 1920, 1080, 256
l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
pc = c.PreSSPC(l, r, d, 35, c.GetCCType("GRD"))
pm = c.CSPatchMatch(l, r, d, 1)
pm.PatchMatch(3, pc, True)
print("stage_ms", pm.stage_ms())
