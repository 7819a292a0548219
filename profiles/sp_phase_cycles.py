"""Diagnostic: where do the cycles of the persistent spatial sweep go (CTA 0)?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
w, h, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x256").split("x"))
l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
pc = c.PreSSPC(l, r, d, 35, c.GetCCType("GRD"))
pm = c.CSPatchMatch(l, r, d, 1)
pm.stage(c.STAGE_INIT, 0, pc)
out = (C.c_uint64 * 8)()
c._lib.dll.cspm_debug_counters(out, 1)
import time
t = time.time(); pm.stage(c.STAGE_SPATIAL, 0, pc); dt = time.time() - t
c._lib.dll.cspm_debug_counters(out, 1)
tot = sum(out[:4])
nd = w + h - 2
print(f"sweep {dt*1e3:.1f} ms, {nd} diagonals, {dt*1e6/nd:.1f} us/diag; CTA0 cycles/diag:",
      {k: round(v / nd) for k, v in zip(["setup", "eval", "accept", "barrier", "midCTA_consumer_busy", "midCTA_producer0_busy", "midCTA_chunkloop"], out)}, "total", round(tot / nd))
