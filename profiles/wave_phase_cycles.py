"""Diagnostic: where do the cycles of the wavefront spatial sweep go (one CTA, grid / 2)?
Needs a -DCSPM_WAVE_PROFILE=1 build (CSPM_LIB_PATH).  usage: wave_phase_cycles.py [WxHxD] [cs]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
w, h, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x256").split("x"))
cs = len(sys.argv) > 2 and sys.argv[2] == "cs"
l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
cc = c.GetCCType("GRD")
pc = c.PreCSPC(l, r, d, 35, 3, cc, 0.3) if cs else c.PreSSPC(l, r, d, 35, cc)
pm = c.CSPatchMatch(l, r, d, 1)
pm.stage(c.STAGE_INIT, 0, pc)
out = (C.c_uint64 * 8)()
for it in range(2):
    c._lib.dll.cspm_debug_counters(out, 1)
    t = time.time(); pm.stage(c.STAGE_SPATIAL, it, pc); dt = time.time() - t
    c._lib.dll.cspm_debug_counters(out, 1)
    nd = max(1, out[7])
    names = ["nbr_wait", "bar1", "eval_w0", "wait_slowest_warp", "decide", "bar3", "publish"]
    print(f"{w}x{h}x{d} {'cs' if cs else 'ss'} sweep {it}: {dt*1e3:.1f} ms wall (incl. fix-up), {nd} diagonals, {dt*1e6/nd:.1f} us/diag; mid-CTA cycles/diag:",
          {k: round(v / nd) for k, v in zip(names, out)}, "total", round(sum(out[:7]) / nd), pm.stats(), flush=True)
    smp = (C.c_double * 256)(); ns = C.c_uint32(0)
    c._lib.dll.cspm_debug_wave_samples(smp, C.byref(ns))
    print(f"  undecided tasks met by the profiled CTA: {ns.value}; first ones (T0, T1, m, flags, (T0-m)/m, (T1-m)/m, (T1-T0)/T0):")
    for k in range(min(ns.value, 24)):
        T0, T1, m, fl = smp[4 * k:4 * k + 4]
        rel = lambda a, b: (a - b) / b if b else float("nan")
        print(f"    {T0!r:>24} {T1!r:>24} {m!r:>24} {int(fl)} {rel(T0, m):.3e} {rel(T1, m):.3e} {rel(T1, T0):.3e}")
