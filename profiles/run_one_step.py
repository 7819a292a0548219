"""Profiling driver: exactly one device-resident step (the bench.py step_device
body) so ncu sees one pass of the hot path.  usage: run_one_step.py [WxHxD] [ss|cs] [schedule 0|1|2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

w, h, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x256").split("x"))
cs = len(sys.argv) > 2 and sys.argv[2] == "cs"
l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
dev = torch.device("cuda", 0)
dl, dr = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
c.set_thread_stream(torch.cuda.current_stream().cuda_stream)
cc = c.GetCCType("GRD")
ptrs = ((dl.data_ptr(), w, h, 3 * w), (dr.data_ptr(), w, h, 3 * w))
pc = c.PreCSPC(None, None, d, 35, 3, cc, 0.3, device_ptrs=ptrs) if cs else c.PreSSPC(None, None, d, 35, cc, device_ptrs=ptrs)
schedule = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pm = c.CSPatchMatch(None, None, d, 1, device_ptrs=ptrs, schedule=schedule)
pm.PatchMatch(3, pc, False)
torch.cuda.synchronize()
print("stage_ms", pm.stage_ms(), pm.stats())
