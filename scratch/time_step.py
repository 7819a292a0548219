"""Variant timing: N device-resident HD steps (first is warm-up), stage_ms of each.  usage: time_step.py [WxHxD] [cs] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

w, h, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x256").split("x"))
cs = len(sys.argv) > 2 and sys.argv[2] == "cs"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
import numpy as np
_cache = f"/tmp/pair_{w}x{h}x{d}.npz"
if os.path.exists(_cache):
    _z = np.load(_cache); l, r = _z["l"], _z["r"]
else:
    l, r, _ = make_pair(w, h, d, SEED_BASE + 3)
    np.savez(_cache, l=l, r=r)
dev = torch.device("cuda", 0)
dl, dr = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
c.set_thread_stream(torch.cuda.current_stream().cuda_stream)
cc = c.GetCCType("GRD")
ptrs = ((dl.data_ptr(), w, h, 3 * w), (dr.data_ptr(), w, h, 3 * w))
import hashlib
for rep in range(reps):
    pc = c.PreCSPC(None, None, d, 35, 3, cc, 0.3, device_ptrs=ptrs) if cs else c.PreSSPC(None, None, d, 35, cc, device_ptrs=ptrs)
    pm = c.CSPatchMatch(None, None, d, 1, device_ptrs=ptrs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pm.PatchMatch(3, pc, False)
    e1.record()
    torch.cuda.synchronize()
    sm = pm.stage_ms()
    hsh = hashlib.sha1(pm.dis(0).tobytes() + pm.dis(1).tobytes()).hexdigest()[:12]
    print(os.environ.get("CSPM_LIB_PATH", "default"), f"rep{rep} total_ms {e0.elapsed_time(e1):.1f}",
          [round(v, 1) for v in sm], "dis_sha", hsh, flush=True)
