#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 PatchMatch engine.

Metric (BASELINE.json): Mpixel*disparities/s = W*H*max_dis / t / 1e6 per stereo
pair on synthetic 1920x1080, max_dis=256 pairs, 3 PatchMatch iterations
(configs[3]); t covers the region the reference times (main.cc:92-123):
plane-cost construction + CSPatchMatch construction + PatchMatch(3, pc, false).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--use-cs] [--impl reference]

One "step" = one stereo pair.  N>1: launched under torchrun, one rank per GPU,
every rank works on its own pair (independent units, no data-path collective;
"scaling": "weak"); NCCL is used only for the barrier and the max-over-ranks
of the elapsed time.

`value`  : device-resident inputs, CUDA events on the stream the library
           launches on (PyTorch's current stream, injected through
           cspm_set_thread_stream), one event pair per step, L2 flushed between
           steps; max over ranks.
`e2e`    : the same pair through the host-pointer API (pinned host images in,
           two uint8 disparity maps out), wall clock with a device synchronise
           on both sides, H2D/D2H inside the timed region.
`roofline`: dominant kernel = the PlaneRefinement launch; achieved = 19 B x
           border-clipped window taps of that launch / its CUDA-event duration
           (SURVEY.md §8d no-reuse gather model), peak = MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own CPU code (oracle/_ref: unmodified reference
           sources + OpenCV shim) or, if that .so is absent, the oracle port,
           on a bounded crop of the same pair with all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_HD, H_HD, D_HD = 1920, 1080, 256
ITERS = 3
WND = 35
PM_SEED = 0x5EED
BYTES_PER_TAP = 19  # 3 B guide pixel + 2 x 8 B cost samples (pre_ss_pc.cc:93,109-110)
# BASELINE.json configs: (w, h, max_dis, cross-scale S=3 lambda=0.3)
CONFIGS = {"C1": (450, 375, 60, False), "C2": (450, 375, 60, True), "C3": (1242, 375, 128, True),
           "C4": (1920, 1080, 256, False), "C4cs": (1920, 1080, 256, True)}
SCHEDULE_NAMES = {0: "raster", 1: "red-black (4 neighbours)", 2: "red-black (8 neighbours)"}
CPU_CROP = (384, 216)  # bounded CPU sample (w, h), same max_dis: ~10-30 s of CPU work per pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--use-cs", action="store_true", help="cross-scale plane cost (PreCSPC, S=3, lambda=0.3)")
    ap.add_argument("--size", default=None, help="WxHxD override (debug only; the headline is 1920x1080x256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE.json config preset (default: C4, the headline)")
    ap.add_argument("--cpu-full", action="store_true", help="CPU arm on the whole image instead of the bounded crop (BASELINE.md table)")
    ap.add_argument("--schedule", type=int, default=0, choices=[0, 1, 2],
                    help="0 raster = the reference's order (headline), 1 / 2 red-black with 4 / 8 neighbours (throughput schedules)")
    ap.add_argument("--pairs-in-flight", type=int, default=1,
                    help="throughput mode: a step is a batch of P independent pairs per GPU (cspm_pm_patchmatch_batch: their "
                         "spatial sweeps share one launch); the headline is P = 1")
    ap.add_argument("--no-worst-case", action="store_true", help="skip the extra step with early termination disabled")
    ap.add_argument("--no-tap-count", action="store_true", help="skip the executed-tap count (instrumented build)")
    return ap.parse_args()


def workload(args):
    if args.config:
        w, h, d, cs = CONFIGS[args.config]
        args.size = f"{w}x{h}x{d}"
        args.use_cs = cs
    if args.size:
        w, h, d = (int(v) for v in args.size.lower().split("x"))
    else:
        w, h, d = W_HD, H_HD, D_HD
    name = f"synthetic {w}x{h} pair, max_dis={d}, {ITERS} PatchMatch iters, wnd=35, GRD, " + (
        "cross-scale S=3 lambda=0.3" if args.use_cs else "single-scale") + ", use_pp=false"
    return w, h, d, name


# ---------------------------------------------------------------- CPU arms
def cpu_lib():
    """The reference's own code if its build travelled with the repo (oracle/_ref: unmodified reference sources +
    OpenCV shim; the -O3 -march=x86-64-v3 timing build when present), else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    if orc.ref_fast_available():
        return orc.ref_fast(), "reference", "-O3 -march=x86-64-v3 -fopenmp -ffp-contract=off (oracle/Makefile)"
    if orc.ref_available():
        return orc.ref(), "reference", "-O2 -fopenmp -ffp-contract=off (oracle/Makefile)"
    return orc.oracle(), "port", "-O2 -fopenmp -ffp-contract=off (oracle/Makefile)"


def cpu_run(l, r, d, use_cs, nthreads, want_result=False):
    """One pass of the reference's timed region (main.cc:92-123) on the CPU; returns (seconds, kind, build, result).
    The reference's OpenMP row loops use `nthreads`; its time(NULL) seeds are interposed in the shim's mode 2
    (every row gets the RNG stream id of the library's RNG contract, whatever the thread count), so the result
    can be compared bit for bit with the GPU's."""
    lib, kind, build = cpu_lib()
    import orc
    if kind == "reference":
        lib.set_time_mode(2)
    t0 = time.perf_counter()
    pc = orc.PlaneCost(lib, l, r, d, WND, 3 if use_cs else 0, 0.3 if use_cs else 0.0)
    pm = orc.PatchMatch(lib, l, r, d, 1, seed=PM_SEED)
    pm.patchmatch(ITERS, pc, False, nthreads)
    t = time.perf_counter() - t0
    res = None
    if want_result:
        res = [pm.dis(0), pm.dis(1)] + [a for v in (0, 1) for a in pm.planes(v)]
    pm.close()
    pc.close()
    return t, kind, build, res


FULL_CPU = {"on": False}


def cpu_sample(w, h, d, seed):
    from crossscalepatchmatch_b200.synth import make_pair

    cw, ch = (w, h) if FULL_CPU["on"] else (min(CPU_CROP[0], w), min(CPU_CROP[1], h))
    l, r, _ = make_pair(w, h, d, seed)
    x0, y0 = (w - cw) // 2, (h - ch) // 2
    return np.ascontiguousarray(l[y0:y0 + ch, x0:x0 + cw]), np.ascontiguousarray(r[y0:y0 + ch, x0:x0 + cw]), cw, ch


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def sample_text(cw, ch, w, h, d):
    if (cw, ch) == (w, h):
        return f"the FULL {w}x{h} pair, max_dis={d}, full pipeline ({ITERS} iters)", "full"
    return f"{cw}x{ch} centre crop of the {w}x{h} pair, max_dis={d}, full pipeline ({ITERS} iters)", "crop"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    w, h, d, name = workload(args)
    from crossscalepatchmatch_b200.synth import SEED_BASE

    l, r, cw, ch = cpu_sample(w, h, d, SEED_BASE + 3)
    cores = host_cores()
    kind, build = "port", ""
    for _ in range(min(args.warmup, 1)):  # one CPU warm-up pass is plenty (no JIT, no caches to prime)
        cpu_run(l, r, d, args.use_cs, cores)
    ts = []
    for _ in range(args.steps):
        t, kind, build, _ = cpu_run(l, r, d, args.use_cs, cores)
        ts.append(t)
    total = sum(ts)
    value = cw * ch * d * args.steps / total / 1e6
    sample, sample_kind = sample_text(cw, ch, w, h, d)
    line = {
        "impl": "reference", "metric": "Mpixel*disparities/s", "value": value, "unit": "Mpixel*disp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "sample": sample, "sample_kind": sample_kind},
        "cpu_baseline": {"value": value, "unit": "Mpixel*disp/s", "cores": cores, "kind": kind, "sample": sample,
                         "sample_kind": sample_kind, "build": build},
        "e2e": {"value": value, "unit": "Mpixel*disp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    # this arm is CPU only: the product library must not even be mapped into the process
    line["product_so_loaded"] = any("libcspm_b200" in m for m in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None
    print(json.dumps(line), flush=True)


def binding_limit(onchip):
    """The on-chip resource of the ncu summary that is closest to its peak: what actually bounds the dominant kernel
    (the 19 B/tap HBM model exceeds 1 by construction once windows are served from L1/L2)."""
    if not isinstance(onchip, dict):
        return None
    names = {"issue_active_pct": "instruction issue slots", "l1_data_pipe_pct": "L1 data pipe (LSU wavefronts)",
             "fp64_pipe_pct": "fp64 pipe", "xu_pipe_pct": "XU pipe (int<->f64 conversions)", "alu_pipe_pct": "ALU pipe"}
    cands = {k: float(onchip[k]) for k in names if isinstance(onchip.get(k), (int, float))}
    if not cands:
        return None
    top = max(cands, key=cands.get)
    return {"resource": names[top], "frac_of_peak": cands[top] / 100.0,
            "all": {names[k]: round(v / 100.0, 4) for k, v in sorted(cands.items(), key=lambda kv: -kv[1])},
            "source": onchip.get("source")}


def count_executed_taps(args, device):
    """Window taps the PlaneRefinement launches actually execute: a second process loads the instrumented build
    (libcspm_b200_count.so, -DCSPM_COUNT_TAPS=1: same kernels + a counter) and replays the step stage by stage."""
    lib = os.path.join(ROOT, "crossscalepatchmatch_b200", "libcspm_b200_count.so")
    if not os.path.exists(lib):
        return None
    w, h, d, _ = workload(args)
    code = f"""
import ctypes, json, sys
sys.path.insert(0, {ROOT!r})
import crossscalepatchmatch_b200 as c
from crossscalepatchmatch_b200 import _lib
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair
l, r, _ = make_pair({w}, {h}, {d}, SEED_BASE + 3)
cc = c.GetCCType("GRD")
pc = c.PreCSPC(l, r, {d}, {WND}, 3, cc, 0.3, device={device}) if {bool(args.use_cs)} else c.PreSSPC(l, r, {d}, {WND}, cc, device={device})
pm = c.CSPatchMatch(l, r, {d}, 1, seed={PM_SEED}, device={device}, schedule={args.schedule})
n = ctypes.c_uint64(0)
pm.stage(c.STAGE_INIT, 0, pc)
tot = 0
for it in range({ITERS}):
    pm.stage(c.STAGE_SPATIAL, it, pc); pm.stage(c.STAGE_VIEW, it, pc)
    _lib.check(_lib.dll.cspm_debug_taps_executed(ctypes.byref(n), 1))
    pm.stage(c.STAGE_REFINE, it, pc)
    _lib.check(_lib.dll.cspm_debug_taps_executed(ctypes.byref(n), 1))
    tot += n.value
print(json.dumps({{"refine_taps_per_launch": tot // {ITERS}}}))
"""
    try:
        out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "CSPM_LIB_PATH": lib}, capture_output=True, text=True, timeout=600)
        res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        res["source"] = "libcspm_b200_count.so (-DCSPM_COUNT_TAPS=1) on the same inputs, mean of the step's 3 refinement launches"
        return res
    except Exception as e:  # the figure is an annotation: never fail the bench line over it
        print(f"executed-tap count unavailable: {e}", file=sys.stderr)
        return None


# ------------------------------------------------------------- clock sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import crossscalepatchmatch_b200 as c
    from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available() and c.device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w, h, d, name = workload(args)

    # this rank's pairs: pair k of the job is variant k of the configuration's seed — the same scene under a
    # different, statistically equivalent texture (include/cspm_synth.h), i.e. distinct images with the same amount
    # of work, so that "weak scaling" really holds the per-GPU work fixed; pinned host copies for the e2e leg
    P = max(1, args.pairs_in_flight)
    host, devp = [], []
    for k in range(P):
        l, r, _ = make_pair(w, h, d, SEED_BASE + 3, variant=rank + 1000 * k)
        hl = torch.from_numpy(l).pin_memory()
        hr = torch.from_numpy(r).pin_memory()
        host.append((hl, hr))
        devp.append((hl.to(dev), hr.to(dev)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    stream = torch.cuda.current_stream()
    c.set_thread_stream(stream.cuda_stream)  # library launches on torch's current stream
    cc = c.GetCCType("GRD")
    S, lam = (3, 0.3) if args.use_cs else (0, 0.0)

    def make_pc(lp, rp, dev_ptrs):
        kw = {"device": local}
        if dev_ptrs:
            kw["device_ptrs"] = ((lp.data_ptr(), w, h, 3 * w), (rp.data_ptr(), w, h, 3 * w))
            lp = rp = None
        else:
            lp, rp = lp.numpy(), rp.numpy()
        if S == 0:
            return c.PreSSPC(lp, rp, d, WND, cc, **kw)
        return c.PreCSPC(lp, rp, d, WND, S, cc, lam, **kw)

    launches = {"n": 0}
    last = {}

    def run_pairs(pms, pcs):
        if P == 1:
            pms[0].PatchMatch(ITERS, pcs[0], False)
        else:
            c.PatchMatchBatch(pms, pcs, ITERS, False)

    def step_device():
        """reference's timed region with inputs resident in HBM (P pairs: the region once per pair, batched)"""
        pcs = [make_pc(dl, dr, True) for dl, dr in devp]
        pms = [c.CSPatchMatch(None, None, d, 1, seed=PM_SEED, device=local, schedule=args.schedule,
                              device_ptrs=((dl.data_ptr(), w, h, 3 * w), (dr.data_ptr(), w, h, 3 * w))) for dl, dr in devp]
        run_pairs(pms, pcs)
        launches["n"] += sum(pc.stats()["kernel_launches"] for pc in pcs) + sum(pm.stats()["kernel_launches"] for pm in pms)
        st = {k: sum(pm.stats()[k] for pm in pms) for k in pms[0].stats()}
        last.update(stage_ms=[v / P for v in pms[0].stage_ms()], taps_eval=pcs[0].taps_per_image_eval(), n_ref=pms[0].n_ref(), stats=st)
        return pcs, pms

    def step_e2e():
        pcs = [make_pc(hl, hr, False) for hl, hr in host]
        pms = [c.CSPatchMatch(hl.numpy(), hr.numpy(), d, 1, seed=PM_SEED, device=local, schedule=args.schedule) for hl, hr in host]
        run_pairs(pms, pcs)
        return [(pm.dis(c.kLeft), pm.dis(c.kRight)) for pm in pms]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    launches["n"] = 0

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    refine_ms = []
    for k in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (outside the timed region)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        refine_ms.append(last["stage_ms"][c.STAGE_REFINE] / ITERS)
    barrier()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    if os.environ.get("CSPM_BENCH_DEBUG"):
        print(f"[rank {rank}] per-step device ms: {[round(a.elapsed_time(b), 1) for a, b in ev]} stage_ms {last['stage_ms']}", file=sys.stderr, flush=True)
    gpu_launches = launches["n"]

    # end-to-end leg: host buffers in, disparity maps out
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None

    from crossscalepatchmatch_b200 import shard

    units, (t_dev, t_e2e) = shard.aggregate(float(w * h * d * args.steps * P), [t_dev, t_e2e], device=dev)
    gpu_launches, _ = shard.aggregate(float(gpu_launches), [0.0], device=dev)  # whole job: summed over ranks

    if rank == 0:
        value = units / t_dev / 1e6
        e2e = units / t_e2e / 1e6
        # roofline of the dominant kernel (k_refine, one launch per iteration, both views)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        taps_launch = 2 * last["n_ref"] * last["taps_eval"]
        ms = float(np.mean(refine_ms))
        achieved = BYTES_PER_TAP * taps_launch / (ms * 1e-3) / 1e9
        key = "cs" if args.use_cs else "ss"
        traffic, onchip = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(key)
        opath = os.path.join(ROOT, "profiles", "r02_onchip.json")
        if os.path.exists(opath):  # ncu --set full summary of the benched kernel version (profiles/summarize.py)
            onchip = json.load(open(opath)).get("k_refine_" + key)
        stage_ms = dict(zip(["init", "spatial", "view", "refine", "plane2disp", "postproc"], last["stage_ms"]))
        roofline = {"bound": "hbm", "kernel": "k_refine", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": BYTES_PER_TAP * taps_launch, "launch_ms": ms,
                    "taps_nominal": int(taps_launch),
                    "model": "no-reuse gather, 19 B per border-clipped window tap (SURVEY.md 8d): an upper bound on the "
                             "bytes a cache-less machine would move, not DRAM traffic; windows of neighbouring pixels "
                             "overlap 34/35 and are served by L1/L2, so frac > 1 is expected and the binding limits "
                             "are the on-chip ones in `onchip`",
                    "onchip": onchip}
        if traffic:
            roofline["dram_frac"] = traffic / (ms * 1e-3) / 1e9 / peak  # measured DRAM bytes / launch time / peak
        roofline["binding"] = binding_limit(onchip)  # the limiter that is really closest to its peak (< 1)
        # taps the launch really executed (exact early termination skips the rest): counted by the instrumented
        # build of the same kernels on the same inputs, outside the timed region
        if not args.no_tap_count:
            ex = count_executed_taps(args, local)
            if ex:
                roofline["taps_executed"] = ex["refine_taps_per_launch"]
                roofline["frac_executed"] = BYTES_PER_TAP * ex["refine_taps_per_launch"] / (ms * 1e-3) / 1e9 / peak
                roofline["taps_executed_source"] = ex["source"]
        line = {
            "metric": "Mpixel*disparities/s", "value": value, "unit": "Mpixel*disp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "pairs_per_gpu_per_step": P, "pairs_in_flight": P, "sharding": "one stereo pair per GPU, no data-path collective",
                       "schedule": SCHEDULE_NAMES[args.schedule] + (" (bit-exact reference order)" if args.schedule == 0 else " (throughput schedule, not the reference's order)"),
                       "l2": "flushed between timed steps (256 MiB write)"},
            "e2e": {"value": e2e, "unit": "Mpixel*disp/s", "h2d_bytes_per_step": 4 * w * h * 3 * P, "d2h_bytes_per_step": 2 * w * h * P,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": roofline,
            "stage_ms": stage_ms,
            "window_taps_per_step": int(last["stats"]["window_taps"]),
            "stage_ms_note": "per pair (batch totals / pairs_in_flight)" if P > 1 else "one pair",
            "spatial_sweep": {k: int(last["stats"][k]) for k in ("spatial_tasks", "spatial_candidates", "spatial_adopted", "spatial_exact_tasks")},
        }
        if world == 1 and not args.no_worst_case and P == 1:
            # data-independent figure: the same step with exact early termination switched off (every window summed in full)
            os.environ["CSPM_NO_EARLY_TERM"] = "1"
            step_device()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record(stream)
            step_device()
            b.record(stream)
            torch.cuda.synchronize()
            del os.environ["CSPM_NO_EARLY_TERM"]
            wms = a.elapsed_time(b)
            line["worst_case"] = {"what": "exact early termination disabled (CSPM_NO_EARLY_TERM=1): every window summed in full, "
                                          "throughput independent of the image content", "ms_per_step": wms,
                                  "value": w * h * d / wms / 1e3, "unit": "Mpixel*disp/s",
                                  "refine_frac_of_model": BYTES_PER_TAP * taps_launch / (last["stage_ms"][c.STAGE_REFINE] / ITERS * 1e-3) / 1e9 / peak}
        if world == 1 and not args.no_cpu_baseline:
            cl, cr, cw, ch = cpu_sample(w, h, d, SEED_BASE + 3)
            cores = host_cores()
            t, kind, build, ref_out = cpu_run(cl, cr, d, args.use_cs, cores, want_result=True)
            sample, sample_kind = sample_text(cw, ch, w, h, d)
            line["cpu_baseline"] = {"value": cw * ch * d / t / 1e6, "unit": "Mpixel*disp/s", "cores": cores, "kind": kind,
                                    "sample": sample + f", 1 pass ({t:.1f} s)", "sample_kind": sample_kind, "build": build}
            # parity on the very sample the CPU arm just computed: the GPU runs the same crop through the C ABI and
            # both disparity maps and every plane / cost field must be bit-identical
            if S == 0:
                pcc = c.PreSSPC(cl, cr, d, WND, cc, device=local)
            else:
                pcc = c.PreCSPC(cl, cr, d, WND, S, cc, lam, device=local)
            pmc = c.CSPatchMatch(cl, cr, d, 1, seed=PM_SEED, device=local, schedule=0)
            pmc.PatchMatch(ITERS, pcc, False)
            got = [pmc.dis(0), pmc.dis(1)] + [a for v in (0, 1) for a in pmc.planes(v)]
            line["parity_checked"] = bool(all(np.array_equal(x, y) for x, y in zip(got, ref_out)))
            line["parity_sample"] = sample + " (uint8 disparity maps, f64 normals / params / min_cost of both views, bit for bit" + (
                "; GPU arm run with the raster schedule for this check)" if args.schedule else ")")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    FULL_CPU["on"] = args.cpu_full
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
