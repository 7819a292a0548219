"""Generate tests/golden/cfg_<name>.npz: what the REFERENCE'S OWN CODE
(oracle/_ref: unmodified /root/reference sources + oracle/cvshim) computes on the
BASELINE.json configurations at their real sizes and disparity ranges
(C1 450x375 D=60, C2 = C1 + S=3 lambda=0.3, C3 1242x375 D=128 S=3,
C4 1920x1080 D=256, C4cs = C4 + S=3).  Per config: sha1 of the two generated
input images, the two uint8 disparity maps, and sha1 digests of the f64 normal /
param / min_cost fields of both views after PatchMatch(3, pc, use_pp=false).

The reference's OpenMP row loops run on all cores; time(NULL) is interposed in
the shim's mode 2, which hands every row the RNG stream id of the single-threaded
run (oracle/cvshim/opencv2/opencv.hpp), so the output does not depend on the
thread count (tests/test_oracle_vs_ref.py::test_ref_parallel_rng_mode).

Run (authoring container only; C4 takes ~1 h on 8 cores, C4cs ~3 h):
    python tests/golden/make_config_vectors.py C1 C2 C3 C4
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import orc
from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

# name -> (w, h, max_dis, scale_num, reg_lambda, dis_scale, seed)
CONFIGS = {
    "C1": (450, 375, 60, 0, 0.0, 4, SEED_BASE + 0),
    "C2": (450, 375, 60, 3, 0.3, 4, SEED_BASE + 1),
    "C3": (1242, 375, 128, 3, 0.3, 1, SEED_BASE + 2),
    "C4": (1920, 1080, 256, 0, 0.0, 1, SEED_BASE + 3),
    "C4cs": (1920, 1080, 256, 3, 0.3, 1, SEED_BASE + 3),
}
ITERS, WND, PM_SEED = 3, 35, 0x5EED


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(names):
    lib = orc.ref_fast() if orc.ref_fast_available() else orc.ref()
    lib.set_time_mode(2)
    threads = len(os.sched_getaffinity(0))
    os.makedirs(os.path.join(ROOT, "profiles", "r02_configs"), exist_ok=True)
    for name in names:
        w, h, d, S, lam, dis_scale, seed = CONFIGS[name]
        l, r, _ = make_pair(w, h, d, seed)
        t0 = time.perf_counter()
        pc = orc.PlaneCost(lib, l, r, d, WND, S, lam)
        pm = orc.PatchMatch(lib, l, r, d, dis_scale, seed=PM_SEED)
        pm.patchmatch(ITERS, pc, False, threads)
        secs = time.perf_counter() - t0
        out = {"whd": np.array([w, h, d]), "scale_num": np.array(S), "reg_lambda": np.array(lam), "dis_scale": np.array(dis_scale),
               "seed": np.array(seed, np.uint64), "pm_seed": np.array(PM_SEED, np.uint64), "iters": np.array(ITERS),
               "sha_l": np.array(sha(l)), "sha_r": np.array(sha(r))}
        for v in (0, 1):
            n, p, c = pm.planes(v)
            out[f"dis{v}"] = pm.dis(v)
            out[f"sha_normal{v}"], out[f"sha_param{v}"], out[f"sha_mincost{v}"] = np.array(sha(n)), np.array(sha(p)), np.array(sha(c))
        np.savez_compressed(os.path.join(HERE, f"cfg_{name}.npz"), **out)
        line = {"impl": "reference", "config": name, "workload": f"synthetic {w}x{h} pair, max_dis={d}, {ITERS} iters, " +
                (f"cross-scale S={S} lambda={lam}" if S else "single-scale") + ", use_pp=false, FULL image",
                "seconds": secs, "value": w * h * d / secs / 1e6, "unit": "Mpixel*disp/s", "cores": threads,
                "kind": "reference", "build": os.path.basename(lib.path),
                "where": "authoring container (not the GPU box); by-product of the fixture run, other jobs may have shared the cores"}
        json.dump(line, open(os.path.join(ROOT, "profiles", "r02_configs", f"cfg_{name}_cpu_full_authoring.json"), "w"))
        print(json.dumps(line), flush=True)
        pm.close()
        pc.close()


if __name__ == "__main__":
    main(sys.argv[1:] or ["C1", "C2", "C3"])
