"""BASELINE.json configurations at their REAL sizes and disparity ranges, through the C ABI, against what the
reference's own code computed (tests/golden/cfg_<name>.npz, written by tests/golden/make_config_vectors.py from
oracle/_ref = the unmodified reference sources): both uint8 disparity maps byte for byte, and the f64 normal /
param / min_cost fields of both views by sha1.  C1 450x375 D=60; C2 = C1 + cross-scale S=3 lambda=0.3;
C3 1242x375 D=128 S=3; C4 1920x1080 D=256 (the benchmarked workload); C4cs = C4 + S=3."""
import hashlib
import os

import numpy as np
import pytest

from crossscalepatchmatch_b200.synth import make_pair

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_config(name, env=None):
    import crossscalepatchmatch_b200 as c

    path = os.path.join(GOLDEN, f"cfg_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (make_config_vectors.py {name})")
    g = np.load(path)
    w, h, d = (int(v) for v in g["whd"])
    S, lam, dis_scale = int(g["scale_num"]), float(g["reg_lambda"]), int(g["dis_scale"])
    l, r, _ = make_pair(w, h, d, int(g["seed"]))
    assert sha(l) == str(g["sha_l"]) and sha(r) == str(g["sha_r"]), "the input generator no longer reproduces the fixture's images"
    cc = c.GetCCType("GRD")
    pc = c.PreSSPC(l, r, d, 35, cc) if S == 0 else c.PreCSPC(l, r, d, 35, S, cc, lam)
    pm = c.CSPatchMatch(l, r, d, dis_scale, seed=int(g["pm_seed"]))
    pm.PatchMatch(int(g["iters"]), pc, False)
    for v in (0, 1):
        dis = pm.dis(v)
        nbad = int((dis != g[f"dis{v}"]).sum())
        assert nbad == 0, f"{name}: {nbad} of {dis.size} disparities of view {v} differ from the reference's"
        n, p, cst = pm.planes(v)
        assert sha(n) == str(g[f"sha_normal{v}"]), f"{name}: plane normals of view {v}"
        assert sha(p) == str(g[f"sha_param{v}"]), f"{name}: plane parameters of view {v}"
        assert sha(cst) == str(g[f"sha_mincost{v}"]), f"{name}: aggregated costs of view {v}"
    return pm.stats()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C4cs"])
def test_config_matches_reference(name):
    st = run_config(name)
    assert st["spatial_exact_tasks"] <= st["spatial_tasks"]


def test_c1_other_sweep_implementations(monkeypatch):
    """The round-1 ring sweep and the wavefront sweep's exact path (every task forced down it) give the same C1."""
    for env in ({"CSPM_SPATIAL": "ring"}, {"CSPM_WAVE_EXACT": "1"}):
        import subprocess
        import sys

        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_configs as t; t.run_config('C1'); print('ok')"
                % (os.path.dirname(GOLDEN), os.path.dirname(os.path.dirname(GOLDEN))))
        out = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0 and "ok" in out.stdout, f"{env}: {out.stdout[-2000:]} {out.stderr[-2000:]}"
