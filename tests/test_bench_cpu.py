"""bench.py's CPU arm (--impl reference) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "64x40x12",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel*disp/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["cpu_baseline"]
    assert d["product_so_loaded"] is False, "the reference arm must not map libcspm_b200.so"
    assert d["cpu_baseline"]["sample_kind"] in ("crop", "full") and "build" in d["cpu_baseline"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "64x40x12",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_binding_limit_from_the_committed_ncu_summary():
    """roofline.binding of the bench line: the on-chip resource closest to its peak, from profiles/r02_onchip.json."""
    sys.path.insert(0, ROOT)
    import bench

    onchip = json.load(open(os.path.join(ROOT, "profiles", "r02_onchip.json")))
    for key in ("k_refine_ss", "k_refine_cs"):
        b = bench.binding_limit(onchip[key])
        assert b and 0.3 < b["frac_of_peak"] < 1.0 and b["resource"] in b["all"] and "r02_ncu_k_refine" in b["source"]
        assert b["frac_of_peak"] == max(b["all"].values()) or abs(b["frac_of_peak"] - max(b["all"].values())) < 1e-4
    assert bench.binding_limit(None) is None and bench.binding_limit({}) is None
