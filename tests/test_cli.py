"""python -m crossscalepatchmatch_b200: the reference's command line (main.cc:23-34, 57-139) with cv2 image I/O."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, "-m", "crossscalepatchmatch_b200", *args], cwd=cwd, env=env,
                          capture_output=True, text=True, timeout=600)


def test_flag_parsing_follows_gflags():
    from crossscalepatchmatch_b200.__main__ import DEFAULTS, parse_flags

    assert parse_flags([]) == DEFAULTS  # main.cc:23-34
    f = parse_flags(["--l_img_file=a.png", "-r_img_file", "b.png", "--max_dis=60", "--dis_scale", "4", "--cc_name=GRD",
                     "--use_cs", "--nouse_pp", "--reg_lambda=0.3", "--seed=0x10"])
    assert (f["l_img_file"], f["r_img_file"], f["max_dis"], f["dis_scale"], f["cc_name"]) == ("a.png", "b.png", 60, 4, "GRD")
    assert f["use_cs"] is True and f["use_pp"] is False and f["reg_lambda"] == 0.3 and f["seed"] == 16
    assert parse_flags(["--use_pp=false"])["use_pp"] is False and parse_flags(["--use_pp=1"])["use_pp"] is True
    for bad in (["--no_such_flag=1"], ["--max_dis"], ["--use_cs=maybe"]):
        with pytest.raises(ValueError):
            parse_flags(bad)


def test_missing_images_and_unknown_flags_fail_like_the_reference(tmp_path):
    r = _run(["--l_img_file=nope_l.png", "--r_img_file=nope_r.png", "--max_dis=16", "--dis_scale=4", "--cc_name=GRD"], tmp_path)
    assert r.returncode == 1 and "PatchMatch Stereo Matching" in r.stdout and "Error: can not open image" in r.stdout
    r = _run(["--bogus=1"], tmp_path)
    assert r.returncode == 1 and "unknown command line flag 'bogus'" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--use_cs", "--reg_lambda=0.3", "--use_pp"]])
def test_cli_equals_the_api(tmp_path, extra):
    import cv2
    import crossscalepatchmatch_b200 as c
    from crossscalepatchmatch_b200.synth import make_pair

    w, h, d = 80, 64, 16
    l, r, _ = make_pair(w, h, d, 0xC5B20000)
    cv2.imwrite(str(tmp_path / "l.png"), l)
    cv2.imwrite(str(tmp_path / "r.png"), r)
    res = _run(["--l_img_file=l.png", "--r_img_file=r.png", "--l_dis_file=ld.png", "--r_dis_file=rd.png", f"--max_dis={d}",
                "--dis_scale=4", "--cc_name=GRD", *extra], tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Total Time:" in res.stdout
    cs = "--use_cs" in extra
    cc = c.GetCCType("GRD")
    pc = c.PreCSPC(l, r, d, 35, 5, cc, 0.3) if cs else c.PreSSPC(l, r, d, 35, cc)
    pm = c.CSPatchMatch(l, r, d, 4)
    pm.PatchMatch(3, pc, "--use_pp" in extra)
    for v, name in ((0, "ld.png"), (1, "rd.png")):
        got = cv2.imread(str(tmp_path / name), cv2.IMREAD_UNCHANGED)
        assert got is not None and got.dtype == np.uint8 and got.shape == (h, w)
        assert np.array_equal(got, pm.dis(v)), name
