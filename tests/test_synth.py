"""The input generator (include/cspm_synth.h, splitmix64, dependency-free C): the Python binding and a C++ program
that includes the header produce byte-identical pairs; images are pinned by digest; loading it does not pull in
the product library."""
import hashlib
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_is_deterministic_and_pinned():
    from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

    l, r, gt = make_pair(96, 64, 20, SEED_BASE)
    l2, r2, gt2 = make_pair(96, 64, 20, SEED_BASE)
    assert np.array_equal(l, l2) and np.array_equal(r, r2) and np.array_equal(gt, gt2)
    assert l.shape == (64, 96, 3) and l.dtype == np.uint8 and gt.min() >= 1.0 and gt.max() <= 19.0
    assert not np.array_equal(make_pair(96, 64, 20, SEED_BASE + 1)[0], l)
    # pinned: a change of the generator invalidates tests/golden/cfg_*.npz (they store the digests of their inputs)
    assert hashlib.sha1(l.tobytes() + r.tobytes()).hexdigest() == PIN_96x64x20, hashlib.sha1(l.tobytes() + r.tobytes()).hexdigest()
    # the right view is the left one shifted by the disparity where nothing is occluded
    y, x = 40, 70
    d = int(round(gt[y, x]))
    assert np.abs(l[y, x].astype(int) - r[y, x - d].astype(int)).max() <= 40


def test_variants_share_the_scene():
    from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

    l0, r0, g0 = make_pair(96, 64, 20, SEED_BASE + 3)
    l1, r1, g1 = make_pair(96, 64, 20, SEED_BASE + 3, variant=1)
    l2, r2, g2 = make_pair(96, 64, 20, SEED_BASE + 3, variant=2)
    assert np.array_equal(g0, g1) and np.array_equal(g0, g2), "variants keep the ground-truth disparity"
    assert not np.array_equal(l0, l1) and not np.array_equal(l1, l2) and not np.array_equal(r0, r1)
    assert abs(float(l0.mean()) - float(l1.mean())) < 6 and abs(float(l0.std()) - float(l1.std())) < 6


PIN_96x64x20 = "d083e3bf1d80963e627df612da9043894bd22fa4"


def test_importing_synth_does_not_load_the_product():
    code = ("import sys; import crossscalepatchmatch_b200.synth as s; s.make_pair(16, 8, 4); "
            "assert 'crossscalepatchmatch_b200._lib' not in sys.modules; "
            "assert not any('libcspm_b200' in l for l in open('/proc/self/maps')); print('ok')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_cpp_program_generates_the_same_pair(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no host compiler")
    src = tmp_path / "g.cc"
    src.write_text('#include <cstdio>\n#include <vector>\n#include "cspm_synth.h"\n'
                   'int main() { const int w = 80, h = 50, d = 16; std::vector<uint8_t> l(3 * w * h), r(3 * w * h);\n'
                   '  if (cspm_synth_pair(w, h, d, CSPM_SYNTH_SEED_BASE + 2, l.data(), r.data(), nullptr)) return 1;\n'
                   '  fwrite(l.data(), 1, l.size(), stdout); fwrite(r.data(), 1, r.size(), stdout); return 0; }\n')
    exe = tmp_path / "g"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = subprocess.run([str(exe)], capture_output=True).stdout
    from crossscalepatchmatch_b200.synth import SEED_BASE, make_pair

    l, rr, _ = make_pair(80, 50, 16, SEED_BASE + 2)
    assert raw == l.tobytes() + rr.tobytes()
