"""The C-ABI library loads without a GPU and exports every symbol that
include/cspm_c.h declares; compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cspm_c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cspm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import crossscalepatchmatch_b200 as c

    dll = ctypes.CDLL(c.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/cspm_c.h but not exported"
    # and the Python binding covers exactly the declared surface
    assert sorted(c._lib.SIGNATURES) == names


def test_no_gpu_fails_loudly():
    import crossscalepatchmatch_b200 as c

    if c.device_count() > 0:
        pytest.skip("a GPU is present")
    img = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(c.CspmError) as e:
        c.PreSSPC(img, img, 4, 35, c.GetCCType("GRD"))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(c.CspmError):
        c.CSPatchMatch(img, img, 4, 1)
    with pytest.raises(c.CspmError):
        c.GrdCC().buildCV(np.zeros((4, 4, 3)), np.zeros((4, 4, 3)), 2)


def test_product_does_not_import_oracle():
    """The product package and the C sources must not reference oracle/."""
    pkg = os.path.join(ROOT, "crossscalepatchmatch_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/ and is test", "") and "import orc" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read().lower() or f == "cspm_c.h"


def test_argument_validation_without_gpu():
    import crossscalepatchmatch_b200 as c

    assert c.GetCCType("GRD") is not None and c.GetCCType("CEN") is not None
    assert c.GetCCType("CG") is None and c.GetCCType("nope") is None  # main.cc:47-54
    with pytest.raises(c.CspmError):
        c.PreSSPC(np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8), 4, 35, c.GetCCType("GRD"))


def test_batch_entry_point_validates_without_gpu():
    import crossscalepatchmatch_b200 as c
    from crossscalepatchmatch_b200 import _lib

    assert _lib.dll.cspm_pm_patchmatch_batch(None, None, 0, 1, 0) == _lib.EINVAL
    assert _lib.dll.cspm_last_error_code() == _lib.EINVAL and "batch" in _lib.last_error()
    with pytest.raises(c.CspmError):
        c.PatchMatchBatch([], [], 1)


def test_headers_compile_as_c_and_cpp(tmp_path):
    """include/cspm_c.h must be plain C (the boundary is a C ABI); include/cspm.hpp is self-contained C++17."""
    import shutil
    import subprocess

    inc = os.path.join(ROOT, "include")
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    c_src = tmp_path / "t.c"
    c_src.write_text('#include "cspm_c.h"\nint main(void) { return cspm_device_count() < 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(c_src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp_src = tmp_path / "t.cc"
    cpp_src.write_text('#include "cspm.hpp"\nint main() { cspm::Plane p; (void)p; return 0; }\n')
    r = subprocess.run([gxx, "-std=c++17", "-Wall", "-fsyntax-only", "-I", inc, str(cpp_src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
