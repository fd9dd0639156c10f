"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/kidmp.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kidmp.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kidmp_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("kidmp_init", "kidmp_finalize", "kidmp_column", "kidmp_step", "kidmp_step_resident", "kidmp_upload",
                 "kidmp_download", "kidmp_step_device", "kidmp_diag", "kidmp_last_error", "kidmp_get_table",
                 "kidmp_kid_interface", "kidmp_mp_gt_driver"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    from kid_b200 import kidmp
    L = kidmp.load()
    for name in declared_functions():
        assert hasattr(L, name), "libkidmp.so does not export %s" % name
    # and the binding covers the same set
    assert sorted(kidmp.SYMBOLS) == declared_functions()


def test_no_torch_or_cxx_types_in_the_abi():
    txt = open(HEADER).read()
    assert 'extern "C"' in txt
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert "torch" not in code and "std::" not in code and "at::" not in code and "template" not in code


def test_built_for_sm_100a():
    from kid_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kid_b200.kidmp import Thompson, KidmpError
    with pytest.raises(KidmpError) as e:
        Thompson()
    assert "no CPU path" in str(e.value)


def test_product_does_not_reference_the_oracle():
    # the product path must never import, link or call anything under oracle/
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|thompson_oracle|kor_", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    from kid_b200 import build
    ldd = subprocess.run(["ldd", build.LIB], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_cxx_host_twin_builds_and_fails_cleanly_without_a_gpu(tmp_path):
    """kid_b200/host: the C++ restatement of `module mphys_thompson09n` links against the C ABI; without a CUDA device the
    no-argument interface call returns the library's error instead of crashing or falling back to a CPU path."""
    import numpy as np
    import torch
    from kid_b200 import build
    exe = build.build_host()
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    nx, nz = 2, 10
    src = tmp_path / "in.bin"
    np.zeros(7 * nx * nz + nz + 21 * nx * nz, np.float32).tofile(src)
    r = subprocess.run([exe, str(src), str(tmp_path / "out.bin"), str(nx), str(nz), "1.0", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


def test_library_is_built_from_the_sources_in_the_tree():
    """kidmp_build_id() = hash of the sources and flags (kid_b200/build.py::source_id): a shipped binary of other sources
    does not pass for the tree."""
    from kid_b200 import build, kidmp
    L = kidmp.load()
    assert L.kidmp_build_id().decode() == build.source_id() == build.library_id()
