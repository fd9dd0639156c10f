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


def test_fortran_shim_binds_only_what_the_header_declares():
    """kid_b200/fortran/mphys_thompson09n.f90 cannot be compiled here (no Fortran compiler), so its iso_c_binding side is checked on
    the text: every bind(C, name=...) is a function of include/kidmp.h, and the bind(C) derived types list the members of the C
    structs in the same order (kidmp_config, kidmp_kid_columns, kidmp_wrf_fields, kidmp_wrf_aerosols)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    f90 = open(os.path.join(root, "kid_b200", "fortran", "mphys_thompson09n.f90")).read()
    hdr = open(os.path.join(root, "include", "kidmp.h")).read()
    declared = set(re.findall(r"\b(kidmp_[a-z0-9_]+)\s*\(", hdr))
    bound = re.findall(r"bind\(C,\s*name='(kidmp_[a-z0-9_]+)'\)", f90)
    assert len(bound) >= 8
    for name in bound:
        assert name in declared, name
    # the module keeps the reference's names (I:9, I:28)
    assert re.search(r"(?i)^\s*module\s+mphys_thompson09n\b", f90, re.M) and re.search(r"(?i)subroutine\s+mphys_thompson09_interfacen\b", f90)

    def c_members(struct):
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?(float|int|long|char)\s*", "", decl.replace("\n", " "))
            for n in names.split(","):
                n = n.strip().lstrip("*").strip()
                n = re.sub(r"^(const\s+)?(float|int|long|char)\s*\**\s*", "", n)
                n = re.sub(r"\[.*\]", "", n)
                if n:
                    out.append(n)
        return out

    def f_members(struct):
        body = re.search(r"type, bind\(C\) :: %s(.*?)end type %s" % (struct, struct), f90, re.S).group(1)
        out = []
        for line in body.split("\n"):
            line = line.split("!")[0]
            if "::" in line:
                for n in line.split("::")[1].split(","):
                    n = re.sub(r"\(.*\)", "", n).strip()
                    if n:
                        out.append(n)
        return out

    for struct in ("kidmp_config", "kidmp_kid_columns", "kidmp_wrf_fields", "kidmp_wrf_aerosols"):
        assert [m.lower() for m in f_members(struct)] == [m.lower() for m in c_members(struct)], struct
