"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/qmprs_b200.h declares; the Python binding table covers the same set."""
import ctypes
import os
import re

from qmprs_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "qmprs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qm_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.qm_version() == 100
    assert lib.qm_prof_num_classes() == 11


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    lib.qm_svd_work_bytes.restype = ctypes.c_longlong
    assert lib.qm_svd_work_bytes(1024, 1024) > 1024 * 2048 * 16
    assert lib.qm_sweep_work_bytes() > 0


def test_sass_uses_fp64_tensor_pipe():
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", build.OUT], capture_output=True, text=True).stdout
    assert out.count("DMMA") > 100          # zgemm + SVD gram/apply kernels
    assert "sm_100a" in out or "SM100" in out.upper() or "EF_CUDA_SM100" in out


def test_sass_has_no_floating_point_atomics():
    """Run-to-run reproducibility: no kernel may accumulate with floating-point
    atomics -- integer tickets / flags (order independent) are the only RED/ATOM instructions allowed."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", build.OUT], capture_output=True, text=True).stdout
    bad = [ln.strip() for ln in out.splitlines()
           if re.search(r"\b(RED|REDG|ATOM|ATOMG|ATOMS)\b", ln) and re.search(r"\.(F64|F32|F16)", ln)]
    assert not bad, bad[:5]
