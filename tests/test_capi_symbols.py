"""CPU: the C-ABI library loads and exports every symbol include/parla_b200.h declares, and the
Python binding table covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

from parla_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "parla_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pla_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("pla_stream_pass_f64", "pla_trsv_upper_f64", "pla_lsqr_step_f64", "pla_sjlt_apply_f64",
                 "pla_sketch_gauss_f64", "pla_gemm_f64", "pla_geqrf_f64", "pla_orgqr_f64"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libparla_b200.so is not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.pla_version() >= 100


def test_argument_checks_do_not_need_a_gpu():
    """Bad arguments are rejected before any CUDA call: negative return = -(argument index)."""
    lib = _lib.load()
    rc = lib.pla_trsv_upper_f64(0, 8, 8, 0, 0, 0, 0, 0)
    assert rc == -1 and b"R is null" in lib.pla_last_error()
    rc = lib.pla_gemm_f64(2, 0, 4, 4, 4, 1.0, 0, 4, 0, 4, 0.0, 0, 4, 0, 0, 0)
    assert rc == -1
    rc = lib.pla_stream_pass_f64(1, 10, 20000, 20000, 0, 0, 0, 0, 1.0, 0.0, 0, 1, 0, 0, 0, 0)
    assert rc == -3
