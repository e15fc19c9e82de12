"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rnacode_cuda.h declares,
and refuses to create a context without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from rnacode_b200 import build as rbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    rbuild.build()
    return C.CDLL(rbuild.LIB)


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rnacode_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rc_[a-z_]+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("rc_create", "rc_score_aln", "rc_score_samples", "rc_batch_create", "rc_batch_run"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), "libRNAcode_cuda.so does not export " + s


def test_binding_lists_same_symbols():
    from rnacode_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    lib.rc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    assert lib.rc_create(C.byref(h), 0) == -2  # RC_ERR_CUDA
    assert not h.value


def test_version_string(lib):
    lib.rc_version.restype = C.c_char_p
    assert b"sm_100a" in lib.rc_version()


def test_product_does_not_reference_oracle():
    """The product (package + csrc + include) must not import / include / link anything under oracle/."""
    bad = []
    for base in ("rnacode_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_py|liboracle|rnacode_oracle|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
