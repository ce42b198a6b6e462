"""CPU-only checks of the C-ABI boundary: the library loads and exports every declared symbol."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "synaptor_b200.h")


@pytest.fixture(scope="module")
def lib():
    from synaptor_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "synaptor_b200", "csrc")])
    return _cabi.lib()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"SYN_API[^;(]*?\b(syn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "syn_chunk_ccs" in syms and "syn_count_overlaps" in syms and len(syms) >= 14


def test_library_exports_every_declared_symbol(lib):
    from synaptor_b200 import _cabi
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature in _cabi.py"
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_no_compute_calls_needed_for_sizes(lib):
    assert lib.syn_version() >= 100
    n = lib.syn_ccs_workspace_bytes(512, 512, 512, 1 << 24, 1 << 19)
    assert n > 512 * 512 * 512 // 8
    assert lib.syn_ccs_workspace_bytes(0, 4, 4, 16, 0) < 0            # empty shape is an error
    assert lib.syn_ccs_workspace_bytes(1 << 20, 1 << 20, 1 << 20, 16, 0) < 0   # beyond 32-bit word indexing
    assert b"" != lib.syn_last_error()


def test_argument_validation_without_gpu(lib):
    # null pointers / bad connectivity are rejected before any CUDA call
    rc = lib.syn_chunk_ccs(None, 4, 4, 4, 0.5, 6, 0, 0, None, None, None, 0, None, None, None, None, 0,
                           None, 0, 16, None, None)
    assert rc == -1
    rc = lib.syn_extract_faces(None, 4, 4, 4, None, None)
    assert rc == -1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import synaptor_b200 as syn
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        syn.connected_components(np.zeros((4, 4, 4), np.float32), 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        syn.proc.tasks.overlap_task(np.ones((2, 2, 2), np.uint32), np.ones((2, 2, 2), np.uint64))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under synaptor_b200/ may reference it."""
    pkg = os.path.join(ROOT, "synaptor_b200")
    for (d, _, files) in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                assert "oracle" not in text.lower(), f"{os.path.join(d, f)} mentions the oracle"


def test_header_is_plain_c(tmp_path):
    """include/synaptor_b200.h is the drop-in boundary: it must compile as C99 (and as C++) on its own."""
    import shutil
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "synaptor_b200.h"\nint main(void) { return syn_version() > 0 ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    if shutil.which("gcc"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                               "-I", inc, str(src)])
    if shutil.which("g++"):
        subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)])
