"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/vqb.h declares;
compute entry points refuse to run without an sm_100 device (no CPU fallback).  CPU only."""
import ctypes
import os
import re

import pytest
import torch

import vqvae_b200 as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vqb.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vqb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_table_agree():
    assert declared_symbols() == sorted(V._lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(V._lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(V._lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/vqb.h but not exported"
    lib.vqb_version.restype = ctypes.c_int
    assert lib.vqb_version() == 100


def test_binding_argument_counts_match_header():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in V._lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", src)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, params, args)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    prev = (V._lib._BACKEND, V._lib._DEVICE)
    V._lib._BACKEND = V._lib._DEVICE = None
    try:
        with pytest.raises(V._lib.VQBError, match="no CPU fallback"):
            V._lib.lib()
        lib = V._lib.load_library()
        rc = lib.vqb_device_check(0)
        assert rc == -2 and b"CUDA" in lib.vqb_last_error() or rc != 0
        d = V._lib.ConvDesc(1, 8, 4, 4, 3, 1, 1, 0, 0)
        assert lib.vqb_conv1d_fwd(ctypes.byref(d), None, None, None, None, None, None) != 0
    finally:
        V._lib._BACKEND, V._lib._DEVICE = prev


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(V._lib.VQBError, match="missing"):
        V._lib.load_library(str(tmp_path / "nope.so"))
