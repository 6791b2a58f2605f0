"""The C-ABI library loads and exports exactly what include/gqe.h declares."""
import ctypes
import os
import re

import pytest
import torch

from graphqembed_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gqe.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gqe_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "libgqe_b200.so does not export %s" % n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.gqe_abi_version() == 3


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Plan) == 9 * 4
    assert ctypes.sizeof(_lib.Segment) == 40 + 16      # plan padded to 8-byte alignment
    assert _lib.Segment.query_begin.offset == 40


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(_lib.GqeError, match="no CUDA device"):
        _lib.Context(0)


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
