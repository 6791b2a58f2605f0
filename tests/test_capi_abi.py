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
    assert lib.gqe_abi_version() == _lib.ABI_VERSION == 6


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Plan) == 9 * 4
    assert ctypes.sizeof(_lib.Segment) == 40 + 16      # plan padded to 8-byte alignment
    assert _lib.Segment.query_begin.offset == 40
    # gqe_store_slice: four device pointers, three int64
    assert ctypes.sizeof(_lib.StoreSliceC) == 7 * 8
    assert _lib.StoreSliceC.block_queries.offset == 32 and _lib.StoreSliceC.pool_size.offset == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(_lib.GqeError, match="no CUDA device"):
        _lib.Context(0)


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_scored_accumulator_column_permutation():
    """gqe_pack permutes the output columns of a chain's last matrix so that, in the tcgen05.ld
    16x256b fragment layout (lane t owns accumulator columns 8j + 2(t%4) + {0,1} of repeat j), the
    four columns a lane owns in repeats (2b, 2b+1) are four CONTIGUOUS output columns -- one
    128-bit load of the anchor row they are scored against (csrc/gqe_tc.cuh score_frag)."""
    lib = _lib.load()
    src = [lib.gqe_debug_score_col_src(n) for n in range(256)]
    assert sorted(src) == list(range(256))                                   # a permutation ...
    assert all(s // 16 == n // 16 for n, s in enumerate(src))                # ... inside 16-column blocks
    for b in range(16):
        for m in range(4):                                                   # lane % 4
            owned = [16 * b + 8 * jp + 2 * m + e for jp in (0, 1) for e in (0, 1)]
            assert [src[c] for c in owned] == [16 * b + 4 * m + i for i in range(4)]
