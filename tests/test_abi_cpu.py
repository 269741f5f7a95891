"""CPU checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/amod_b200.h, and fails loudly (no fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from amod_b200 import capi, pool


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "amod_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(amod_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(capi.EXPORTS)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.amod_version()


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    # amod_epoch: 6 int32 + int64 + 16 pointers; amod_pool: 6 int64 + 9 pointers
    assert C.sizeof(capi.AmodEpoch) == 6 * 4 + 8 + 16 * 8
    assert C.sizeof(capi.AmodPool) == 6 * 8 + 9 * 8
    assert C.sizeof(capi.AmodStats) == 6 * 8 + 2 * 4 + 2 * 8 + 2 * 4 + 3 * 8 + 2 * 16 * 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    with pytest.raises(pool.AmodError) as ei:
        pool.PoolBuilder(np.zeros((4, 4), np.float32))
    assert ei.value.code == capi.E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_missing_library_raises(tmp_path):
    with pytest.raises(RuntimeError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "amod_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "amod_oracle" not in src, f
