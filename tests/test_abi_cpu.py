"""CPU-side checks of the drop-in boundary: libalphagpu.so loads, exports every symbol include/alphagpu.h
declares, reports the plugin constants, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from alphagpu_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_match_header(lib):
    from alphagpu_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "alphagpu.h")).read()
    declared = set(re.findall(r"\b(agpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in alphagpu.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.agpu_abi_version() == 1


def test_struct_sizes_match_header():
    from alphagpu_b200 import _lib
    assert C.sizeof(_lib.Config) == 40
    assert C.sizeof(_lib.GameInfo) == 20
    assert C.sizeof(_lib.Samples) == 72
    assert C.sizeof(_lib.RunStats) == 64
    assert C.sizeof(_lib.TreeDump) == 88
    assert C.sizeof(_lib.KernelTimes) == 8 * 8 * 2 + 16


def test_game_info_matches_reference_constants(lib):
    import alphagpu_b200 as ag
    want = {("connect4", 0, 0): (7, 42, 42, 42, 104), ("gobang", 3, 3): (9, 9, 9, 9, 104), ("gobang", 9, 5): (81, 81, 81, 81, 104),
            ("hex", 7, 0): (49, 64, 64, 49, 104), ("reversi8", 0, 0): (65, 64, 64, 70, 152), ("reversi6", 0, 0): (37, 36, 36, 50, 152)}
    for (name, n, nv), w in want.items():
        s = ag.GameSpec.named(name, n, nv)
        assert (s.maxActions, s.VectorizedState, s.FeatureSize, s.maxLengthGame, s.position_dtype.itemsize) == w
    with pytest.raises(ValueError):
        ag.GameSpec.named("gobang", 15, 5).maxActions      # 225 > 192 bits (Bitboard.jl:22)


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import alphagpu_b200 as ag
    with pytest.raises(ag._lib.AlphaGPUError) as e:
        ag.Context(ag.GameSpec.named("connect4"), 8, 4, 128, 6)
    assert e.value.code in (ag._lib.ERR_NO_DEVICE, ag._lib.ERR_CUDA)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under alphagpu_b200/ may reference it."""
    pkg = os.path.join(ROOT, "alphagpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
                assert "oracle/" not in txt and "liboracle" not in txt, f


def test_poolsample_ring():
    import alphagpu_b200 as ag
    spec = ag.GameSpec.named("connect4")
    buf = ag.PoolSample(spec, 10)
    mk = lambda n, v: (np.full((n, 84), v, np.int8), np.full((n, 7), v, np.float32), np.full(n, 1, np.int8), np.full(n, v, np.float32), np.full((n, 42), v, np.int8))
    buf.push_block(*mk(6, 1))
    assert buf.length_buffer() == 6 and not buf.full
    buf.push_block(*mk(6, 2))
    assert buf.full and buf.length_buffer() == 10 and buf.currentIndex == 2
    assert buf.value[0] == 2 and buf.value[1] == 2 and buf.value[2] == 1 and buf.value[9] == 2
