"""CPU-only: the C-ABI library loads and exports every symbol include/b200pose.h declares, the ctypes
signature table mirrors the header, and size queries (pure host code) work without a GPU."""
import os
import re

import pytest

from rnnpose_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200pose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200pose_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    hdr = open(os.path.join(ROOT, "include", "b200pose.h")).read()
    assert _lib.lib().b200pose_version() == int(re.search(r"#define B200POSE_VERSION (\d+)", hdr).group(1))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 19
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200pose.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_host_side_size_queries():
    L = _lib.lib()
    assert L.b200pose_packed_weights_bytes() > 12_000_000           # 12.3 MB of fp32 weights + padding
    B, h, w = 2, 30, 40
    assert L.b200pose_pyramid_floats(B, h, w) == B * 1200 * (1200 + 300 + 70 + 15)
    assert L.b200pose_refine_workspace_bytes(1, 240, 320) < L.b200pose_refine_workspace_bytes(2, 240, 320)
    # B=32 at 240x320 fills the machine: chained convolutions (1 launch); B=1 runs the layers one by one (11 + flow head 2)
    assert L.b200pose_refine_launch_count(32, 240, 320, 4, 3) == 12 + 4 * (3 + 1 + 1 + 1)
    assert L.b200pose_refine_launch_count(1, 240, 320, 4, 3) == 15 + 4 * (3 + 13 + 1 + 1)
    assert L.b200pose_refine_launch_count(1, 100, 320, 4, 3) == 0
    assert b"workspace" in L.b200pose_error_string(-3)


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rnnpose_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.corr_pyramid(torch.zeros(1, 256, 16, 16), torch.zeros(1, 256, 16, 16))
