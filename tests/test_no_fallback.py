"""The product path has no CPU fallback: without a CUDA device, or without the built library, it raises."""
import os

import numpy as np
import pytest
import torch

from theta_rrt_b200 import _lib


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_planner_raises_without_cuda():
    from theta_rrt_b200.planner import Planner
    with pytest.raises(_lib.TrrtError):
        Planner(object())  # the device check comes first: nothing is computed on the host


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "libthetarrt.so"))
    with pytest.raises(_lib.TrrtError):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the package may import it."""
    pkg = os.path.dirname(os.path.abspath(_lib.__file__))
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "import oracle" not in src and "from oracle" not in src, name


def test_host_calls_reject_bad_arguments():
    """Argument validation of the C ABI runs before any device work (error codes, no exceptions across the ABI)."""
    import ctypes as C
    lib = _lib.load()
    assert lib.trrt_los_batch(None, 1, 10, 12, None, None, 5, None, None) == 2   # non-square map
    assert lib.trrt_los_batch(None, 1, 40000, 40000, None, None, 5, None, None) == 3  # map too large
    assert lib.trrt_los_batch(None, 1, 10, 10, None, None, 5, None, None) == 1   # null pointers
    assert lib.trrt_rrt_batch(None, None) == 1
    assert lib.trrt_theta_batch(None, None) == 1
    assert lib.trrt_error_string(2).decode().startswith("map must be square")
    assert lib.trrt_nearest_workspace_bytes(1 << 20, 4096) > 0
