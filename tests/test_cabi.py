"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol the header declares, and its
compute entry points fail loudly (no silent CPU path) when there is no device."""
import ctypes

import pytest
import torch

from contrastors_b200 import _lib


def test_exports_every_declared_symbol(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.cx_version() >= 100


def test_workspace_size_is_pure_host_math(lib):
    small = lib.cx_infonce_workspace_bytes(128, 128, 64)
    big = lib.cx_infonce_workspace_bytes(2048, 16384, 768)
    assert 0 < small < big
    assert big >= 2048 * 16384 * 2


def test_argument_validation_sets_error(lib):
    rc = lib.cx_gemm_bf16(None, None, None, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1.0, None)
    assert rc != 0
    assert b"null" in lib.cx_last_error()


def test_cpu_tensors_are_rejected_not_emulated():
    import torch.distributed as dist
    from contrastors_b200 import clip_loss
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    import os
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        q = torch.randn(4, 8)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            clip_loss(q, q, lambda x: x)
    finally:
        dist.destroy_process_group()


def test_library_is_not_stale():
    """The in-tree .so (which is what travels to the GPU box) must be newer than every source it is built from."""
    import os
    from contrastors_b200 import build
    lib_m = os.path.getmtime(_lib.LIB_PATH)
    srcs = [os.path.join(build.CSRC, f) for f in os.listdir(build.CSRC)] + [_lib.HEADER]
    stale = [s for s in srcs if os.path.getmtime(s) > lib_m]
    assert not stale, f"rebuild with `python -m contrastors_b200.build`: {stale}"
