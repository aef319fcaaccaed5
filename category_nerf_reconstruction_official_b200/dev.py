"""Development hooks (include/cnf_dev.h) -- NOT part of the product: they live in the development build libcnf_sm100_dev.so
(`_lib.build_dev()`, the same sources with -DCNF_DEV).  The hardware self-tests in tests/ use `umma_selftest` to pin the tcgen05
descriptor encodings; tuning scripts call `use_dev_library()` first so that the whole package runs on the instrumented build."""
import os

import torch

from . import _lib
from ._lib import check, ptr, stream


def use_dev_library() -> str:
    """Make `_lib.load()` (and therefore every op of the package) use the development build.  Call before the first op."""
    path = _lib.build_dev()
    assert _lib._lib is None or os.environ.get("CNF_LIB_PATH") == path, "the product library is already loaded in this process"
    os.environ["CNF_LIB_PATH"] = path
    return path


def umma_selftest(mode: int, A: torch.Tensor, B: torch.Tensor, N: int, Kd: int) -> torch.Tensor:
    """D[128,N] = A * B^T through one tcgen05 tile (cnf_umma_selftest)."""
    if not A.is_cuda:
        raise _lib.CnfError("cnf dev ops run only on CUDA (sm_100a) tensors")
    lib = _lib.load_dev()
    D = torch.empty(128, N, device=A.device, dtype=torch.float32)
    check(lib.cnf_umma_selftest(mode, N, Kd, ptr(A.float().contiguous()), ptr(B.float().contiguous()), ptr(D), stream()))
    return D
