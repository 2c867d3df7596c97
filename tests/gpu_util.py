"""Helpers for the -m gpu tests: bf16 plane packing and a ctypes GEMM call."""
import ctypes as C

import torch

from diarizen_b200 import _lib


def rup(x, m):
    return (x + m - 1) // m * m


def to_planes(x: torch.Tensor, ld: int = None) -> torch.Tensor:
    """fp32 (..., rows, cols) -> bf16 planes (2, ..., rows, ld), zero padded columns."""
    cols = x.shape[-1]
    ld = ld or rup(cols, 8)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    out = torch.zeros((2,) + tuple(x.shape[:-1]) + (ld,), dtype=torch.bfloat16, device=x.device)
    out[0, ..., :cols] = hi
    out[1, ..., :cols] = lo
    return out


def planes_value(p: torch.Tensor, npass: int) -> torch.Tensor:
    """What the kernels see: hi only (bf16 mode) or hi + lo (bf16x3)."""
    v = p[0].double()
    if npass == 3:
        v = v + p[1].double()
    return v


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run_gemm(d: _lib.GemmDesc, impl: int, force_bn: int = 0):
    _lib.check(_lib.lib().dz_gemm(C.byref(d), impl, force_bn, None))
    torch.cuda.synchronize()


def act_ref(x, act):
    if act == 1:
        return torch.nn.functional.gelu(x)
    if act == 2:
        return x * torch.sigmoid(x)
    if act == 3:
        return torch.relu(x)
    return x
