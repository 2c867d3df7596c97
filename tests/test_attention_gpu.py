"""Attention kernels (tcgen05 and CUDA-core) vs a float64 torch softmax-attention on the same bf16 operands.
reference semantics: components.py:455-480 with the gated bias of :690-725."""
import ctypes as C

import pytest
import torch

from diarizen_b200 import _lib
from gpu_util import ptr, rup, to_planes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vrow", [0, 1], ids=["vT", "vrow"])
@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("B,T,h,bias", [(2, 49, 1, True), (2, 249, 3, True), (1, 799, 2, True), (2, 130, 4, False), (1, 64, 1, True),
                                        (3, 799, 5, True)])
def test_attention(impl, B, T, h, bias, vrow):
    torch.manual_seed(T + h)
    dev = "cuda"
    q = torch.randn(B, T, h, 64, device=dev) * 0.5
    k = torch.randn(B, T, h, 64, device=dev)
    v = torch.randn(B, T, h, 64, device=dev)
    cols = [q.reshape(B * T, h * 64), k.reshape(B * T, h * 64)] + ([v.reshape(B * T, h * 64)] if vrow else [])
    qkp = to_planes(torch.cat(cols, dim=1))
    Tp = rup(T, 8)
    vtp = to_planes(v.permute(0, 2, 3, 1).reshape(B, h * 64, T), Tp)
    planes = 2 if impl == 1 else 1
    tab = torch.randn(h, 2 * T - 1, device=dev) if bias else None
    gate = (1.0 + torch.rand(B, h, T, device=dev)) if bias else None
    out = torch.zeros(2, B * T, h * 64, device=dev, dtype=torch.bfloat16)
    a = _lib.AttnArgs()
    a.T, a.nheads = T, h
    a.q = a.k = ptr(qkp).value
    a.qk_plane, a.ldqk, a.q_col, a.k_col = qkp[0].numel(), (3 if vrow else 2) * h * 64, 0, h * 64
    a.planes = planes
    if vrow:   # V row-major next to q | k (what the engine's single QKV projection writes)
        a.v, a.v_col = ptr(qkp).value, 2 * h * 64
    else:
        a.vt, a.vt_plane, a.ldvt = ptr(vtp).value, vtp[0].numel(), Tp
    a.bias_tab = ptr(tab).value if bias else None
    a.gate = ptr(gate).value if bias else None
    a.out, a.out_plane, a.ldo, a.out_planes = ptr(out).value, out[0].numel(), h * 64, 2
    _lib.check(_lib.lib().dz_attention(C.byref(a), B, impl, None))
    torch.cuda.synchronize()

    def val(p):
        x = p[0].double()
        return x + p[1].double() if planes == 2 else x
    qv = val(qkp)[:, :h * 64].view(B, T, h, 64).permute(0, 2, 1, 3)
    kv = val(qkp)[:, h * 64:2 * h * 64].view(B, T, h, 64).permute(0, 2, 1, 3)
    vv = val(vtp)[..., :T].view(B, h, 64, T).permute(0, 1, 3, 2)
    s = qv @ kv.transpose(-1, -2)
    if bias:
        idx = (torch.arange(T, device=dev)[None, :] - torch.arange(T, device=dev)[:, None]) + T - 1
        s = s + gate.double()[..., None] * tab.double()[:, idx][None]
    ref = (torch.softmax(s, dim=-1) @ vv).permute(0, 2, 1, 3).reshape(B * T, h * 64)
    got = out[0].double() + out[1].double()
    err = (got - ref).abs().max().item()
    # tc path rounds P to bf16 (rel 2^-9) before P.V; simt path is fp32 throughout
    tol = 1.5e-2 if impl == 0 else 2e-4
    assert err < tol, f"max err {err:.3e}"


@pytest.mark.parametrize("B,T,h,bias", [(2, 49, 1, True), (2, 249, 3, True), (1, 799, 2, True), (2, 130, 4, False), (1, 64, 1, True),
                                        (3, 799, 5, True)])
def test_attention_split_precision_tensor_core(B, T, h, bias):
    """attention_tc3_kernel (bf16 hi + lo operand planes, three tensor-core passes per product: the fp32-class mode) vs float64
    torch on the hi + lo operand values; fp32-class tolerance."""
    torch.manual_seed(T + h)
    dev = "cuda"
    q = torch.randn(B, T, h, 64, device=dev) * 0.5
    k = torch.randn(B, T, h, 64, device=dev)
    v = torch.randn(B, T, h, 64, device=dev)
    qkp = to_planes(torch.cat([q.reshape(B * T, h * 64), k.reshape(B * T, h * 64), v.reshape(B * T, h * 64)], dim=1))
    tab = torch.randn(h, 2 * T - 1, device=dev) if bias else None
    gate = (1.0 + torch.rand(B, h, T, device=dev)) if bias else None
    out = torch.zeros(2, B * T, h * 64, device=dev, dtype=torch.bfloat16)
    a = _lib.AttnArgs()
    a.T, a.nheads = T, h
    a.q = a.k = a.v = ptr(qkp).value
    a.qk_plane, a.ldqk, a.q_col, a.k_col, a.v_col = qkp[0].numel(), 3 * h * 64, 0, h * 64, 2 * h * 64
    a.planes = 2
    a.bias_tab = ptr(tab).value if bias else None
    a.gate = ptr(gate).value if bias else None
    a.out, a.out_plane, a.ldo, a.out_planes = ptr(out).value, out[0].numel(), h * 64, 2
    _lib.check(_lib.lib().dz_attention(C.byref(a), B, 0, None))
    torch.cuda.synchronize()
    x = qkp[0].double() + qkp[1].double()
    qv = x[:, :h * 64].view(B, T, h, 64).permute(0, 2, 1, 3)
    kv = x[:, h * 64:2 * h * 64].view(B, T, h, 64).permute(0, 2, 1, 3)
    vv = x[:, 2 * h * 64:].view(B, T, h, 64).permute(0, 2, 1, 3)
    s = qv @ kv.transpose(-1, -2)
    if bias:
        idx = (torch.arange(T, device=dev)[None, :] - torch.arange(T, device=dev)[:, None]) + T - 1
        s = s + gate.double()[..., None] * tab.double()[:, idx][None]
    ref = (torch.softmax(s, dim=-1) @ vv).permute(0, 2, 1, 3).reshape(B * T, h * 64)
    got = out[0].double() + out[1].double()
    err = (got - ref).abs().max().item()
    assert err < 3e-4, f"max err {err:.3e}"      # bf16 hi+lo carries ~16 bits; dropped lo*lo terms and ex2.approx bound the error
