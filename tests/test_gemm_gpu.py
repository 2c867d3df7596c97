"""Parity of the tcgen05 GEMM (and its CUDA-core twin) against a float64 torch contraction of the same
bf16-rounded operands.  Tolerances: bf16 mode 2e-3 relative to the output scale (fp32 accumulation of exact
bf16 products: only summation order differs), bf16x3 mode 2e-5."""
import pytest
import torch

from diarizen_b200 import _lib
from gpu_util import act_ref, planes_value, ptr, run_gemm, rup, to_planes

pytestmark = pytest.mark.gpu


def _tol(npass):
    return 5e-4 if npass == 1 else 8e-5


def _check(name, got, ref, npass):
    scale = ref.abs().max().item() + 1e-6
    err = (got.double() - ref).abs().max().item() / scale
    assert err < _tol(npass), f"{name}: rel err {err:.3e}"


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("M,N,K,bn", [(300, 200, 136, 0), (128, 64, 64, 64), (257, 666, 768, 128), (513, 1092, 1024, 256),
                                     (96, 11, 256, 0), (1000, 384, 53, 0)])
def test_linear_epilogue(impl, npass, M, N, K, bn):
    torch.manual_seed(M + N + K)
    dev = "cuda"
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(rup(N, 64) + 64, device=dev)
    res = torch.randn(M, rup(N, 8), device=dev)
    Ap, Wp = to_planes(A), to_planes(W)
    ldn = rup(N, 8)
    out_f = torch.full((M, ldn), 7.0, device=dev)
    out_b = torch.full((2, M, ldn), 7.0, device=dev, dtype=torch.bfloat16)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass = M, N, K, npass
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_rows_alloc = ptr(Ap).value, Ap[0].numel(), Ap.shape[-1], K, M
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(Wp).value, Wp[0].numel(), Wp.shape[-1], Wp[0].numel()
    d.bias, d.act, d.alpha = ptr(bias).value, 1, 0.5
    d.residual, d.ldr = ptr(res).value, ldn
    d.out_f32, d.ldo = ptr(out_f).value, ldn
    d.out_bf, d.ob_plane, d.ldob, d.out_planes, d.zero_pad_to = ptr(out_b).value, out_b[0].numel(), ldn, 2, ldn
    run_gemm(d, impl, bn)
    ref = 0.5 * act_ref(planes_value(Ap, npass)[:, :K] @ planes_value(Wp, npass)[:, :K].T + bias[:N].double(), 1) + res[:, :N].double()
    _check("f32", out_f[:, :N], ref, npass)
    got_b = out_b[0].double() + out_b[1].double()
    _check("bf planes", got_b[:, :N], ref, 3)
    assert (out_b[:, :, N:ldn] == 0).all(), "pad columns must be zeroed"
    assert (out_f[:, N:] == 7.0).all(), "fp32 output must not touch pad columns"


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("B,Tin,Cin,Cout,k", [(3, 401, 24, 40, 3), (2, 1000, 153, 224, 3), (2, 300, 90, 161, 2)])
def test_conv1d_as_strided_gemm(impl, npass, B, Tin, Cin, Cout, k):
    """conv1d(k, stride 2) over channels-last input == GEMM over an overlapping-row view (components.py:119)."""
    torch.manual_seed(Tin)
    dev = "cuda"
    x = torch.randn(B, Tin, Cin, device=dev)
    w = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
    Cp = rup(Cin, 8)
    xp = to_planes(x, Cp)                                    # (2, B, Tin, Cp)
    wr = torch.zeros(Cout, k, Cp, device=dev)
    wr[:, :, :Cin] = w.permute(0, 2, 1)
    wp = to_planes(wr.reshape(Cout, k * Cp))
    Tout = (Tin - k) // 2 + 1
    ldo = rup(Cout, 8)
    out = torch.zeros(B, Tout, ldo, device=dev)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass, d.batches = Tout, Cout, k * Cp, npass, B
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_bstride, d.a_rows_alloc = ptr(xp).value, xp[0].numel(), 2 * Cp, k * Cp, Tin * Cp, Tout
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(wp).value, wp[0].numel(), wp.shape[-1], wp[0].numel()
    d.act = 1
    d.out_f32, d.ldo, d.of_bstride = ptr(out).value, ldo, Tout * ldo
    run_gemm(d, impl)
    xv = planes_value(xp, npass)[..., :Cin].permute(0, 2, 1)
    wv = planes_value(wp, npass).reshape(Cout, k, Cp)[:, :, :Cin].permute(0, 2, 1)
    ref = torch.nn.functional.gelu(torch.nn.functional.conv1d(xv, wv, stride=2)).permute(0, 2, 1)
    _check("conv", out[..., :Cout], ref, npass)


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("B,T,D", [(2, 49, 128), (2, 249, 768), (1, 300, 1024)])
def test_grouped_posconv(impl, npass, B, T, D):
    """Grouped conv1d(k=128, pad=64, groups=16) + bias + GELU + residual (components.py:366-380, :981)."""
    torch.manual_seed(T)
    dev = "cuda"
    G, KT = 16, 128
    Dg = D // G
    x = torch.randn(B, T, D, device=dev)
    w = torch.randn(D, Dg, KT, device=dev) / (Dg * KT) ** 0.5
    bias = torch.randn(D + 64, device=dev)
    stage = torch.zeros(B, T + 128, G, 64, device=dev)
    stage[:, 64:64 + T, :, :Dg] = x.view(B, T, G, Dg)
    sp = to_planes(stage.view(B, T + 128, G * 64))
    wr = torch.zeros(G, Dg, KT, 64, device=dev)
    wr[:, :, :, :Dg] = w.view(G, Dg, Dg, KT).permute(0, 1, 3, 2)
    wp = to_planes(wr.view(G * Dg, KT * 64))
    res = x.clone().view(B * T, D).contiguous()
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass, d.batches, d.groups = T, Dg, KT * 64, npass, B, G
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_kouter, d.a_gstride = ptr(sp).value, sp[0].numel(), G * 64, 64, G * 64, 64
    d.a_bstride, d.a_rows_alloc = (T + 128) * G * 64, T
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(wp).value, wp[0].numel(), KT * 64, Dg * KT * 64
    d.bias, d.act, d.group_cols = ptr(bias).value, 1, Dg
    d.residual, d.res_bstride, d.ldr = ptr(res).value, T * D, D
    d.out_f32, d.of_bstride, d.ldo = ptr(res).value, T * D, D
    run_gemm(d, impl)
    xv = planes_value(sp, npass).view(B, T + 128, G, 64)[:, 64:64 + T, :, :Dg].reshape(B, T, D)
    wv = planes_value(wp, npass).view(G, Dg, KT, 64)[..., :Dg].permute(0, 1, 3, 2).reshape(D, Dg, KT)
    pc = torch.nn.functional.conv1d(xv.permute(0, 2, 1), wv, bias[:D].double(), padding=64, groups=G)[..., :-1]
    ref = x.double() + torch.nn.functional.gelu(pc).permute(0, 2, 1)
    _check("posconv", res.view(B, T, D), ref, npass)


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
def test_transposed_output(impl):
    """q|k row-major + v^T planes from one projection GEMM."""
    torch.manual_seed(5)
    dev = "cuda"
    B, T, D, h = 2, 99, 256, 3
    M, N = B * T, 3 * h * 64
    A = torch.randn(M, D, device=dev)
    W = torch.randn(N, D, device=dev) / D ** 0.5
    Ap, Wp = to_planes(A), to_planes(W)
    Tp = rup(T, 8)
    qk = torch.zeros(2, M, 2 * h * 64, device=dev, dtype=torch.bfloat16)
    vt = torch.zeros(2, B, h * 64, Tp, device=dev, dtype=torch.bfloat16)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass = M, N, D, 3
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_rows_alloc = ptr(Ap).value, Ap[0].numel(), D, D, M
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(Wp).value, Wp[0].numel(), D, Wp[0].numel()
    d.out_bf, d.ob_plane, d.ldob, d.out_planes = ptr(qk).value, qk[0].numel(), 2 * h * 64, 2
    d.out_t, d.ot_plane, d.ot_bstride, d.ldt, d.tr_col0, d.seq_len = ptr(vt).value, vt[0].numel(), h * 64 * Tp, Tp, 2 * h * 64, T
    run_gemm(d, impl)
    ref = planes_value(Ap, 3) @ planes_value(Wp, 3).T
    _check("qk", qk[0].double() + qk[1].double(), ref[:, :2 * h * 64], 3)
    vref = ref[:, 2 * h * 64:].view(B, T, h * 64).permute(0, 2, 1)
    _check("vt", (vt[0].double() + vt[1].double())[..., :T], vref, 3)
    assert (vt[..., T:] == 0).all()


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("B,H,W,Cin,Cout,ks,stride,res", [(2, 10, 50, 32, 32, 3, 1, True), (1, 20, 199, 32, 64, 3, 2, False),
                                                         (2, 8, 130, 64, 128, 1, 2, False), (1, 6, 300, 128, 128, 3, 1, True),
                                                         (1, 5, 77, 256, 256, 3, 1, True)])
def test_conv2d_as_gemm(impl, npass, B, H, W, Cin, Cout, ks, stride, res):
    """3x3 / 1x1 conv2d (+ folded BN bias, residual from 16-bit planes, ReLU after the add) over a zero-bordered NHWC
    image == GEMM whose k loop walks the input rows of the window (resnet.py:139-144)."""
    torch.manual_seed(W)
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    bias = torch.randn(rup(Cout, 64) + 64, device=dev)
    pad = 1 if ks == 3 else 0
    Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
    Wp, Wop = W + 2, Wo + 2
    xin = torch.zeros(B, H, Wp, Cin, device=dev)
    xin[:, :, 1:W + 1] = x.permute(0, 2, 3, 1)
    xp = to_planes(xin.reshape(B * H * Wp, Cin))                       # (2, B*H*Wp, Cin)
    xp = torch.cat([xp, torch.zeros(2, 64, Cin, device=dev, dtype=xp.dtype)], dim=1).contiguous()   # slack for tail boxes
    run_len = ks * Cin
    krun = rup(run_len, 64)
    wr = torch.zeros(Cout, ks, krun, device=dev)
    wr[:, :, :run_len] = w.permute(0, 2, 3, 1).reshape(Cout, ks, ks * Cin)
    wp = to_planes(wr.reshape(Cout, ks * krun))
    resid = torch.randn(B, Ho, Wop, Cout, device=dev)
    resid[:, :, 0] = 0; resid[:, :, -1] = 0
    rp = to_planes(resid.reshape(B * Ho * Wop, Cout))
    out = torch.zeros(2, B * Ho * Wop, Cout, device=dev, dtype=torch.bfloat16)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass, d.batches = Wo, Cout, ks * krun, npass, B * Ho
    d.a, d.a_plane, d.a_rstride, d.a_bstride, d.a_hstride = ptr(xp).value, xp[0].numel(), stride * Cin, H * Wp * Cin, Wp * Cin
    d.a_kinner = d.K
    d.conv_runs, d.conv_run_len, d.conv_x0, d.conv_h0, d.conv_hs, d.conv_Ho, d.conv_H = ks, run_len, (0 if ks == 3 else Cin), -pad, stride, Ho, H
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(wp).value, wp[0].numel(), ks * krun, wp[0].numel()
    d.bias, d.act, d.act_after_res = ptr(bias).value, 3, 1
    if res:
        d.res16, d.res16_plane, d.res16_bstride, d.ldr16, d.res16_row_off = ptr(rp).value, rp[0].numel(), Wop * Cout, Cout, 1
    d.out_bf, d.ob_plane, d.ob_bstride, d.ldob, d.out_row_off, d.out_planes = ptr(out).value, out[0].numel(), Wop * Cout, Cout, 1, 2
    run_gemm(d, impl)
    xv = planes_value(xp, npass)[:B * H * Wp].view(B, H, Wp, Cin)[:, :, 1:W + 1].permute(0, 3, 1, 2)
    wv = planes_value(wp, npass).view(Cout, ks, krun)[:, :, :run_len].reshape(Cout, ks, ks, Cin).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(xv, wv, bias[:Cout].double(), stride=stride, padding=pad)
    if res:
        ref = ref + planes_value(rp, 3).view(B, Ho, Wop, Cout)[:, :, 1:Wo + 1].permute(0, 3, 1, 2)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    got = (out[0].double() + out[1].double()).view(B, Ho, Wop, Cout)
    _check("conv2d", got[:, :, 1:Wo + 1], ref, npass if npass == 1 else 3)
    assert (got[:, :, 0] == 0).all() and (got[:, :, -1] == 0).all(), "zero border must stay untouched"


@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("B,Tin,Cin,Cout", [(2, 1000, 153, 224), (2, 777, 512, 153), (3, 401, 24, 40), (1, 300, 224, 255)])
def test_conv1d_with_fused_layernorm_gelu(npass, B, Tin, Cin, Cout):
    """conv1d -> LayerNorm(channels) -> GELU in ONE launch (row LayerNorm on the accumulator row in TMEM): the conv stack of
    the layer-norm feature extractor (components.py:119-122)."""
    torch.manual_seed(Tin + Cout)
    dev, k = "cuda", 3
    x = torch.randn(B, Tin, Cin, device=dev)
    w = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
    gamma = torch.zeros(rup(Cout, 32), device=dev)      # zero padded to a multiple of 32 columns (contract of ln_gamma / ln_beta)
    beta = torch.zeros(rup(Cout, 32), device=dev)
    gamma[:Cout] = 1.0 + 0.2 * torch.randn(Cout, device=dev)
    beta[:Cout] = 0.3 * torch.randn(Cout, device=dev)
    Cp = rup(Cin, 8)
    xp = to_planes(x, Cp)
    wr = torch.zeros(Cout, k, Cp, device=dev)
    wr[:, :, :Cin] = w.permute(0, 2, 1)
    wp = to_planes(wr.reshape(Cout, k * Cp))
    Tout = (Tin - k) // 2 + 1
    ldo = rup(Cout, 8)
    out_b = torch.full((2, B, Tout, ldo), 7.0, device=dev, dtype=torch.bfloat16)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K, d.npass, d.batches = Tout, Cout, k * Cp, npass, B
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_bstride, d.a_rows_alloc = ptr(xp).value, xp[0].numel(), 2 * Cp, k * Cp, Tin * Cp, Tout
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(wp).value, wp[0].numel(), wp.shape[-1], wp[0].numel()
    d.act = 1
    d.ln_gamma, d.ln_beta, d.ln_eps = ptr(gamma).value, ptr(beta).value, 1e-5
    d.out_bf, d.ob_plane, d.ldob, d.ob_bstride, d.out_planes, d.zero_pad_to = ptr(out_b).value, out_b[0].numel(), ldo, Tout * ldo, 2, ldo
    run_gemm(d, 0)
    xv = planes_value(xp, npass)[..., :Cin].permute(0, 2, 1)
    wv = planes_value(wp, npass).reshape(Cout, k, Cp)[:, :, :Cin].permute(0, 2, 1)
    y = torch.nn.functional.conv1d(xv, wv, stride=2).permute(0, 2, 1)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(y, (Cout,), gamma[:Cout].double(), beta[:Cout].double(), 1e-5))
    got = out_b[0].double() + out_b[1].double()
    _check("ln+gelu planes", got[..., :Cout], ref, 3 if npass == 3 else 1)
    assert (out_b[..., Cout:ldo] == 0).all(), "pad columns must be zeroed"
