"""Micro-benchmark of the tcgen05 GEMM with epilogue variants (run under gpurun)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from diarizen_b200 import _lib
from gpu_util import ptr, rup, to_planes

L = _lib.lib()
dev = "cuda"


def bench(M, N, K, variant, bn=0, iters=20):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
    Ap, Wp = to_planes(A), to_planes(W)
    ldn = rup(N, 8)
    bias = torch.randn(rup(N, 64) + 64, device=dev)
    res = torch.randn(M, ldn, device=dev)
    of = torch.empty(M, ldn, device=dev); ob = torch.empty(2, M, ldn, device=dev, dtype=torch.bfloat16)
    d = _lib.GemmDesc.default()
    d.M, d.N, d.K = M, N, K
    d.a, d.a_plane, d.a_rstride, d.a_kinner, d.a_rows_alloc = ptr(Ap).value, Ap[0].numel(), Ap.shape[-1], K, M
    d.b, d.b_plane, d.ldb, d.b_gstride = ptr(Wp).value, Wp[0].numel(), Wp.shape[-1], Wp[0].numel()
    if "bias" in variant: d.bias = ptr(bias).value
    if "gelu" in variant: d.act = 1
    if "swish" in variant: d.act = 2
    if "res" in variant: d.residual, d.ldr = ptr(res).value, ldn
    if "f32" in variant: d.out_f32, d.ldo = ptr(of).value, ldn
    if "bf" in variant: d.out_bf, d.ob_plane, d.ldob, d.zero_pad_to = ptr(ob).value, ob[0].numel(), ldn, ldn
    p = L.dz_gemm_plan_create(C.byref(d), bn)
    assert p, L.dz_last_error()
    for _ in range(3): _lib.check(L.dz_gemm_plan_launch(p, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.dz_gemm_plan_launch(p, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    L.dz_gemm_plan_destroy(p)
    print(f"M={M} N={N} K={K} bn={bn or 'auto'} {variant:24s} {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    M = 63744
    for var in ("none", "bf", "f32", "bias+swish+bf", "bias+gelu+bf", "bias+res+f32", "bias+res+f32+bf"):
        bench(M, 1024, 256, var)
    for var in ("none", "bf", "bias+res+f32"):
        bench(M, 768, 768, var)
    for bn in (64, 128, 256):
        bench(M, 768, 768, "bf", bn)
    bench(M, 768, 3072, "none"); bench(M, 3072, 768, "none"); bench(8192, 8192, 8192, "none"); bench(8192, 8192, 8192, "bf")
