"""A/B of the attention kernels on the pipeline's shapes (96 windows x h heads x 799 frames): CUDA-event time per launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from diarizen_b200 import _lib
from gpu_util import ptr, to_planes
B, T = 96, 799
for h in (1, 5, 10):
    q = torch.randn(B * T, 3 * h * 64, device="cuda")
    qkp = to_planes(q)
    tab = torch.randn(h, 2 * T - 1, device="cuda"); gate = 1.0 + torch.rand(B, h, T, device="cuda")
    out = torch.zeros(2, B * T, h * 64, device="cuda", dtype=torch.bfloat16)
    a = _lib.AttnArgs(); a.T, a.nheads = T, h
    a.q = a.k = a.v = ptr(qkp).value
    a.qk_plane, a.ldqk, a.q_col, a.k_col, a.v_col = qkp[0].numel(), 3 * h * 64, 0, h * 64, 2 * h * 64
    a.planes = 1; a.fp16 = 0
    a.bias_tab, a.gate = ptr(tab).value, ptr(gate).value
    a.out, a.out_plane, a.ldo, a.out_planes = ptr(out).value, out[0].numel(), h * 64, 1
    L = _lib.lib()
    for _ in range(3): _lib.check(L.dz_attention(C.byref(a), B, 0, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): _lib.check(L.dz_attention(C.byref(a), B, 0, None))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"h={h}: {ms*1e3:.1f} us  {4.0*T*T*64*h*B/ms/1e9:.0f} TFLOP/s")
