"""Times dz_attention (tcgen05 path) alone at the wavlm_large_s80_md pipeline shape (run under gpurun).
usage: attn_bench.py [B] [T] [heads] [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from diarizen_b200 import _lib
from gpu_util import ptr, rup, to_planes

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
T = int(sys.argv[2]) if len(sys.argv) > 2 else 799
h = int(sys.argv[3]) if len(sys.argv) > 3 else 5
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = "cuda"
torch.manual_seed(0)
q = torch.randn(B, T, h, 64, device=dev) * 0.5
k = torch.randn(B, T, h, 64, device=dev)
v = torch.randn(B, T, h, 64, device=dev)
VROW = os.environ.get("DZ_ATTN_VT", "0") != "1"
qkp = to_planes(torch.cat([q.reshape(B * T, h * 64), k.reshape(B * T, h * 64)] + ([v.reshape(B * T, h * 64)] if VROW else []), dim=1))
Tp = rup(T, 8)
vtp = to_planes(v.permute(0, 2, 3, 1).reshape(B, h * 64, T), Tp)
tab = torch.randn(h, 2 * T - 1, device=dev)
gate = 1.0 + torch.rand(B, h, T, device=dev)
out = torch.zeros(2, B * T, h * 64, device=dev, dtype=torch.bfloat16)
a = _lib.AttnArgs()
a.T, a.nheads = T, h
a.q = a.k = ptr(qkp).value
a.qk_plane, a.ldqk, a.q_col, a.k_col = qkp[0].numel(), (3 if VROW else 2) * h * 64, 0, h * 64
a.planes = 1
if VROW:
    a.v, a.v_col = ptr(qkp).value, 2 * h * 64
else:
    a.vt, a.vt_plane, a.ldvt = ptr(vtp).value, vtp[0].numel(), Tp
a.bias_tab, a.gate = ptr(tab).value, ptr(gate).value
a.out, a.out_plane, a.ldo, a.out_planes = ptr(out).value, out[0].numel(), h * 64, 1
L = _lib.lib()
for _ in range(3):
    _lib.check(L.dz_attention(C.byref(a), B, 0, None))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    _lib.check(L.dz_attention(C.byref(a), B, 0, None))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 4.0 * B * h * T * T * 64
print(f"attention B={B} T={T} h={h} vrow={int(VROW)}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TF/s", flush=True)
# correctness spot check against fp64 torch on window 0 / head 0
qv = qkp[0].double()[:T, :64]
kv = qkp[0].double()[:T, h * 64:h * 64 + 64]
vv = vtp[0].double()[0, :64, :T].T
idx = (torch.arange(T, device=dev)[None, :] - torch.arange(T, device=dev)[:, None]) + T - 1
s = qv @ kv.T + gate[0, 0].double()[:, None] * tab[0].double()[idx]
ref = torch.softmax(s, dim=-1) @ vv
print("  max err vs fp64: %.3e" % (out[0][:T, :64].double() - ref).abs().max().item(), flush=True)
