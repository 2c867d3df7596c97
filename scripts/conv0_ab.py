"""A/B of the two conv0 kernels (tcgen05 vs CUDA-core) on the same input: compares the "conv0" tap and the log-probs.
usage: conv0_ab.py <arch> <B> <N>   (run under gpurun; each variant in its own subprocess with a timeout)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 4 and sys.argv[4] == "child":
    import torch
    from diarizen_b200.archs import get_arch, init_state_dict
    from diarizen_b200.segmentation import SegmentationModel
    name, B, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    a = get_arch(name)
    m = SegmentationModel(a, init_state_dict(a, 1), precision="fp16")
    wav = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(1234))
    logp, ml = m.hard(wav.unsqueeze(1))
    torch.cuda.synchronize()
    tap = m.tap("conv0").float().cpu()
    torch.save({"logp": logp.cpu(), "tap": tap}, sys.argv[5])
    print("child ok", name, os.environ.get("DZ_CONV0_TC"), tuple(tap.shape), flush=True)
    sys.exit(0)

import torch
name, B, N = sys.argv[1], sys.argv[2], sys.argv[3]
outs = {}
for simt in ("1", "0"):
    f = f"/tmp/conv0_{simt}.pt"
    env = dict(os.environ, DZ_CONV0_TC="0" if simt == "1" else "1")
    try:
        r = subprocess.run([sys.executable, __file__, name, B, N, "child", f], env=env, timeout=25, capture_output=True, text=True)
        print(r.stdout[-600:], r.stderr[-600:], flush=True)
        if r.returncode == 0:
            outs[simt] = torch.load(f)
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT simt=" + simt, (e.stdout or b"")[-800:], flush=True)
if len(outs) == 2:
    d = (outs["0"]["tap"] - outs["1"]["tap"]).abs()
    print(f"{name}: conv0 tap max |tc - simt| = {d.max().item():.3e} (scale {outs['1']['tap'].abs().max().item():.3e}), "
          f"logp max diff {(outs['0']['logp'] - outs['1']['logp']).abs().max().item():.3e}", flush=True)
