"""One segmentation engine call at the pipeline's batch (96 windows x 16 s, wavlm_large_s80_md, fp16) three times: the command
ncu wraps to capture individual kernels of the segmentation network (-k regex:<kernel> -s <skip> -c <n>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diarizen_b200.segmentation import SegmentationModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
arch = sys.argv[2] if len(sys.argv) > 2 else "wavlm_large_s80_md"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256000
m = SegmentationModel.random_init(arch, seed=0, precision=os.environ.get("DZ_PRECISION", "fp16"))
wav = (0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(0))).cuda()
for _ in range(int(os.environ.get("DZ_REPS", "2"))):
    m.hard(wav, want_logp=False)
torch.cuda.synchronize()
print("ok", m.last_launches)
