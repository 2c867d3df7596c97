"""Runs the embedding engine alone on 32 windows of 16 s (the pipeline's engine batch) three times and prints the per-step
profile; under ncu it exposes the ResNet conv GEMM launches (36 tensor-core GEMMs per forward)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diarizen_b200.archs import init_resnet_state_dict
from diarizen_b200.embedding import EmbeddingModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = EmbeddingModel(init_resnet_state_dict(0), precision="fp16")
wav = 0.1 * torch.randn(B, 256000, generator=torch.Generator().manual_seed(0)).cuda()
masks = torch.ones(B, 4, 799).cuda()
for _ in range(3):
    out = m.embed_windows(wav, masks)
torch.cuda.synchronize()
if os.environ.get("DZ_PROFILE", "1") == "1":
    ep = m.profile(); ep = m.profile()
    print("EMB per-batch(%d) total %.2f ms" % (B, sum(p[1] for p in ep)))
    for n, ms, fl in ep:
        print("  %-16s %7.3f ms  %6.1f TF/s" % (n, ms, fl / ms / 1e9 if fl else 0))
