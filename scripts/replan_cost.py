"""How long does a change of engine batch shape cost (re-plan: workspace reallocation + tensor maps)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diarizen_b200.segmentation import SegmentationModel
from diarizen_b200.embedding import EmbeddingModel
from diarizen_b200.archs import init_resnet_state_dict
m = SegmentationModel.random_init("wavlm_large_s80_md", seed=0, precision="fp16")
w = (0.1 * torch.randn(96, 256000)).cuda()
for B in (96, 96, 81, 81, 96, 75, 94):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.hard(w[:B], want_logp=False); torch.cuda.synchronize()
    print(f"seg B={B}: {1e3*(time.perf_counter()-t0):.1f} ms")
e = EmbeddingModel(init_resnet_state_dict(0), precision="fp16")
mk = torch.ones(32, 4, 799).cuda()
for B in (32, 32, 31, 31, 32, 30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e.embed_windows(w[:B], mk[:B]); torch.cuda.synchronize()
    print(f"emb B={B}: {1e3*(time.perf_counter()-t0):.1f} ms")
