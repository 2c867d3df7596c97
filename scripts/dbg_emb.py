import sys; sys.path.insert(0,'/root/repo')
import torch
from oracle.emb_oracle import init_resnet_state_dict
from diarizen_b200.embedding import EmbeddingModel
sd=init_resnet_state_dict(0)
m=EmbeddingModel(sd, precision="bf16x3", gemm_impl="simt")
wav=0.1*torch.randn(2,32000); masks=torch.ones(2,4,99)
try:
    out=m.embed_windows(wav,masks); torch.cuda.synchronize(); print("ok", out.shape, out.abs().max())
except Exception as e: print("ERR", e)
