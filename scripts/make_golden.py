"""Generates tests/golden/*.npz from the REFERENCE modules (run in the container where /root/reference is mounted).

The fixtures travel with the repo so that the oracle (and, through it, the CUDA path) stays pinned to the reference on
the GPU box, where /root/reference does not exist.
  seg_<arch>.npz   : seeded weights' output of the reference WavLM + Conformer stack (RefSegModel = the reference
                     wav2vec2_model + ConformerEncoder assembled as model_wavlm_conformer.py:238-264)
  emb_resnet.npz   : reference ResNet34(feat_dim=80, embed_dim=256, TSTP) on a seeded fbank + masks
  stats_pool.npz   : the 5 known-answer cases of pyannote-audio/tests/test_stats_pool.py run through the reference StatsPool
  powerset.npz     : reference Powerset(4, 2).mapping is not importable (pyannote.core); the expected mapping of
                     pyannote-audio/tests/utils/test_powerset.py semantics is stored from the documented order instead
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/pyannote-audio/pyannote/audio/"


def load_ref_resnet():
    for n in ["pyannote", "pyannote.audio", "pyannote.audio.models", "pyannote.audio.models.blocks", "pyannote.audio.utils"]:
        if n not in sys.modules:
            m = types.ModuleType(n); m.__path__ = []; sys.modules[n] = m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec); sys.modules[name] = m; spec.loader.exec_module(m)
        return m
    load("pyannote.audio.utils.receptive_field", REF + "utils/receptive_field.py")
    pool = load("pyannote.audio.models.blocks.pooling", REF + "models/blocks/pooling.py")
    rn = load("ref_resnet", REF + "models/embedding/wespeaker/resnet.py")
    return rn, pool


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    for name, N in (("tiny_base", 16000), ("tiny_large", 16000)):
        a = get_arch(name)
        sd = init_state_dict(a, seed=11)
        m = ref_loader.RefSegModel(a).eval()
        m.load_state_dict(sd, strict=False)
        wav = 0.1 * torch.randn(2, N, generator=torch.Generator().manual_seed(5))
        with torch.inference_mode():
            y = m(wav[:, None])
        np.savez_compressed(os.path.join(OUT, f"seg_{name}.npz"), wav=wav.numpy(), logp=y.numpy(), seed=11)
        print(name, y.shape)
    rn, pool = load_ref_resnet()
    net = rn.ResNet34(80, 256, pooling_func="TSTP", two_emb_layer=False).eval()
    sd = init_resnet_state_dict(13, "")
    net.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(3)
    fb = torch.randn(2, 198, 80, generator=g)
    masks = (torch.rand(2, 3, 99, generator=g) > 0.5).float()
    masks[1, 2] = 0
    with torch.inference_mode():
        emb = torch.stack([net(fb.clone(), weights=masks[:, s])[1] for s in range(3)], 1)
    np.savez_compressed(os.path.join(OUT, "emb_resnet.npz"), fbank=fb.numpy(), masks=masks.numpy(), emb=emb.numpy(), seed=13)
    # StatsPool KAT inputs (tests/test_stats_pool.py:28-131) through the reference module
    sp = pool.StatsPool()
    x = torch.Tensor([[[2.0, 4.0], [2.0, 4.0]], [[1.0, 1.0], [1.0, 1.0]]])
    cases = {
        "unweighted": (x, None),
        "one_speaker": (x, torch.Tensor([[0.5, 0.01], [0.2, 0.1]])),
        "multi_speaker": (x, torch.Tensor([[[0.1, 0.2], [0.2, 0.3]], [[0.001, 0.001], [0.2, 0.3]]])),
        "frame_mismatch": (x, torch.Tensor([[[0.2], [0.3]], [[0.001], [0.3]]])),
        "all_zero": (x, torch.Tensor([[0.5, 0.01], [0.0, 0.0]])),
    }
    out = {}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, (xx, w) in cases.items():
            out[k + "_x"] = xx.numpy()
            if w is not None:
                out[k + "_w"] = w.numpy()
            out[k + "_y"] = sp(xx, weights=w).numpy()
    np.savez_compressed(os.path.join(OUT, "stats_pool.npz"), **out)
    make_vbx()
    print("golden fixtures written to", OUT)


def make_vbx():
    """VBx fixture through the reference's diarizen/clustering/VBx.py (vbx_setup + cluster_vbx) on a synthetic PLDA."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vbx_util
    from oracle.pipeline_oracle import filter_embeddings
    spec = importlib.util.spec_from_file_location("ref_vbx", os.path.join(ref_loader.REF, "diarizen/clustering/VBx.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seed, Fa, Fb, iters, lda_dim = 5, 0.07, 0.8, 20, 128
    emb, seg = vbx_util.make_embeddings(seed)
    train, _, _ = filter_embeddings(emb, seg)
    with tempfile.TemporaryDirectory() as d:
        vbx_util.write_plda(d, seed)
        x_tf, plda_tf, psi = ref.vbx_setup(d)
    fea = plda_tf(x_tf(train), lda_dim=lda_dim)
    labels = np.random.default_rng(seed).integers(0, 6, size=len(train))
    from scipy.special import softmax
    q0 = np.zeros((len(labels), labels.max() + 1))
    q0[np.arange(len(labels)), labels] = 1.0
    q0 = softmax(q0 * 7.0, axis=1)
    gamma, pi = ref.cluster_vbx(labels, fea, psi[:lda_dim], Fa=Fa, Fb=Fb, maxIters=iters)
    np.savez_compressed(os.path.join(OUT, "vbx.npz"), seed=seed, Fa=Fa, Fb=Fb, max_iters=iters, train=train, fea=fea,
                        psi=psi, phi=psi[:lda_dim], labels=labels, q0=q0, gamma=gamma, pi=pi)


if __name__ == "__main__":
    main()
