"""Synthetic PLDA backend + clustered x-vectors for the VBx tests (no checkpoint is reachable offline)."""
import numpy as np


def make_plda(seed=0, d_in=256, d_out=128):
    r = np.random.default_rng(seed)
    lda, _ = np.linalg.qr(r.standard_normal((d_in, d_out)))
    xt = {"mean1": 0.05 * r.standard_normal(d_in), "mean2": 0.05 * r.standard_normal(d_out), "lda": lda}
    q, _ = np.linalg.qr(r.standard_normal((d_out, d_out)))
    tr = q * np.linspace(0.7, 1.4, d_out)[:, None] + 0.02 * r.standard_normal((d_out, d_out))
    plda = {"mu": 0.1 * r.standard_normal(d_out), "tr": tr, "psi": np.linspace(6.0, 0.05, d_out) ** 2 + 0.01}
    return xt, plda


def write_plda(dirname, seed=0):
    import os
    xt, plda = make_plda(seed)
    os.makedirs(dirname, exist_ok=True)
    np.savez(os.path.join(dirname, "xvec_transform.npz"), **xt)
    np.savez(os.path.join(dirname, "plda.npz"), **plda)
    return xt, plda


def make_embeddings(seed=0, C=60, S=4, T=50, D=256, n_spk=3, noise=0.35):
    """(C,S,D) embeddings drawn around n_spk directions + (C,T,S) hard segmentations with a few overlapped frames."""
    r = np.random.default_rng(seed)
    spk = r.standard_normal((n_spk, D))
    emb = np.full((C, S, D), np.nan)
    seg = np.zeros((C, T, S), dtype=np.float32)
    for c in range(C):
        k = int(r.integers(1, min(S, n_spk) + 1))
        who = r.permutation(n_spk)[:k]
        for s, w in enumerate(who):
            emb[c, s] = spk[w] + noise * r.standard_normal(D)
            a = int(r.integers(0, T // 2))
            seg[c, a:a + int(r.integers(T // 4, T // 2)), s] = 1.0
    return emb, seg
