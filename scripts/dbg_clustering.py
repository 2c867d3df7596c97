"""GPU debug: for the reference clustering goldens, compare every device step with its CPU counterpart."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from scipy.cluster.hierarchy import linkage, fcluster
from scipy.optimize import linear_sum_assignment
from diarizen_b200 import clustering as cl
from oracle import pipeline_oracle as po

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
z = np.load(os.path.join(G, "glue_clustering.npz"))
for name in [str(n) for n in z["names"]]:
    prm = eval(str(z[f"{name}__params"]), {"__builtins__": {}}, {"dict": dict})
    if "vbx" in prm:
        continue
    emb, seg = z[f"{name}__embeddings"], z[f"{name}__segmentations"].astype(np.float32)
    a = cl.AgglomerativeClustering(); a.threshold, a.min_cluster_size = prm["threshold"], prm["mcs"]
    af, sf, T = cl.frame_statistics(seg)
    train, ci, si = a.filter_embeddings(emb, af, sf, T)
    n = train.shape[0]
    num, lo, hi = a.set_num_clusters(n, prm.get("num"), prm["min"], prm["max"])
    unit = train / np.linalg.norm(train, axis=-1, keepdims=True)
    dd = cl.DeviceDendrogram(unit)
    Z = dd.Z(); Zs = linkage(unit, method="centroid", metric="euclidean")
    zeq = np.array_equal(Z, Zs)
    ms = min(prm["mcs"], max(1, round(0.1 * n)))
    labels, info = dd.cut(prm["threshold"], ms, lo, hi, num)
    ref_tc = po.ahc_cluster(train.copy(), prm["threshold"], prm["mcs"], lo, hi, num)
    got_tc = cl.absorb_small_clusters(unit, labels.astype(np.int64), ms)
    print(f"{name}: n={n} Z bitwise {zeq} (first diff row {np.argwhere((Z != Zs).any(1))[:3].ravel().tolist()}), cut info {info}, train clusters equal {np.array_equal(got_tc, ref_tc)}")
    if not zeq:
        i = int(np.argwhere((Z != Zs).any(1))[0])
        print("   Z dev", Z[max(0, i - 1):i + 2].tolist()); print("   Z ref", Zs[max(0, i - 1):i + 2].tolist())
    hard, soft, cent = a.assign_embeddings(emb, ci, si, ref_tc)
    sc = np.nan_to_num(soft, nan=np.nanmin(soft))
    hh = -2 * np.ones(sc.shape[:2], dtype=np.int8)
    for c, cost in enumerate(sc):
        for s, k in zip(*linear_sum_assignment(cost, maximize=True)):
            hh[c, s] = k
    bad = np.argwhere((hh != hard).any(1)).ravel()
    print(f"   assign vs scipy on the reference clusters: {len(bad)} chunks differ; golden equal {np.array_equal(hh.astype(np.int16), z[f'{name}__hard'])}")
    for c in bad[:3]:
        print("   chunk", c, "device", hard[c].tolist(), "scipy", hh[c].tolist(), "soft", np.round(sc[c], 6).tolist())
