"""Checkpoint averaging for the recipes' stage-4 inference.

reference: diarizen/ckpt_utils.py:16-60 (average_checkpoints / average_states / load_metric_summary) and
recipes/diar_ssl/infer_avg.py:268-286 (which checkpoints are averaged).  Model-loading-time host code: the averaged
state dict is what `SegmentationModel` uploads; nothing here runs per recording."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Sequence, Union

import torch

Ckpt = Union[str, Path, Dict]


def average_states(states_list: Sequence[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Key-wise mean of state dicts (ckpt_utils.py:32-43).  Like the reference, the sum runs in the dtype of the first
    checkpoint and the final `/ qty` is a true division (integer buffers such as `num_batches_tracked` come out as floats);
    unlike it the inputs are left untouched."""
    if len(states_list) == 0:
        raise ValueError("no checkpoints to average")
    qty = len(states_list)
    avg = {k: v.clone() for k, v in states_list[0].items()}
    for st in states_list[1:]:
        if st.keys() != avg.keys():
            raise KeyError("checkpoints to average have different keys")
        for k in avg:
            avg[k] += st[k].to(avg[k].device)
    return {k: v / qty for k, v in avg.items()}


def _path_of(ckpt: Ckpt) -> Path:
    return Path(ckpt["bin_path"]) if isinstance(ckpt, dict) else Path(ckpt)


def average_checkpoints(checkpoint_list: Sequence[Ckpt]) -> Dict[str, torch.Tensor]:
    """ckpt_utils.py:16-30 without the nn.Module round trip: entries are paths or the `{'bin_path': ...}` records that
    `load_metric_summary` produces."""
    states = [torch.load(str(_path_of(c)), map_location="cpu") for c in checkpoint_list]
    return average_states(states)


def load_metric_summary(metric_file: Union[str, Path], ckpt_path: Union[str, Path]) -> List[Dict]:
    """ckpt_utils.py:45-60: one record per validation line of the trainer's summary file."""
    out = []
    ckpt_path = Path(ckpt_path)
    with open(metric_file, "r") as f:
        for line in f:
            if not line.strip():
                continue
            if "Validation Loss/DER" not in line:
                raise ValueError(f"unexpected line in metric summary: {line!r}")
            tok = line.split()
            epoch = tok[4].split(":")[0]
            out.append({"epoch": int(epoch), "bin_path": ckpt_path / f"epoch_{str(epoch).zfill(4)}/pytorch_model.bin",
                        "Loss": float(tok[-3]), "DER": float(tok[-1])})
    return out


def select_checkpoints(val_metric_lst: List[Dict], val_metric: str = "Loss", val_mode: str = "best", avg_ckpt_num: int = 5) -> List[Dict]:
    """infer_avg.py:270-286: `best` = the n best epochs, `prev` = the best epoch and the n-1 before it, `center` = a window of
    n epochs centred on the best one."""
    ranked = sorted(val_metric_lst, key=lambda r: r[val_metric])
    best = val_metric_lst.index(ranked[0])
    if val_mode == "best":
        sel = ranked[:avg_ckpt_num]
    elif val_mode == "prev":
        sel = val_metric_lst[best - avg_ckpt_num + 1: best + 1]
    elif val_mode == "center":
        sel = val_metric_lst[best - avg_ckpt_num // 2: best + avg_ckpt_num // 2 + 1]
    else:
        raise ValueError(f"unknown val_mode {val_mode!r}")
    if len(sel) != avg_ckpt_num:
        raise AssertionError(f"selected {len(sel)} checkpoints, wanted {avg_ckpt_num}")
    return sel


def write_hub_snapshot(root: Union[str, Path], arch, seg_state_dict: Dict[str, torch.Tensor], emb_state_dict: Dict[str, torch.Tensor],
                       inference_args: Dict, clustering_args: Dict, plda: tuple = None, fusion: Dict = None) -> Path:
    """Saves a model pair in the directory layout `DiariZenPipeline(diarizen_hub=...)` / `from_pretrained(<dir>)` read
    (diarizen/pipelines/inference.py:34-50,95-119): `config.toml`, `pytorch_model.bin` (full segmentation state dict),
    `<arch>.pt` = the WavLM `{config, state_dict}` checkpoint `model.args.wavlm_src` points at
    (model_wavlm_conformer.py:209-221), `wespeaker/pytorch_model.bin` in the lightning `{"state_dict": ...}` wrapping
    (core/model.py:460-473) and, when `plda = (xvec_transform dict, plda dict)` is given, `plda/*.npz` for VBx.  `fusion`
    ({"fusion_dim", "fusion_heads", "fusion_layers"}) marks a multi-channel model (its state dict holds `channel_fusions.*`)."""
    import numpy as np
    from .archs import to_reference_config
    root = Path(root)
    (root / "wespeaker").mkdir(parents=True, exist_ok=True)
    torch.save(dict(seg_state_dict), root / "pytorch_model.bin")
    src = root / f"{arch.name}.pt"
    torch.save({"config": to_reference_config(arch),
                "state_dict": {k[len("wavlm_model."):]: v for k, v in seg_state_dict.items() if k.startswith("wavlm_model.")}}, src)
    torch.save({"state_dict": dict(emb_state_dict)}, root / "wespeaker" / "pytorch_model.bin")
    if plda is not None:
        (root / "plda").mkdir(exist_ok=True)
        np.savez(root / "plda" / "xvec_transform.npz", **plda[0])
        np.savez(root / "plda" / "plda.npz", **plda[1])

    def table(d):
        def val(v):
            if isinstance(v, bool):
                return "true" if v else "false"
            if isinstance(v, str):
                return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
            return repr(v)
        return "".join(f"{k} = {val(v)}\n" for k, v in d.items())
    head = {"attention_in": arch.head_dim_model, "ffn_hidden": arch.head_ffn, "num_head": arch.head_heads, "num_layer": arch.head_layers,
            "kernel_size": arch.head_kernel, "wavlm_layer_num": arch.num_layers + 1, "wavlm_feat_dim": arch.embed_dim}
    if fusion is not None:   # multi-channel model (model_wavlm_conformer_mc.py:26-57): channel_fusion_dim / _layers / _heads
        head.update({"channel_fusion_dim": fusion["fusion_dim"], "channel_fusion_layers": fusion["fusion_layers"],
                     "channel_fusion_heads": fusion["fusion_heads"]})
    with open(root / "config.toml", "w") as f:
        f.write(f"[model.args]\nwavlm_src = \"{src}\"\n{table(head)}\n[inference.args]\n{table(inference_args)}\n[clustering.args]\n{table(clustering_args)}")
    return root
