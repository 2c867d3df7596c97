"""Batch command line over a wav.scp.

Two modes, one flag set each, same names and defaults as the reference scripts:
  * hub mode      - diarizen/pipelines/inference.py:194-368: `--diarizen_hub DIR --embedding_model CKPT --in_wav_scp SCP`
  * experiment mode - recipes/diar_ssl/infer_avg.py:100-345: `-C exp/config.toml` plus either `--segmentation_model CKPT` or
    `--val_metric_summary FILE` (the `--avg_ckpt_num` checkpoints picked by `--val_metric` / `--val_mode` are averaged),
    RTTMs written to `--out_dir`.
Host-side glue only: every recording goes through `DiariZenPipeline.__call__`."""
from __future__ import annotations

import argparse
import os
from pathlib import Path
from typing import Dict, List, Optional


def scp2path(scp_file) -> List[str]:
    """diarizen/pipelines/utils.py:4-7"""
    return [line.strip().split()[1] for line in open(scp_file) if line.strip()]


def load_scp(scp_file) -> Dict[str, str]:
    """`<session> <path>` per line (recipes' wav.scp)"""
    out = {}
    for line in open(scp_file):
        if line.strip():
            k, v = line.strip().split(None, 1)
            out[k] = v.strip()
    return out


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser("diarizen_b200: diarize every recording of a wav.scp and write one RTTM per session", add_help=True,
                                usage="%(prog)s [options]")
    p.add_argument("-i", "--in_wav_scp", type=str, required=True, help="list of recordings, one `<session> <path>` per line", dest="in_wav_scp")
    p.add_argument("--diarizen_hub", type=str, default=None, help="model directory (config.toml, pytorch_model.bin, plda/)")
    p.add_argument("--embedding_model", type=str, required=True, help="WeSpeaker ResNet34 checkpoint")
    # experiment mode (infer_avg.py)
    p.add_argument("-C", "--configuration", type=str, default=None, help="config.toml of a training run (switches to experiment mode)")
    p.add_argument("-o", "--out_dir", type=str, default=None, help="where the RTTMs go in experiment mode")
    p.add_argument("--avg_ckpt_num", type=int, default=5, help="how many checkpoints are averaged")
    p.add_argument("--val_metric", type=str, default="Loss", choices=["Loss", "DER"], help="column of the metric summary used to rank epochs")
    p.add_argument("--val_mode", type=str, default="best", choices=["best", "prev", "center"], help="best: the n best epochs; prev: the best and the n-1 before it; center: a window around the best")
    p.add_argument("--val_metric_summary", type=str, default="", help="per-epoch validation summary written by the trainer")
    p.add_argument("--segmentation_model", type=str, default="", help="single segmentation checkpoint (experiment mode without averaging)")
    # inference parameters
    p.add_argument("--seg_duration", type=int, default=16, help="window length in seconds")
    p.add_argument("--segmentation_step", type=float, default=0.1, help="window hop as a fraction of the window length")
    p.add_argument("--batch_size", type=int, default=32, help="windows per batch in the configuration (results do not depend on it)")
    p.add_argument("--apply_median_filtering", action=argparse.BooleanOptionalAction, default=True,
                   help="11-frame median filter on the window decisions")
    # clustering parameters
    p.add_argument("--clustering_method", type=str, default="VBxClustering", choices=["VBxClustering", "AgglomerativeClustering"],
                   help="global clustering of the window-level speakers")
    p.add_argument("--min_speakers", type=int, default=1, help="lower bound on the number of speakers")
    p.add_argument("--max_speakers", type=int, default=20, help="upper bound on the number of speakers")
    p.add_argument("--ahc_criterion", type=str, default="distance", help="flat-cluster criterion of the AHC initialisation (VBx)")
    p.add_argument("--ahc_threshold", type=float, default=0.6, help="linkage threshold (distance) or cluster count (maxclust)")
    p.add_argument("--min_cluster_size", type=int, default=13, help="smaller AHC clusters are merged into their nearest large one")
    p.add_argument("--Fa", type=float, default=0.07, help="VBx: scale of the sufficient statistics")
    p.add_argument("--Fb", type=float, default=0.8, help="VBx: speaker regularisation")
    p.add_argument("--lda_dim", type=int, default=128, help="VBx: dimensions kept after the PLDA transform")
    p.add_argument("--max_iters", type=int, default=20, help="VBx: iteration cap")
    p.add_argument("--rttm_out_dir", type=str, default=None, required=False, help="where the RTTMs go in hub mode")
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "bf16", "bf16x3"], help="operand precision mode")
    return p


def config_from_args(args) -> Dict:
    """inference.py:323-354: the `config_parse` override built from the command line."""
    inference_config = {"seg_duration": args.seg_duration, "segmentation_step": args.segmentation_step,
                        "batch_size": args.batch_size, "apply_median_filtering": args.apply_median_filtering}
    clustering_config = {"method": args.clustering_method, "min_speakers": args.min_speakers, "max_speakers": args.max_speakers}
    if args.clustering_method == "AgglomerativeClustering":
        clustering_config.update({"ahc_threshold": args.ahc_threshold, "min_cluster_size": args.min_cluster_size})
    elif args.clustering_method == "VBxClustering":
        clustering_config.update({"ahc_criterion": args.ahc_criterion, "ahc_threshold": args.ahc_threshold, "Fa": args.Fa,
                                  "Fb": args.Fb, "lda_dim": args.lda_dim, "max_iters": args.max_iters})
    else:
        raise ValueError(f"Unsupported clustering method: {args.clustering_method}")
    return {"inference": {"args": inference_config}, "clustering": {"args": clustering_config}}


def checkpoints_from_args(args) -> Optional[list]:
    """infer_avg.py:265-286: the checkpoint(s) of an experiment directory that make up the segmentation model."""
    if args.configuration is None:
        return None
    from .checkpoints import load_metric_summary, select_checkpoints
    if args.val_metric_summary:
        ckpt_path = Path(args.configuration).expanduser().absolute().parent / "checkpoints"
        return select_checkpoints(load_metric_summary(args.val_metric_summary, ckpt_path), args.val_metric, args.val_mode,
                                  args.avg_ckpt_num)
    if not args.segmentation_model:
        raise SystemExit("experiment mode needs --segmentation_model or --val_metric_summary")
    return [args.segmentation_model]


def main(argv=None, pipeline_factory=None) -> int:
    args = build_parser().parse_args(argv)
    print(args)
    config_parse = config_from_args(args)
    if pipeline_factory is None:
        from .pipeline import DiariZenPipeline, _load_toml
        pipeline_factory = DiariZenPipeline
    else:
        _load_toml = None
    if args.configuration is not None:
        # experiment mode: model section from the training configuration, weights from (averaged) checkpoints
        if args.out_dir is None:
            raise SystemExit("experiment mode needs --out_dir")
        exp = _load_toml(Path(args.configuration).expanduser().absolute()) if _load_toml else {"model": {"args": {}}}
        config = {"model": exp["model"], **config_parse}
        if args.clustering_method == "VBxClustering":
            if not args.diarizen_hub:
                raise SystemExit("VBxClustering needs --diarizen_hub (for <hub>/plda)")
            config["clustering"]["args"]["plda_dir"] = os.path.join(args.diarizen_hub, "plda")
        pipe = pipeline_factory(None, args.embedding_model, rttm_out_dir=args.out_dir, precision=args.precision, _config=config,
                                segmentation=checkpoints_from_args(args))
        # a multi-channel model (recipes/diar_ssl_mc: `channel_fusion_*` in the model section) uses every channel of the file
        run = pipe.diarize_session if getattr(pipe, "_segmentation_mc", None) is not None else pipe
        for sess, in_wav in load_scp(args.in_wav_scp).items():
            print(f"Diarizing Session: {sess}")
            run(in_wav, sess_name=sess)
        return 0
    if not args.diarizen_hub:
        raise SystemExit("hub mode needs --diarizen_hub")
    pipe = pipeline_factory(Path(args.diarizen_hub), args.embedding_model, config_parse=config_parse, rttm_out_dir=args.rttm_out_dir,
                            precision=args.precision)
    for audio_file in scp2path(args.in_wav_scp):
        sess_name = Path(audio_file).stem.split(".")[0]
        print(f"Prosessing: {sess_name}")
        pipe(audio_file, sess_name=sess_name)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
