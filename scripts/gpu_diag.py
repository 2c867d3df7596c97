"""First-contact diagnostics on the GPU box: runs the GEMM and segmentation checks without stopping at the
first failure and prints per-tap errors against the oracle (run under gpurun)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from diarizen_b200.archs import get_arch, init_state_dict
from diarizen_b200.segmentation import SegmentationModel
from oracle.seg_oracle import seg_forward


def seg_diag(name, B, N, precision, gemm_impl, attn_impl):
    print(f"=== {name} B={B} N={N} {precision} gemm={gemm_impl} attn={attn_impl}", flush=True)
    a = get_arch(name)
    sd = init_state_dict(a, 1)
    wav = 0.1 * torch.randn(B, N, generator=torch.Generator().manual_seed(1234))
    taps = {}
    ref = seg_forward(a, sd, wav, taps)
    try:
        m = SegmentationModel(a, sd, precision=precision, gemm_impl=gemm_impl, attn_impl=attn_impl)
        logp, ml = m.hard(wav.unsqueeze(1))
        torch.cuda.synchronize()
        print("  logp max err %.3e  (launches %d)" % ((logp.cpu() - ref).abs().max().item(), m.last_launches), flush=True)

        def cmp(tap, refv):
            got = m.tap(tap).cpu().view(refv.shape)
            sc = refv.abs().max().item() + 1e-9
            print("  tap %-10s max abs err %.3e  (scale %.3e)  nan=%d" % (tap, (got - refv).abs().max().item(), sc, int(torch.isnan(got).sum())), flush=True)
        cmp("feats_raw", taps["feats"] / sd["wavlm_model.feature_extractor.dummy_weight"])
        cmp("proj", taps["proj"])
        for i, r in enumerate(taps["reps"]):
            cmp(f"rep{i}", r)
        cmp("head_in", taps["head_in"])
        cmp(f"C{a.head_layers - 1}_out", taps["head_out"])
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["tiny"]
    if "tiny" in which:
        for gi in ("simt", "tc"):
            for nm in ("tiny_base", "tiny_large"):
                seg_diag(nm, 2, 16000, "bf16x3", gi, "simt")
        seg_diag("tiny_base", 2, 16000, "bf16", "tc", "simt")
    if "s80" in which:
        seg_diag("wavlm_base_s80_md", 2, 80000, "bf16x3", "tc", "simt")
        seg_diag("wavlm_base_s80_md", 2, 80000, "bf16", "tc", "simt")
        seg_diag("wavlm_large_s80_md", 1, 64000, "bf16x3", "tc", "simt")
