"""ctypes binding of libdiarizen_b200.so (include/diarizen_b200.h).  No torch types cross this boundary:
tensors are passed as raw device pointers.  The library is built in-tree by `diarizen_b200/build.py`;
if it is missing the import fails loudly - there is no Python/CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiarizen_b200.so")

DZ_MAX_LAYERS = 32
DZ_MAX_HEADS = 16


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("npass", C.c_int32), ("batches", C.c_int32),
        ("groups", C.c_int32),
        ("a", C.c_void_p),
        ("a_plane", C.c_int64), ("a_bstride", C.c_int64), ("a_gstride", C.c_int64), ("a_rstride", C.c_int64),
        ("a_kinner", C.c_int32), ("fp16", C.c_int32),
        ("a_kouter", C.c_int64), ("a_rows_alloc", C.c_int64),
        ("b", C.c_void_p),
        ("b_plane", C.c_int64), ("b_gstride", C.c_int64),
        ("ldb", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p),
        ("alpha", C.c_float),
        ("group_cols", C.c_int32), ("out_row_off", C.c_int32), ("ldr", C.c_int32),
        ("residual", C.c_void_p),
        ("res_bstride", C.c_int64),
        ("out_f32", C.c_void_p),
        ("of_bstride", C.c_int64),
        ("ldo", C.c_int32), ("ldob", C.c_int32),
        ("out_bf", C.c_void_p),
        ("ob_plane", C.c_int64), ("ob_bstride", C.c_int64),
        ("out_planes", C.c_int32), ("zero_pad_to", C.c_int32),
        ("out_t", C.c_void_p),
        ("ot_plane", C.c_int64), ("ot_bstride", C.c_int64),
        ("ldt", C.c_int32), ("tr_col0", C.c_int32), ("seq_len", C.c_int32), ("act_after_res", C.c_int32),
        ("conv_runs", C.c_int32), ("conv_run_len", C.c_int32), ("conv_x0", C.c_int32), ("conv_h0", C.c_int32),
        ("conv_hs", C.c_int32), ("conv_Ho", C.c_int32), ("conv_H", C.c_int32), ("_pad2", C.c_int32),
        ("a_hstride", C.c_int64),
        ("res16", C.c_void_p),
        ("res16_plane", C.c_int64), ("res16_bstride", C.c_int64),
        ("ldr16", C.c_int32), ("res16_row_off", C.c_int32),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("_pad3", C.c_int32),
    ]

    @classmethod
    def default(cls) -> "GemmDesc":
        d = cls()
        d.npass = 1
        d.batches = 1
        d.groups = 1
        d.alpha = 1.0
        d.out_planes = 1
        d.seq_len = 1
        return d


class SegArchC(C.Structure):
    _fields_ = [
        ("large", C.c_int32),
        ("conv_channels", C.c_int32 * 7),
        ("embed_dim", C.c_int32), ("total_heads", C.c_int32), ("num_layers", C.c_int32),
        ("num_heads", C.c_int32 * DZ_MAX_LAYERS),
        ("head_index", (C.c_int32 * DZ_MAX_HEADS) * DZ_MAX_LAYERS),
        ("ffn", C.c_int32 * DZ_MAX_LAYERS),
        ("head_dim_model", C.c_int32), ("head_ffn", C.c_int32), ("head_heads", C.c_int32),
        ("head_layers", C.c_int32), ("head_kernel", C.c_int32), ("num_classes", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("nheads", C.c_int32),
        ("q", C.c_void_p), ("k", C.c_void_p), ("qk_plane", C.c_int64), ("ldqk", C.c_int32), ("q_col", C.c_int32),
        ("k_col", C.c_int32), ("fp16", C.c_int32),
        ("vt", C.c_void_p), ("vt_plane", C.c_int64), ("ldvt", C.c_int32), ("planes", C.c_int32),
        ("bias_tab", C.c_void_p), ("gate", C.c_void_p),
        ("out", C.c_void_p), ("out_plane", C.c_int64), ("ldo", C.c_int32), ("out_planes", C.c_int32),
        ("v", C.c_void_p), ("v_col", C.c_int32), ("_pad", C.c_int32),
    ]


class DzError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m diarizen_b200.build` (nvcc, sm_100a). "
            "diarizen_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.dz_last_error.restype = C.c_char_p
    L.dz_abi_version.restype = C.c_int
    L.dz_relpos_bucket.restype = C.c_int
    L.dz_relpos_bucket.argtypes = [C.c_int]
    L.dz_gemm.restype = C.c_int
    L.dz_gemm.argtypes = [C.POINTER(GemmDesc), C.c_int, C.c_int, C.c_void_p]
    L.dz_gemm_plan_create.restype = C.c_void_p
    L.dz_gemm_plan_create.argtypes = [C.POINTER(GemmDesc), C.c_int]
    L.dz_gemm_plan_launch.restype = C.c_int
    L.dz_gemm_plan_launch.argtypes = [C.c_void_p, C.c_void_p]
    L.dz_gemm_plan_destroy.restype = None
    L.dz_gemm_plan_destroy.argtypes = [C.c_void_p]
    L.dz_layernorm.restype = C.c_int
    L.dz_layernorm.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_float,
                               C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.dz_attention.restype = C.c_int
    L.dz_attention.argtypes = [C.POINTER(AttnArgs), C.c_int, C.c_int, C.c_void_p]
    L.dz_seg_create.restype = C.c_void_p
    L.dz_seg_create.argtypes = [C.POINTER(SegArchC), C.c_int, C.c_int, C.c_int]
    L.dz_seg_destroy.restype = None
    L.dz_seg_destroy.argtypes = [C.c_void_p]
    L.dz_seg_set_param.restype = C.c_int
    L.dz_seg_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.dz_seg_finalize.restype = C.c_int
    L.dz_seg_finalize.argtypes = [C.c_void_p]
    L.dz_seg_num_frames.restype = C.c_int
    L.dz_seg_num_frames.argtypes = [C.c_void_p, C.c_int]
    L.dz_seg_forward.restype = C.c_int
    L.dz_seg_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dz_seg_forward_host.restype = C.c_int
    L.dz_seg_forward_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.dz_seg_tap.restype = C.c_int64
    L.dz_seg_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.dz_seg_last_launches.restype = C.c_int
    L.dz_seg_last_launches.argtypes = [C.c_void_p]
    L.dz_seg_num_steps.restype = C.c_int
    L.dz_seg_num_steps.argtypes = [C.c_void_p]
    L.dz_seg_step_info.restype = C.c_int
    L.dz_seg_step_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.dz_seg_profile.restype = C.c_int
    L.dz_seg_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.dz_seg_plan.restype = C.c_int
    L.dz_seg_plan.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.dz_seg_tap_info.restype = C.c_int
    L.dz_seg_tap_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dz_seg_run_steps.restype = C.c_int
    L.dz_seg_run_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dz_fusion_create.restype = C.c_void_p
    L.dz_fusion_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.dz_fusion_destroy.restype = None
    L.dz_fusion_destroy.argtypes = [C.c_void_p]
    L.dz_fusion_set_param.restype = C.c_int
    L.dz_fusion_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.dz_fusion_finalize.restype = C.c_int
    L.dz_fusion_finalize.argtypes = [C.c_void_p]
    L.dz_fusion_forward.restype = C.c_int
    L.dz_fusion_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_float,
                                    C.c_void_p, C.c_void_p]
    L.dz_rows_to_planes.restype = C.c_int
    L.dz_rows_to_planes.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.dz_channel_mean.restype = C.c_int
    L.dz_channel_mean.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.dz_emb_create.restype = C.c_void_p
    L.dz_emb_create.argtypes = [C.c_int, C.c_int]
    L.dz_emb_destroy.restype = None
    L.dz_emb_destroy.argtypes = [C.c_void_p]
    L.dz_emb_set_param.restype = C.c_int
    L.dz_emb_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.dz_emb_finalize.restype = C.c_int
    L.dz_emb_finalize.argtypes = [C.c_void_p]
    L.dz_emb_num_fbank_frames.restype = C.c_int
    L.dz_emb_num_fbank_frames.argtypes = [C.c_int]
    L.dz_emb_forward.restype = C.c_int
    L.dz_emb_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.dz_emb_last_launches.restype = C.c_int
    L.dz_emb_last_launches.argtypes = [C.c_void_p]
    L.dz_emb_tap_fbank.restype = C.c_int64
    L.dz_emb_tap_fbank.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.dz_emb_num_steps.restype = C.c_int
    L.dz_emb_num_steps.argtypes = [C.c_void_p]
    L.dz_emb_profile.restype = C.c_int
    L.dz_emb_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    for name, args in {
        "dz_median_filter": [vp, vp, i32, i32, i32, i32, vp],
        "dz_speaker_count": [vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "dz_embedding_masks": [vp, i32, i32, i32, i32, vp, vp, vp],
        "dz_reconstruct": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
        "dz_pdist": [vp, i32, i32, vp, vp],
        "dz_linkage_centroid": [vp, i32, vp, vp, vp],
        "dz_linkage_centroid_variant": [vp, i32, vp, vp, vp, i32],
        "dz_assign": [vp, i32, i32, i32, vp, vp],
        "dz_conv3x3": [vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp],
        "dz_dendrogram_cut": [vp, i32, C.c_double, i32, i32, i32, i32, i32, vp, vp, vp, vp],
        "dz_vbx_model": [vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, vp],
        "dz_vbx_resp": [vp, vp, vp, vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, vp, vp],
    }.items():
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = args
    L.dz_linkage_workspace_bytes.restype = i64
    L.dz_linkage_workspace_bytes.argtypes = [i32]
    L.dz_dendrogram_cut_workspace_bytes.restype = i64
    L.dz_dendrogram_cut_workspace_bytes.argtypes = [i32]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc < 0:
        raise DzError(lib().dz_last_error().decode("utf-8", "replace") + f" (code {rc})")


EXPORTS = [
    "dz_last_error", "dz_abi_version", "dz_gemm", "dz_gemm_plan_create", "dz_gemm_plan_launch", "dz_gemm_plan_destroy",
    "dz_layernorm", "dz_attention",
    "dz_seg_create", "dz_seg_destroy", "dz_seg_set_param", "dz_seg_finalize", "dz_seg_num_frames",
    "dz_seg_forward", "dz_seg_forward_host", "dz_seg_tap", "dz_seg_last_launches",
    "dz_seg_num_steps", "dz_seg_step_info", "dz_seg_profile", "dz_seg_plan", "dz_seg_tap_info", "dz_seg_run_steps",
    "dz_fusion_create", "dz_fusion_destroy", "dz_fusion_set_param", "dz_fusion_finalize", "dz_fusion_forward", "dz_channel_mean", "dz_rows_to_planes",
    "dz_emb_create", "dz_emb_destroy", "dz_emb_set_param", "dz_emb_finalize", "dz_emb_num_fbank_frames", "dz_emb_forward",
    "dz_emb_last_launches", "dz_emb_tap_fbank", "dz_emb_num_steps", "dz_emb_profile",
    "dz_median_filter", "dz_speaker_count", "dz_embedding_masks", "dz_reconstruct", "dz_pdist", "dz_linkage_workspace_bytes",
    "dz_conv3x3", "dz_linkage_centroid", "dz_linkage_centroid_variant", "dz_dendrogram_cut_workspace_bytes", "dz_dendrogram_cut", "dz_assign", "dz_vbx_model", "dz_vbx_resp",
]
