"""ctypes binding of ``libradvlm_b200.so`` (the C ABI declared in ``include/radvlm_b200.h``).

There is no CPU fallback: if the shared library is missing, importing the product path raises with the
build instruction.  The planner entry points are CPU-only and work without a GPU; every GPU entry
point returns ``RADVLM_ERR_UNSUPPORTED_DEVICE`` when the current device is not sm_100.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RADVLM_B200_LIB: tuning builds only (tools/build_variant.sh); the shipped library is the in-tree one.
LIB_PATH = os.environ.get("RADVLM_B200_LIB") or os.path.join(_HERE, "libradvlm_b200.so")

OK = 0
ERR_BAD_ARGUMENT = 1
ERR_UNSUPPORTED_SHAPE = 2
ERR_CUDA = 3
ERR_WORKSPACE_TOO_SMALL = 4
ERR_UNSUPPORTED_DEVICE = 5

DT_F32, DT_BF16, DT_F16 = 0, 1, 2

EPI_BIAS_BF16 = 0
EPI_GELU_TANH_BF16 = 1
EPI_GELU_ERF_BF16 = 2
EPI_RESID_F32 = 3
EPI_POS_F32 = 4
EPI_BIAS_F32 = 6
EPI_ATOMIC_F32 = 7
EPI_BIAS_F16 = 9

SEG_PAD, SEG_TEXT, SEG_IMAGE = 0, 1, 2
MERGE_ANYRES, MERGE_SINGLE, MERGE_FLAT, MERGE_VIDEO = 0, 1, 2, 3
MAX_PEERS = 8
POOL_NONE, POOL_BILINEAR, POOL_AVERAGE, POOL_MAX = 0, 1, 2, 3
NEWLINE_NONE, NEWLINE_ONE, NEWLINE_FRAME, NEWLINE_GRID = 0, 1, 2, 3
ANYRES_NO_NEWLINE, ANYRES_NO_BASE = 1, 2


class RadvlmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("radvlm_b200 status %d: %s" % (status, message))
        self.status = status
        self.message = message


class VitLayerWeights(C.Structure):
    _fields_ = [
        ("ln1_gamma", C.c_void_p), ("ln1_beta", C.c_void_p),
        ("qkv_w", C.c_void_p), ("qkv_b", C.c_void_p),
        ("out_w", C.c_void_p), ("out_b", C.c_void_p),
        ("ln2_gamma", C.c_void_p), ("ln2_beta", C.c_void_p),
        ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p),
        ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
        # LayerNorm folded into the QKV / fc1 GEMMs (all None: stand-alone LayerNorm kernels)
        ("qkv_wf", C.c_void_p), ("qkv_sf", C.c_void_p), ("qkv_bf", C.c_void_p),
        ("fc1_wf", C.c_void_p), ("fc1_sf", C.c_void_p), ("fc1_bf", C.c_void_p),
    ]


class SiglipWeights(C.Structure):
    _fields_ = [
        ("hidden", C.c_int), ("intermediate", C.c_int), ("heads", C.c_int), ("num_layers", C.c_int),
        ("image_size", C.c_int), ("patch_size", C.c_int), ("channels", C.c_int), ("patch_k_pad", C.c_int),
        ("ln_eps", C.c_float),
        ("patch_w", C.c_void_p), ("patch_b", C.c_void_p), ("pos_embed", C.c_void_p),
        ("layers", C.POINTER(VitLayerWeights)),
    ]


class ProjectorWeights(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int), ("hidden", C.c_int),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
    ]


class VitLayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_gamma", "ln1_beta", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_gamma", "ln2_beta", "fc1_w", "fc1_b",
        "fc2_w", "fc2_b")]


class SiglipGrads(C.Structure):
    _fields_ = [("patch_w", C.c_void_p), ("patch_b", C.c_void_p), ("pos_embed", C.c_void_p),
                ("layers", C.POINTER(VitLayerGrads))]


class ProjectorGrads(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class ImagePlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "best_w", "best_h", "grid_w", "grid_h", "resized_w", "resized_h",
        "paste_x", "paste_y", "n_tiles", "crop_r0", "crop_c0", "crop_h", "crop_w", "pool",
        "out_h", "out_w", "n_tokens")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SpliceSegment(C.Structure):
    _fields_ = [
        ("dst_row", C.c_int64), ("length", C.c_int32), ("kind", C.c_int32), ("src_off", C.c_int32),
        ("image", C.c_int32), ("pos0", C.c_int32), ("reserved", C.c_int32),
    ]


class MergeImage(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "tile_base", "mode", "grid_w", "crop_r0", "crop_c0", "crop_h", "crop_w", "pool", "out_h",
        "out_w", "n_tokens", "reserved")]


class PreprocessImage(C.Structure):
    _fields_ = [
        ("src_offset", C.c_int64), ("scratch_offset", C.c_int64),
        ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
        ("grid_w", C.c_int32), ("grid_h", C.c_int32), ("resized_w", C.c_int32), ("resized_h", C.c_int32),
        ("paste_x", C.c_int32), ("paste_y", C.c_int32), ("tile_base", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/radvlm_b200.h declares
_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_pi = C.POINTER(C.c_int)
SIGNATURES = {
    "radvlm_last_error": (C.c_char_p, []),
    "radvlm_abi_version": (_i, []),
    "radvlm_profile_enable": (_i, [_i]),
    "radvlm_profile_read": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_int64), _i]),
    "radvlm_tmap_cache_stats": (_i, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "radvlm_gemm_bf16": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i, _vp, _i64, _vp, _i, _i, _vp]),
    "radvlm_gemm_bf16_ex": (_i, [_vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp, _i64, _vp, _i, _i, _vp]),
    "radvlm_gemm_bf16_ln": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp]),
    "radvlm_gemm_schedule_stats": (_i, [_i, _i, _i, _pi, _pi, _pi, _pi]),
    "radvlm_gemm_set_mode": (_i, [_i]),
    "radvlm_gemm_qkv_split": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "radvlm_attention_prepare_vt": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp]),
    "radvlm_attention_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "radvlm_attention_bwd_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "radvlm_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "radvlm_attention_fwd_lse": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "radvlm_layernorm_f32_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "radvlm_patch_im2col": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "radvlm_cast_f32_bf16": (_i, [_vp, _vp, _sz, _vp]),
    "radvlm_ln_row_stats_bf16": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "radvlm_encode_workspace_bytes": (_sz, [C.POINTER(SiglipWeights), C.POINTER(ProjectorWeights), _i]),
    "radvlm_siglip_tower_forward": (_i, [C.POINTER(SiglipWeights), _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "radvlm_projector_forward": (_i, [C.POINTER(ProjectorWeights), _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "radvlm_encode_images": (_i, [C.POINTER(SiglipWeights), C.POINTER(ProjectorWeights), _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "radvlm_tower_saved_bytes": (_sz, [C.POINTER(SiglipWeights), _i]),
    "radvlm_tower_saved_hidden_offset": (_sz, [C.POINTER(SiglipWeights), _i]),
    "radvlm_siglip_tower_forward_train": (_i, [C.POINTER(SiglipWeights), _vp, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "radvlm_tower_backward_workspace_bytes": (_sz, [C.POINTER(SiglipWeights), _i]),
    "radvlm_siglip_tower_backward": (_i, [C.POINTER(SiglipWeights), C.POINTER(SiglipGrads), _vp, _i, _i, _vp, _sz, _vp,
                                          _vp, _sz, _vp]),
    "radvlm_siglip_tower_backward_range": (_i, [C.POINTER(SiglipWeights), C.POINTER(SiglipGrads), _vp, _i, _i, _vp, _sz,
                                                _vp, _vp, _sz, _i, _i, _vp]),
    "radvlm_projector_backward_workspace_bytes": (_sz, [C.POINTER(ProjectorWeights), _i]),
    "radvlm_projector_backward": (_i, [C.POINTER(ProjectorWeights), C.POINTER(ProjectorGrads), _vp, _vp, _i, _vp, _vp,
                                       _sz, _vp]),
    "radvlm_colsum_bf16": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "radvlm_gelu_fwd_bwd_bf16": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "radvlm_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "radvlm_merge_splice_backward": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "radvlm_plan_select_best_resolution": (_i, [_i, _i, C.POINTER(C.c_int32), _i, _pi, _pi]),
    "radvlm_plan_image": (_i, [_i, _i, C.POINTER(C.c_int32), _i, _i, _i, _i, C.POINTER(ImagePlan)]),
    "radvlm_plan_splice": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int32), _i, _i64, _i,
                                C.POINTER(SpliceSegment), _i, _pi, C.POINTER(C.c_int32), _i, _pi,
                                C.POINTER(C.c_int32), _pi]),
    "radvlm_merge_splice": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i64,
                                 _vp, _vp, _vp, _vp, _i64, _vp]),
    "radvlm_merge_splice_scatter": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i64,
                                         C.POINTER(C.c_void_p), _i, _i, _vp, _vp, _vp, _i64, _vp]),
    "radvlm_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p), C.POINTER(C.c_uint8)]),
    "radvlm_peer_open": (_i, [C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]),
    "radvlm_peer_close": (_i, [_vp]),
    "radvlm_peer_free": (_i, [_vp]),
    "radvlm_peer_signal_wait": (_i, [_vp, _vp, _i, _i, C.c_uint64, C.c_double, _vp, _vp]),
    "radvlm_peer_copy": (_i, [_vp, _vp, _sz, _vp]),
    "radvlm_preprocess_scratch_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "radvlm_preprocess_anyres": (_i, [_vp, _vp, C.POINTER(PreprocessImage), _i, _i, _vp, _i, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and declare the argument types.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "radvlm_b200: %s is missing.  Build it with `make` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`).  There is no CPU or PyTorch fallback for the encode path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


PROF_CLASSES = ("gemm_patch_proj", "attention", "layernorm", "misc", "preprocess", "merge_splice",
                "gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2",
                "bwd_recompute", "bwd_dgrad", "bwd_wgrad", "bwd_attention", "bwd_elementwise")


PROFILING = False   # per-launch CUDA-event instrumentation on: the encoder then launches eagerly (no graph replay)


def profile_enable(on: bool):
    global PROFILING
    load().radvlm_profile_enable(1 if on else 0)
    PROFILING = bool(on)


def profile_read():
    """-> ({class: milliseconds}, {class: kernel launches}) since the previous read."""
    n = len(PROF_CLASSES)
    ms = (C.c_float * n)()
    cnt = (C.c_int64 * n)()
    check(load().radvlm_profile_read(ms, cnt, n))
    return dict(zip(PROF_CLASSES, [float(v) for v in ms])), dict(zip(PROF_CLASSES, [int(v) for v in cnt]))


def tmap_cache_stats():
    """-> (hits, misses) of the calling thread's TMA-descriptor cache (host_util.cu)."""
    h, m = C.c_uint64(0), C.c_uint64(0)
    check(load().radvlm_tmap_cache_stats(C.byref(h), C.byref(m)))
    return int(h.value), int(m.value)


def last_error() -> str:
    return load().radvlm_last_error().decode("utf-8", "replace")


def check(status: int):
    if status != OK:
        raise RadvlmError(status, last_error())
