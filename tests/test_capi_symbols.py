"""The C-ABI library loads and exports every symbol include/radvlm_b200.h declares (no compute calls, no GPU)."""
import ctypes
import os
import re

from radvlm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "radvlm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(radvlm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_lib.ImagePlan) == 19 * 4
    assert ctypes.sizeof(_lib.SpliceSegment) == 32
    assert ctypes.sizeof(_lib.MergeImage) == 12 * 4
    assert ctypes.sizeof(_lib.PreprocessImage) == 16 + 10 * 4
    assert ctypes.sizeof(_lib.VitLayerWeights) == 18 * 8
    assert ctypes.sizeof(_lib.VitLayerGrads) == 12 * 8


def test_gpu_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    st = lib.radvlm_layernorm_f32_bf16(None, None, None, None, 1, 1152, 1e-6, None)
    assert st == _lib.ERR_UNSUPPORTED_DEVICE
    assert "no CPU fallback" in _lib.last_error() or "sm_" in _lib.last_error()
    from radvlm_b200 import synthetic, mm_arch
    import pytest
    host = synthetic.build_host(hidden_size=64, vocab=16, dtype=torch.float32, device="cpu",
                                vision_cfg=synthetic.siglip_config(hidden_size=144, intermediate_size=272,
                                                                   num_hidden_layers=1, num_attention_heads=2))
    with pytest.raises(RuntimeError):
        host.encode_images(torch.zeros(1, 3, 384, 384))


def test_scheduled_gemm_tile_lists_cover_every_tile_and_balance():
    """Host logic of the scheduled GEMM (no GPU): every (row block, column tile) is dealt exactly once, in order inside a
    pair's list, N = 1152 / 3456 are cut exactly (4 x 256 + 128, 13 x 256 + 128), and list scheduling leaves the pairs
    within one tile of each other."""
    import ctypes as C
    from radvlm_b200 import _lib
    lib = _lib.load()
    for M, N, pairs, want_tiles in [(58320, 1152, 74, 228 * 5), (58320, 3456, 74, 228 * 14), (58320, 4304, 74, 228 * 17),
                                    (7290, 1152, 74, 29 * 5), (1458, 144, 74, 6), (300, 200, 74, 2), (58320, 1152, 10, 1140)]:
        per = (C.c_int * pairs)()
        n, hi, lo = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.radvlm_gemm_schedule_stats(M, N, pairs, per, C.byref(n), C.byref(hi), C.byref(lo)))
        assert n.value == want_tiles == sum(per), (M, N, n.value, want_tiles)
        assert hi.value - lo.value <= 100, (M, N, hi.value, lo.value)
    # beyond the parameter block (4000 tiles): not covered, reported as such
    assert lib.radvlm_gemm_schedule_stats(600000, 4304, 74, None, None, None, None) == _lib.ERR_UNSUPPORTED_SHAPE
