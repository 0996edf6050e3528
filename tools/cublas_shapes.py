#!/usr/bin/env python
"""Context for the GEMM rooflines: cuBLAS (torch.matmul / F.linear, bf16) on the four layer shapes of an 80-tile tower
call, run back to back for ~2 s each (sustained, power-capped like a long step), next to this library's kernels with
their fused epilogues on the same shapes and the same schedule.  Output: one JSON line.  (A tuning tool, not a product
path; run under gpurun.)"""
import ctypes as C
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from radvlm_b200 import _lib  # noqa: E402

M = 58320
SHAPES = {"qkv": (3456, 1152), "out_proj": (1152, 1152), "fc1": (4304, 1152), "fc2": (1152, 4304)}


def sustained(fn, seconds=2.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n, t0 = 0, time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    out = {}
    st = torch.cuda.current_stream().cuda_stream
    for name, (N, K) in SHAPES.items():
        a = (torch.randn(M, K, device=dev, generator=g)).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * 0.03).bfloat16()
        b = torch.randn(N, device=dev, generator=g)
        flops = 2.0 * M * N * K
        ms_cublas = sustained(lambda: torch.nn.functional.linear(a, w))
        ms_cublas_bias = sustained(lambda: torch.nn.functional.linear(a, w, b.bfloat16()))
        if name in ("out_proj", "fc2"):
            res = torch.randn(M, N, device=dev, generator=g)
            o = torch.empty(M, N, device=dev, dtype=torch.float32)
            epi = _lib.EPI_RESID_F32
            fn = lambda: _lib.check(lib.radvlm_gemm_bf16(a.data_ptr(), K, w.data_ptr(), K, M, N, K, b.data_ptr(), epi,
                                                         o.data_ptr(), N, res.data_ptr(), 0, 0, st))
        else:
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            epi = _lib.EPI_GELU_TANH_BF16 if name == "fc1" else _lib.EPI_BIAS_BF16
            fn = lambda: _lib.check(lib.radvlm_gemm_bf16(a.data_ptr(), K, w.data_ptr(), K, M, N, K, b.data_ptr(), epi,
                                                         o.data_ptr(), N, None, 0, 0, st))
        ms_ours = sustained(fn)
        out[name] = {"M": M, "N": N, "K": K,
                     "cublas_tflops": flops / ms_cublas / 1e9, "cublas_with_bias_tflops": flops / ms_cublas_bias / 1e9,
                     "this_library_fused_epilogue_tflops": flops / ms_ours / 1e9,
                     "epilogue": {"qkv": "bias -> bf16", "out_proj": "bias + fp32 residual -> fp32", "fc1": "bias + GELU-tanh -> bf16",
                                  "fc2": "bias + fp32 residual -> fp32"}[name]}
        del a, w
    sq = torch.randn(8192, 8192, device=dev).bfloat16()
    out["cublas_8192_cubed_tflops"] = 2.0 * 8192 ** 3 / sustained(lambda: sq @ sq) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
