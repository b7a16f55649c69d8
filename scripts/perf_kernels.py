"""Kernel micro-benchmarks at the SCAIL-14B shapes (config A).  CUDA-event timing, warm-up, inputs > L2."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = "cuda"
    res = {}
    M = 27904 * int(os.environ.get("B", "2"))
    d, f = 5120, 13824
    for name, (N, K, epi) in {"qkv": (3 * d, d, 0), "out": (d, d, 2), "fc1": (f, d, 1), "fc2": (d, f, 2)}.items():
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.01
        b = torch.randn(N, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = {}
        if epi == 2:
            kw = dict(gate=torch.randn(2, N, device=dev, dtype=torch.bfloat16), residual=out, rows_per_batch=27904)
        ms = timeit(lambda: ops.gemm(a, w, b, out=out, epilogue=epi, **kw))
        tf = 2 * M * N * K / ms / 1e9
        ms_t = timeit(lambda: torch.matmul(a, w.t(), out=out))
        res["gemm_" + name] = dict(ms=round(ms, 3), tflops=round(tf, 1), cublas_ms=round(ms_t, 3),
                                   cublas_tflops=round(2 * M * N * K / ms_t / 1e9, 1))
        print(name, res["gemm_" + name], flush=True)
        del a, w, out
    B, H, N = int(os.environ.get("B", "2")), 40, 27904
    qkv = torch.randn(B * N, 3 * d, device=dev, dtype=torch.bfloat16)
    out = torch.empty(B * N, d, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, H, N, N), iters=3, warm=1)
    fl = 4 * B * H * N * N * 128
    res["attn_self"] = dict(ms=round(ms, 3), tflops=round(fl / ms / 1e9, 1))
    print("attn_self", res["attn_self"], flush=True)
    q4 = qkv.view(B, N, 3, H, 128)
    qh, kh, vh = (q4[:, :, i].transpose(1, 2).contiguous() for i in range(3))
    ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), iters=3, warm=1)
    res["attn_self"]["torch_sdpa_ms"] = round(ms_t, 3)
    res["attn_self"]["torch_sdpa_tflops"] = round(fl / ms_t / 1e9, 1)
    print("attn_self", res["attn_self"], flush=True)
    x = torch.randn(B, N, d, device=dev, dtype=torch.bfloat16)
    mod = torch.randn(B, 6, d, device=dev, dtype=torch.bfloat16)
    o = torch.empty_like(x)
    ms = timeit(lambda: ops.ln_modulate(x, out=o, shift=mod[:, 0], scale=mod[:, 1]))
    res["ln_modulate"] = dict(ms=round(ms, 3), gbs=round(2 * x.numel() * 2 / ms / 1e6, 1))
    wq = torch.ones(d, device=dev, dtype=torch.bfloat16)
    cos = torch.randn(N, 128, device=dev)
    ms = timeit(lambda: ops.rmsnorm_rope(qkv, N, d, [(0, wq), (d, wq)], cos, cos))
    res["rmsnorm_rope"] = dict(ms=round(ms, 3), gbs=round(2 * 2 * B * N * d * 2 / ms / 1e6, 1))
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_kernels.json", "w"), indent=1)


if __name__ == "__main__":
    main()
