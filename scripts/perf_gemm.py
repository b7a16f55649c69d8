"""One GEMM shape of the 14B step (SHAPE=qkv|out|fc1|fc2), `ITERS` back-to-back launches (default 40: long enough for the
power cap to settle), CUDA-event time; knobs via env: SCAIL_GEMM_CG, SCAIL_GEMM_GROUP_M, SCAIL_GEMM_L2_HINTS."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops
d, f, M = 5120, 13824, 2 * 27904
N, K, epi = {"qkv": (3 * d, d, 0), "out": (d, d, 2), "fc1": (f, d, 1), "fc2": (d, f, 2)}[os.environ.get("SHAPE", "qkv")]
it = int(os.environ.get("ITERS", 40))
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.01
b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = dict(gate=torch.randn(2, N, device="cuda", dtype=torch.bfloat16), residual=out, rows_per_batch=27904) if epi == 2 else {}
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
ms = t(lambda: ops.gemm(a, w, b, out=out, epilogue=epi, **kw))
res = dict(shape=os.environ.get("SHAPE", "qkv"), cg=os.environ.get("SCAIL_GEMM_CG"), group=os.environ.get("SCAIL_GEMM_GROUP_M"),
           hints=os.environ.get("SCAIL_GEMM_L2_HINTS"), ms=round(ms, 3), tflops=round(2 * M * N * K / ms / 1e9, 1))
if os.environ.get("CUBLAS"):
    ms_t = t(lambda: torch.nn.functional.linear(a, w, b))
    res["cublas_ms"] = round(ms_t, 3)
print(json.dumps(res), flush=True)
