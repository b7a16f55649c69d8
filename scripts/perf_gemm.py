import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops
M, d, f = 55808, 5120, 13824
res = []
for name, (N, K, epi) in {"qkv": (3 * d, d, 0), "out": (d, d, 2), "fc1": (f, d, 1), "fc2": (d, f, 2)}.items():
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.01
    b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    kw = dict(gate=torch.randn(2, N, device="cuda", dtype=torch.bfloat16), residual=out, rows_per_batch=27904) if epi == 2 else {}
    fn = lambda: ops.gemm(a, w, b, out=out, epilogue=epi, **kw)
    fn(); fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    res.append("%s %.3f ms %.0f TF" % (name, ms, 2 * M * N * K / ms / 1e9))
    del a, w, out
print("group_m=%s: " % os.environ.get("SCAIL_GEMM_GROUP_M", "24") + " | ".join(res))
