"""Self-attention at the SCAIL-14B shape (b=2, 40 heads, N=27904 by default; N=... env for config B): CUDA-event timing of
scail_attention, torch SDPA (cuDNN/flash) beside it, and the relL2 between the two outputs.  SCAIL_LIB_VARIANT picks an
A/B build of the library (scripts/build_variants.sh)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops
B, H, N, d = int(os.environ.get("B", 2)), 40, int(os.environ.get("N", 27904)), 5120
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * d, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B * N, d, device="cuda", dtype=torch.bfloat16)
f = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, H, N, N)


def timeit(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


ms = timeit(f)
fl = 4 * B * H * N * N * 128
res = dict(variant=os.environ.get("SCAIL_LIB_VARIANT", "product"), B=B, N=N, ms=round(ms, 3), tflops=round(fl / ms / 1e9, 1))
if os.environ.get("SDPA", "1") == "1":
    q4 = qkv.view(B, N, 3, H, 128)
    qh, kh, vh = (q4[:, :, i].transpose(1, 2) for i in range(3))
    g = lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    ms_t = timeit(g, 3)
    ref = g().transpose(1, 2).reshape(B * N, d)
    f(); torch.cuda.synchronize()
    rel = float((out.float() - ref.float()).norm() / ref.float().norm())
    res.update(torch_sdpa_ms=round(ms_t, 3), torch_sdpa_tflops=round(fl / ms_t / 1e9, 1), rel_vs_sdpa=rel)
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/perf_attn.jsonl", "a").write(json.dumps(res) + "\n")
