import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops
B, H, N, d = 2, 40, 27904, 5120
qkv = torch.randn(B * N, 3 * d, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B * N, d, device="cuda", dtype=torch.bfloat16)
f = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, H, N, N)
f(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): f()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print("attn_self ms %.3f  TF/s %.1f" % (ms, 4 * B * H * N * N * 128 / ms / 1e9))
