"""Per-iteration clock64 trace of attention CTA (0,0,0).  Needs the experiment hooks compiled in:
    SCAIL_NVCC_EXTRA=-DSCAIL_ATTN_EXPERIMENTS python -c "from scail_b200 import _lib; _lib.build(force=True)"
(ablation modes: SCAIL_ATTN_DEBUG=1 softmax skipped, 2 MMA ignores P barriers, 3 TMEM read only, 5 PV only, 6 QK only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops, _lib
B, H, N, d = 1, 2, 27904, 256
qkv = torch.randn(B * N, 3 * d, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B * N, d, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(64 * 8, device="cuda", dtype=torch.int64)
f = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], out, B, H, N, N)
f(); torch.cuda.synchronize()
_lib.lib().scail_debug_set_attention_trace(tr.data_ptr())
f(); torch.cuda.synchronize()
_lib.lib().scail_debug_set_attention_trace(None)
t = tr.view(64, 8).cpu()
base = int(t[8, 0])
print("j: mma_wait0_start  wait0_end(+dt)  s0_commit_issued  wait1_end | softmax0: wait_start  wake")
for j in range(8, 24):
    r = [int(x) - base for x in t[j]]
    print(j, r[0], r[1], "(+%d)" % (r[1] - r[0]), r[2], r[3], "|", r[4], r[5], "(+%d)" % (r[5] - r[4]))
per = (int(t[40, 0]) - int(t[8, 0])) / 32
print("cycles per kv step:", per)
