"""Per-substep clock64 trace of attention CTA (0,0,0).  Needs the experiment hooks compiled in
(-DSCAIL_ATTN_EXPERIMENTS; scripts/build_variants.sh builds libscail_b200_v2x.so; run with SCAIL_LIB_VARIANT=v2x).
Ablation modes: SCAIL_ATTN_DEBUG=1 softmax skipped, 5 no QK^T UMMAs, 6 no PV UMMAs."""
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
base = int(t[16, 0])
print("s: MMA: P(t0) wait start, end(+dt) | P(t1) wait start, end(+dt) || softmax0: S wait start, wake(+dt)")
for s in range(16, 40):
    r = [int(x) - base for x in t[s]]
    print(s, r[0], "(+%d)" % (r[1] - r[0]), "|", r[2], "(+%d)" % (r[3] - r[2]), "||", r[4], "(+%d)" % (r[5] - r[4]))
print("cycles per 128-key step:", 2 * (int(t[56, 0]) - int(t[16, 0])) / 40)
