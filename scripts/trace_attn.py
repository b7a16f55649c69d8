"""Per-KV-step clock64 trace of attention CTA (0,0,0).  Needs the experiment hooks compiled in
(-DSCAIL_ATTN_EXPERIMENTS; scripts/build_variants.sh builds libscail_b200_x*.so; run with SCAIL_LIB_VARIANT=<tag>).
Ablation modes: SCAIL_ATTN_DEBUG=1 softmax skipped, 5 no QK^T UMMAs, 6 no PV UMMAs.
Stamps per step j: MMA warp 0 = before P(tile0) wait, 1 = PV0 issued, 2 = QK0(j+1) issued, 3 = PV1 issued;
softmax warp 0: 4 = before S wait, 5 = S seen, 6 = last P group handed over; TMA warp: 7 = K(j) load issued."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops, _lib
B, H, N, d = 1, int(os.environ.get("H", 2)), 27904, 128 * int(os.environ.get("H", 2))
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
print("variant", os.environ.get("SCAIL_LIB_VARIANT"), "debug", os.environ.get("SCAIL_ATTN_DEBUG", "0"))
print("j | MMA: P0wait@ PV0issued(+) QK0issued(+) PV1issued(+) | softmax0: Swait@ seen(+) done(+) | K(j) load@")
for j in range(16, 28):
    r = [int(x) - base for x in t[j]]
    print(j, "|", r[0], "+%d" % (r[1] - r[0]), "+%d" % (r[2] - r[1]), "+%d" % (r[3] - r[2]), "|", r[4], "+%d" % (r[5] - r[4]),
          "+%d" % (r[6] - r[5]), "|", r[7])
print("cycles per KV step:", (int(t[56, 0]) - int(t[16, 0])) / 40)
