"""Wan2.1 VAE decode timing at config 5 (latent [1,16,21,64,64] -> [1,3,81,512,512]), random weights."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scail_b200 import ops  # noqa: E402
from scail_b200.wan_vae import WanVAE  # noqa: E402

VAE_TFLOP_A = 180.5


def main():
    T, h, w = int(os.environ.get("T", 21)), int(os.environ.get("H", 64)), int(os.environ.get("W", 64))
    torch.manual_seed(7)
    vae = WanVAE(dim=96)
    z = torch.randn(16, T, h, w, device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        out = vae.decode([z])
        torch.cuda.synchronize()
        print("out", tuple(out.shape), "finite", bool(torch.isfinite(out).all()), "mem GB", torch.cuda.max_memory_allocated() / 2**30)
        del out
        times = []
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ops.LAUNCHES
            s.record()
            out = vae.decode([z])
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
            del out
    ms = min(times)
    scale = (T * h * w) / (21 * 64 * 64)
    res = {"latent": [T, h, w], "ms": ms, "all_ms": times, "launches": ops.LAUNCHES - l0,
           "tflops": VAE_TFLOP_A * scale / ms * 1e3 if (T, h, w) == (21, 64, 64) else None}
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_vae.json", "w"))


if __name__ == "__main__":
    main()
