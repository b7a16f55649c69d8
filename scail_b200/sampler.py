"""Rectified-flow Euler sampler step with classifier-free guidance: the public per-step API.

Mirrors the reference call chain (paths relative to /root/reference):
  RFSampler.sampler_step / denoise      sgm/modules/diffusionmodules/sampling.py:950-963
  VanillaCFG.prepare_inputs / __call__  sgm/modules/diffusionmodules/guiders.py:41-57
  Denoiser.forward + RFScaling          denoiser.py:25-43, denoiser_scaling.py:71-79  (c_in=c_out=1, c_skip=0, c_noise=1000*sigma)
  make_flow_timesteps                   sampling.py:888-903
The latent stays fp32 across steps (diffusion_video.py:470); the DiT consumes/produces bf16.
"""
import numpy as np
import torch

from . import ops


def make_flow_timesteps(num_steps=50, shift_scale=5.0, t_start=0.0):
    s = np.linspace(t_start, 1.0, num_steps + 1, endpoint=True)
    s = s / (shift_scale + s - shift_scale * s)
    return 1 - torch.tensor(s, dtype=torch.float32)


def sampler_step(model, x, sigma, next_sigma, cond, uc, scale=4.0):
    """One reference sampler step: x [1,t,16,h,w] fp32 (updated in place and returned).
    cond/uc: dicts with 'crossattn' [1,L,text_dim]; cond also carries ref_concat, concat_smpl_render,
    image_clip_features, concat_images (shared by both CFG branches, guiders.py:50-56)."""
    x2 = torch.cat([x, x], 0)
    ts = torch.full((2,), float(sigma) * 1000.0, device=x.device, dtype=torch.float32)
    ctx = torch.cat([uc["crossattn"], cond["crossattn"]], 0)
    v = model(x2, timesteps=ts, context=ctx, y=None, ref_concat=cond["ref_concat"],
              concat_smpl_render=cond["concat_smpl_render"], image_clip_features=cond["image_clip_features"],
              concat_images=cond.get("concat_images"))
    return ops.cfg_euler_(x, v.contiguous(), scale, float(next_sigma) - float(sigma))


def sample(model, x, cond, uc, num_steps=50, shift_scale=5.0, scale=4.0):
    """RFSampler.__call__ (sampling.py:965-982): the full Euler loop."""
    sig = make_flow_timesteps(num_steps, shift_scale)
    for i in range(num_steps):
        x = sampler_step(model, x, sig[i], sig[i + 1], cond, uc, scale)
    return x


class HostStep:
    """End-to-end step through host buffers: pinned host -> device copies of the step's inputs, one sampler
    step, device -> pinned host copy of the updated latent.  This is what bench.py's `e2e` times."""

    def __init__(self, model, host_inputs, device="cuda"):
        self.model, self.device = model, device
        self.host = {k: v.pin_memory() for k, v in host_inputs.items()}
        self.out_host = torch.empty_like(self.host["x"]).pin_memory()
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.d2h_bytes = self.out_host.numel() * self.out_host.element_size()

    def __call__(self, sigma, next_sigma, scale=4.0):
        d = {k: v.to(self.device, non_blocking=True) for k, v in self.host.items()}
        cond = dict(crossattn=d["context_cond"], ref_concat=d["ref_concat"], concat_smpl_render=d["concat_smpl_render"],
                    image_clip_features=d["image_clip_features"], concat_images=None)
        x = sampler_step(self.model, d["x"], sigma, next_sigma, cond, dict(crossattn=d["context_uncond"]), scale)
        self.out_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.out_host
