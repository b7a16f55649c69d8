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


def prepare_context(cond, uc):
    """VanillaCFG.prepare_inputs (guiders.py:47-57) for the 'crossattn' key: the batch is [uncond, cond]; when the two text
    contexts differ in length the UNCOND one is extended by repeating its last token row |Lc - Lu| times (guiders.py:52-53)
    before the concat (so, as in the reference, an uncond context longer than the cond one is an error)."""
    u, c = uc["crossattn"], cond["crossattn"]
    if u.shape[1] != c.shape[1]:
        u = torch.cat([u, u[:, -1:].repeat(1, abs(c.shape[1] - u.shape[1]), 1)], dim=1)
    return torch.cat((u, c), 0)


def sampler_step(model, x, sigma, next_sigma, cond, uc, scale=4.0, _ctx=None, **model_kwargs):
    """One reference sampler step: x [1,t,16,h,w] fp32 (updated in place and returned).
    cond/uc: dicts with 'crossattn' [1,L,text_dim]; cond also carries ref_concat, concat_smpl_render,
    image_clip_features, concat_images (shared by both CFG branches, guiders.py:50-56)."""
    x2 = torch.cat([x, x], 0)
    ts = torch.full((2,), float(sigma) * 1000.0, device=x.device, dtype=torch.float32)
    ctx = _ctx if _ctx is not None else prepare_context(cond, uc)
    v = model(x2, timesteps=ts, context=ctx, y=None, ref_concat=cond["ref_concat"],
              concat_smpl_render=cond["concat_smpl_render"], image_clip_features=cond["image_clip_features"],
              concat_images=cond.get("concat_images"), **model_kwargs)
    return ops.cfg_euler_(x, v.contiguous(), scale, float(next_sigma) - float(sigma))


def sample(model, x, cond, uc, num_steps=50, shift_scale=5.0, scale=4.0):
    """RFSampler.__call__ (sampling.py:965-982): the full Euler loop.  With `adaln_layer.cache_cross_kv` the step-invariant
    conditioning (text / CLIP embeddings and all layers' cross-attention K,V) is computed once for this call."""
    sig = make_flow_timesteps(num_steps, shift_scale)
    ctx = prepare_context(cond, uc)
    cache = getattr(model.mixins["adaln_layer"], "cache_cross_kv", False) if hasattr(model, "mixins") else False
    if cache:
        model.set_conditioning(ctx, cond["image_clip_features"], batch=2)
    try:
        for i in range(num_steps):
            x = sampler_step(model, x, sig[i], sig[i + 1], cond, uc, scale, _ctx=ctx)
    finally:
        if cache:
            model.clear_conditioning()
    return x


class HostStep:
    """End-to-end step through host buffers: pinned host -> device copies of the step's inputs, one sampler
    step, device -> pinned host copy of the updated latent.  This is what bench.py's `e2e` times."""

    def __init__(self, model, host_inputs, device="cuda", step_fn=None):
        self.model, self.device = model, device
        self.step_fn = step_fn or sampler_step  # bench.py's library arm passes baseline.torchlib.sampler_step
        self.host = {k: v.pin_memory() for k, v in host_inputs.items()}
        self.out_host = torch.empty_like(self.host["x"]).pin_memory()
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.d2h_bytes = self.out_host.numel() * self.out_host.element_size()

    def __call__(self, sigma, next_sigma, scale=4.0):
        d = {k: v.to(self.device, non_blocking=True) for k, v in self.host.items()}
        cond = dict(crossattn=d["context_cond"], ref_concat=d["ref_concat"], concat_smpl_render=d["concat_smpl_render"],
                    image_clip_features=d["image_clip_features"], concat_images=None)
        x = self.step_fn(self.model, d["x"], sigma, next_sigma, cond, dict(crossattn=d["context_uncond"]), scale)
        self.out_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.out_host
