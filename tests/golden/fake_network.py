"""Deterministic stand-in for the DiT used to pin the SAMPLERS (not the network) against the reference:
shared by tests/golden/gen_sampler_long.py (reference side) and tests/test_host_cpu.py (product side)."""
import torch


def fake_network(x, c_noise, cond):
    """Stand-in for the DiT: x [B,t,c,h,w], c_noise [B] (= 1000 sigma), cond['crossattn'] [B,L,D], cond['concat_smpl_render']
    [1|B,t,c,h/2,w/2] -> velocity [B,t,c,h,w].  Depends on every input the sampler is responsible for routing."""
    ctx = cond["crossattn"].float().mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)
    pose = cond["concat_smpl_render"].float()
    pose = torch.nn.functional.interpolate(pose.flatten(0, 1), scale_factor=2.0, mode="nearest").view(pose.shape[0], *x.shape[1:3], *x.shape[3:])
    t = (c_noise.float() / 1000.0).view(-1, 1, 1, 1, 1)
    return torch.sin(x.float() * (1.0 + t)) * 0.5 + 0.1 * ctx + 0.05 * pose * t
