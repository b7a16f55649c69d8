"""Three-segment 3-D RoPE tables for the ref || noise || pose token sequence.

Host-side, computed once per latent geometry and cached (not per step).  Follows the reference's
own op order so the tables are bit-identical to what its mixin produces in fp32:
Rotary3DPositionEmbeddingMixin.__init__ (dit_video_crossattn_sc_xc.py:404-513, interleaved_rope branch)
and the slices / 2x2 average pooling of rotary / rotary_ref / rotary_pose (:525-645).
"""
import functools

import torch
import torch.nn.functional as F


def _axis_angles(pos, dim, theta):
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    return torch.einsum("..., f -> ... f", pos, inv).repeat_interleave(2, dim=-1)  # '... n -> ... (n r)', r=2


def _grid(head_dim, pos_t, pos_h, pos_w, theta):
    dim_t = head_dim - 4 * (head_dim // 6)
    dim_h = (head_dim // 6) * 2
    dim_w = (head_dim // 6) * 2
    at, ah, aw = _axis_angles(pos_t, dim_t, theta), _axis_angles(pos_h, dim_h, theta), _axis_angles(pos_w, dim_w, theta)
    nt, nh, nw = at.shape[0], ah.shape[0], aw.shape[0]
    return torch.cat([at[:, None, None, :].expand(nt, nh, nw, dim_t), ah[None, :, None, :].expand(nt, nh, nw, dim_h),
                      aw[None, None, :, :].expand(nt, nh, nw, dim_w)], dim=-1).contiguous()


@functools.lru_cache(maxsize=16)
def build_tables_cpu(head_dim, T, H, W, h_shift=0, w_shift=0, global_h=0, global_w=120, theta=10000.0):
    """Returns (cos, sin): fp32 [n_ref + n_seq + n_pose, head_dim] on CPU.
    T/H/W = rope_T / rope_H / rope_W of DiffusionTransformer.forward (:1566-1568)."""
    f32 = torch.float32
    # main grid: t in 1..T, h/w from 0 (:424-426); only the extent that is sliced is materialised
    main = _grid(head_dim, torch.arange(1, T + 1, dtype=f32), torch.arange(global_h + h_shift + H, dtype=f32),
                 torch.arange(global_w + w_shift + W, dtype=f32), theta)
    ext = _grid(head_dim, torch.tensor([0], dtype=f32), torch.arange(h_shift + H, dtype=f32),
                torch.arange(w_shift + W, dtype=f32), theta)  # :428-430
    out = []
    for fn in (torch.cos, torch.sin):
        m, e = fn(main), fn(ext)
        ref = e[0:1, h_shift:H + h_shift, w_shift:W + w_shift].reshape(-1, head_dim)
        noise = m[:T, h_shift:H + h_shift, w_shift:W + w_shift].reshape(-1, head_dim)
        pose = m[:T, global_h + h_shift:global_h + H + h_shift, global_w + w_shift:global_w + W + w_shift]
        pose = F.avg_pool2d(pose.permute(0, 3, 1, 2), kernel_size=2, stride=2).permute(0, 2, 3, 1)
        out.append(torch.cat([ref, noise, pose.reshape(-1, head_dim)], 0).contiguous())
    return out[0], out[1]


_DEV_CACHE = {}


def build_tables(device, head_dim, T, H, W, h_shift=0, w_shift=0, global_h=0, global_w=120, theta=10000.0):
    key = (str(device), head_dim, T, H, W, h_shift, w_shift, global_h, global_w, theta)
    if key not in _DEV_CACHE:
        cos, sin = build_tables_cpu(head_dim, T, H, W, h_shift, w_shift, global_h, global_w, theta)
        _DEV_CACHE[key] = (cos.to(device), sin.to(device))
    return _DEV_CACHE[key]
