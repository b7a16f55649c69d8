"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch, fp32/fp64) of the
reference's SCAIL DiT denoising forward.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import this module;
the product path (`scail_b200/`) never does and has no CPU fallback.

Parity pin: this restatement is checked against golden vectors produced by
running the UNMODIFIED reference (`tests/golden/gen_golden.py`, which imports
/root/reference through `oracle/ref_harness.py`); see tests/test_oracle_golden.py.
The reference itself ships no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to
/root/reference).  `sd` is a state_dict with the reference's parameter names
(SURVEY.md §8b), values fp32.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# geometry / index maps (integer work: must be bit-exact)
# ----------------------------------------------------------------------------


def segment_lengths(t, h, w, patch=(1, 2, 2)):
    """dit_video_crossattn_sc_xc.py:1557-1559 — (ref_length, seq_length, pose_length)."""
    pp = patch[0] * patch[1] * patch[2]
    return (1 * h * w // pp, t * h * w // pp, t * (h // 2) * (w // 2) // pp)


def token_coords(t, h, w, patch=(1, 2, 2)):
    """Integer (segment, t, y, x) coordinate of every token in ref||noise||pose
    order, derived from the rearranges at dit_video_crossattn_sc_xc.py:110-124.
    segment: 0=ref, 1=noise, 2=pose.  Returns int64 [N, 4]."""
    hp, wp = h // patch[1], w // patch[2]
    hq, wq = (h // 2) // patch[1], (w // 2) // patch[2]
    out = []
    for seg, tt, hh, ww in ((0, 1, hp, wp), (1, t, hp, wp), (2, t, hq, wq)):
        ti, yi, xi = torch.meshgrid(torch.arange(tt), torch.arange(hh), torch.arange(ww), indexing="ij")
        out.append(torch.stack([torch.full_like(ti, seg), ti, yi, xi], -1).reshape(-1, 4))
    return torch.cat(out, 0)


def rope_freq_grid(head_dim, n_t, n_h, n_w, theta=10000.0, t_start=1):
    """dit_video_crossattn_sc_xc.py:404-459 (interleaved_rope=True branch).
    Returns the [n_t, n_h, n_w, head_dim] angle table (fp32)."""
    dim_t = head_dim - 4 * (head_dim // 6)
    dim_h = (head_dim // 6) * 2
    dim_w = (head_dim // 6) * 2

    def fr(d):
        return 1.0 / (theta ** (torch.arange(0, d, 2)[: d // 2].float() / d))

    gt = torch.arange(t_start, t_start + n_t, dtype=torch.float32)
    gh = torch.arange(n_h, dtype=torch.float32)
    gw = torch.arange(n_w, dtype=torch.float32)
    ft = (gt[:, None] * fr(dim_t)[None]).repeat_interleave(2, -1)
    fh = (gh[:, None] * fr(dim_h)[None]).repeat_interleave(2, -1)
    fw = (gw[:, None] * fr(dim_w)[None]).repeat_interleave(2, -1)
    return torch.cat(
        [
            ft[:, None, None, :].expand(n_t, n_h, n_w, dim_t),
            fh[None, :, None, :].expand(n_t, n_h, n_w, dim_h),
            fw[None, None, :, :].expand(n_t, n_h, n_w, dim_w),
        ],
        -1,
    )


def rope_tables(head_dim, T, H, W, max_frames, max_h, max_w, h_shift=0, w_shift=0,
                global_h=0, global_w=120):
    """cos/sin tables [N, head_dim] (fp32) for the ref||noise||pose sequence.
    Restates Rotary3DPositionEmbeddingMixin.rotary / rotary_ref / rotary_pose
    (dit_video_crossattn_sc_xc.py:525-645) with the slices of :1566-1585.
    T,H,W are rope_T/H/W (patch-grid extents); max_* are the ctor grid extents
    (compressed_num_frames, latent_height//2, latent_width//2; :1390-1393)."""
    main = rope_freq_grid(head_dim, max_frames, max_h, max_w + 120, t_start=1)  # :424-426
    ext = rope_freq_grid(head_dim, 1, max_h, max_w, t_start=0)  # :428-430
    out = []
    for fn in (torch.cos, torch.sin):
        m, e = fn(main), fn(ext)
        ref = e[0:1, h_shift:H + h_shift, w_shift:W + w_shift].reshape(-1, head_dim)  # :581-586
        noise = m[:T, h_shift:H + h_shift, w_shift:W + w_shift].reshape(-1, head_dim)  # :544-549
        pose = m[:T, global_h + h_shift:global_h + H + h_shift,
                 global_w + w_shift:global_w + W + w_shift]  # :617-629
        pose = F.avg_pool2d(pose.permute(0, 3, 1, 2), kernel_size=2, stride=2).permute(0, 2, 3, 1)  # :630-634
        out.append(torch.cat([ref, noise, pose.reshape(-1, head_dim)], 0))
    return out[0], out[1]


def rotate_half(x):
    """dit_video_crossattn_sc_xc.py:336-340 (interleaved pairs)."""
    x1, x2 = x[..., 0::2], x[..., 1::2]
    return torch.stack((-x2, x1), -1).flatten(-2)


def unpatchify(x, ref_len, seq_len, T, H, W, c=16, patch=(1, 2, 2)):
    """dit_video_crossattn_sc_xc.py:764-784: 'b (t h w) (o p q c) -> b (t o) c (h p) (w q)'."""
    b = x.shape[0]
    o, p, q = patch
    x = x[:, ref_len:ref_len + seq_len]
    x = x.reshape(b, T, H, W, o, p, q, c)
    x = x.permute(0, 1, 4, 7, 2, 5, 3, 6)  # b t o c h p w q
    return x.reshape(b, T * o, c, H * p, W * q)


# ----------------------------------------------------------------------------
# float ops
# ----------------------------------------------------------------------------


def timestep_embedding(timesteps, dim, max_period=10000):
    """sgm/modules/diffusionmodules/util.py:207-231 (freqs fp64 -> args fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float64, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1).float()


def rmsnorm(x, w, eps=1e-6):
    """dit_video_crossattn_sc_xc.py:61-68 (fp32 math, weight multiply in fp32, one cast back)."""
    dt = x.dtype
    x = x.float()
    return (w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))).to(dt)


def layernorm(x, w=None, b=None, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def modulate(x, shift, scale):
    """dit_video_crossattn_sc_xc.py:760-761."""
    return x * (1 + scale) + shift


def linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def patch_embed(sd, images, ref_concat, pose):
    """ImagePatchEmbeddingMixin.word_embedding_forward, dit_video_crossattn_sc_xc.py:99-130.
    Conv3d with kernel=stride=(1,2,2) restated as an explicit patch gather + matmul."""

    def conv(x, wname):
        w, b = sd[wname + ".weight"], sd[wname + ".bias"]
        bsz, t, c, h, ww = x.shape
        x = x.reshape(bsz, t, c, h // 2, 2, ww // 2, 2).permute(0, 1, 3, 5, 2, 4, 6)  # b t h' w' c p q
        x = x.reshape(bsz, t * (h // 2) * (ww // 2), c * 4)
        return x @ w.reshape(w.shape[0], -1).t() + b

    main = conv(torch.cat([ref_concat, images], 1), "mixins.patch_embed.proj")
    return torch.cat([main, conv(pose, "mixins.patch_embed.proj_pose")], 1)


def heads(x, n_heads):
    b, n, d = x.shape
    return x.view(b, n, n_heads, d // n_heads).permute(0, 2, 1, 3)


def merge_heads(x):
    b, h, n, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, n, h * d)


USE_LIBRARY_SDPA = False  # bench.py's CPU-baseline leg sets this: full-N blocks cannot materialise the N x N scores


def sdpa(q, k, v):
    """sat/transformer_defaults.py:67-72 (non-causal, scale 1/sqrt(d), no mask)."""
    if USE_LIBRARY_SDPA or q.dtype != torch.float32:  # "reference as shipped" mode: the library SDPA the reference calls
        return F.scaled_dot_product_attention(q, k, v)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.softmax(s, -1) @ v


def self_attention(sd, l, x, n_heads, cos, sin):
    """AdaLNMixin.attention_forward (dit_video_crossattn_sc_xc.py:1058-1105) +
    Rotary3DPositionEmbeddingMixin.attention_fn (:653-757)."""
    p = f"transformer.layers.{l}.attention"
    qkv = linear(sd, p + ".query_key_value", x)
    q, k, v = qkv.chunk(3, -1)  # split_tensor_along_last_dim, stride 3
    q = rmsnorm(q, sd[f"mixins.adaln_layer.query_layernorm_list.{l}.weight"])
    k = rmsnorm(k, sd[f"mixins.adaln_layer.key_layernorm_list.{l}.weight"])
    q, k, v = heads(q, n_heads), heads(k, n_heads), heads(v, n_heads)
    q = q * cos + rotate_half(q) * sin
    k = k * cos + rotate_half(k) * sin
    return linear(sd, p + ".dense", merge_heads(sdpa(q, k, v)))


def cross_attention(sd, l, x, n_heads, text, clip):
    """AdaLNMixin.cross_attention_forward (dit_video_crossattn_sc_xc.py:1107-1203)."""
    p = f"transformer.layers.{l}.cross_attention"
    m = "mixins.adaln_layer"
    q = rmsnorm(linear(sd, p + ".query", x), sd[f"{m}.cross_query_layernorm_list.{l}.weight"])
    k, v = linear(sd, p + ".key_value", text).chunk(2, -1)
    k = rmsnorm(k, sd[f"{m}.cross_key_layernorm_list.{l}.weight"])
    kc, vc = linear(sd, f"{m}.clip_feature_key_value_list.{l}", clip).chunk(2, -1)
    kc = rmsnorm(kc, sd[f"{m}.clip_feature_key_layernorm_list.{l}.weight"])
    q = heads(q, n_heads)
    ctx = merge_heads(sdpa(q, heads(k, n_heads), heads(v, n_heads)))
    ctx = ctx + merge_heads(sdpa(q, heads(kc, n_heads), heads(vc, n_heads)))
    return linear(sd, p + ".dense", ctx)


def mlp(sd, l, x):
    """sat/transformer_defaults.py:172-175 with nn.GELU(approximate='tanh')
    (dit_video_crossattn_sc_xc.py:1296)."""
    p = f"transformer.layers.{l}.mlp"
    return linear(sd, p + ".dense_4h_to_h", F.gelu(linear(sd, p + ".dense_h_to_4h", x), approximate="tanh"))


def block(sd, l, x, adaln_emb, n_heads, cos, sin, text, clip):
    """AdaLNMixin.layer_forward, dit_video_crossattn_sc_xc.py:1009-1051 (share_adaln)."""
    d = x.shape[-1]
    mod = adaln_emb.unflatten(1, (6, d)) + sd[f"mixins.adaln_layer.adaLN_modulations.{l}"]
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, 1)
    a = modulate(layernorm(x), shift_msa, scale_msa)
    x = x + gate_msa * self_attention(sd, l, a, n_heads, cos, sin)
    pl = f"transformer.layers.{l}.post_cross_attention_layernorm"
    c = layernorm(x, sd[pl + ".weight"], sd[pl + ".bias"])
    x = x + cross_attention(sd, l, c, n_heads, text, clip)
    m = modulate(layernorm(x), shift_mlp, scale_mlp)
    return x + gate_mlp * mlp(sd, l, m)


def final_layer(sd, x, emb):
    """FinalLayerMixin.final_forward, dit_video_crossattn_sc_xc.py:818-826 (before unpatchify)."""
    shift, scale = (emb.unsqueeze(1) + sd["mixins.final_layer.adaLN_modulation"]).chunk(2, 1)
    return linear(sd, "mixins.final_layer.linear", modulate(layernorm(x), shift, scale))


def embeddings(sd, timesteps, context, clip_feats, time_freq_dim=256, dtype=torch.float32):
    """DiffusionTransformer.forward, dit_video_crossattn_sc_xc.py:1505-1555."""
    text = linear(sd, "text_embedding.2", F.gelu(linear(sd, "text_embedding.0", context), approximate="tanh"))
    c = layernorm(clip_feats, sd["clip_proj.proj.0.weight"], sd["clip_proj.proj.0.bias"], eps=1e-5)
    c = linear(sd, "clip_proj.proj.3", F.gelu(linear(sd, "clip_proj.proj.1", c)))
    clip = layernorm(c, sd["clip_proj.proj.4.weight"], sd["clip_proj.proj.4.bias"], eps=1e-5)
    t_emb = timestep_embedding(timesteps, time_freq_dim).to(device=context.device, dtype=dtype)
    emb = linear(sd, "time_embed.2", F.silu(linear(sd, "time_embed.0", t_emb)))
    adaln = linear(sd, "adaln_projection.1", F.silu(emb))
    return text, clip, emb, adaln


def dit_forward(sd, x, timesteps, context, ref_concat, concat_smpl_render, image_clip_features,
                n_heads, n_layers, max_frames=21, max_h=150, max_w=150, return_hidden=False, dtype=torch.float32):
    """DiffusionTransformer.forward (dit_video_crossattn_sc_xc.py:1452-1587) ->
    BaseTransformer.forward (sat/model/transformer.py:572-746), fp32, CPU.
    x [b,t,16,h,w]; ref_concat [1|b,1,16,h,w]; concat_smpl_render [1|b,t,16,h/2,w/2];
    context [b,L,text_dim]; image_clip_features [1|b,257,1280]; timesteps [b]."""
    # dtype=torch.bfloat16 runs the same op chain in bf16 (rounding after every op, like the reference as shipped:
    # SURVEY §3.5) — used only to put "reference-style bf16" error next to ours in the GPU parity tests
    sd = {k: v.to(dtype) for k, v in sd.items()}
    b, t, _, h, w = x.shape
    x = x.to(dtype)
    dev = x.device

    def rep(a):
        return a.to(dtype).repeat(b // a.shape[0], *([1] * (a.dim() - 1)))

    images = torch.cat([x, torch.zeros(b, t, 4, h, w, dtype=dtype, device=dev)], 2)  # :1468,1503
    ref = torch.cat([rep(ref_concat), torch.ones(b, 1, 4, h, w, dtype=dtype, device=dev)], 2)  # :1483-1486
    pose = torch.cat([rep(concat_smpl_render), torch.ones(b, t, 4, h // 2, w // 2, dtype=dtype, device=dev)], 2)  # :1496-1501
    text, clip, emb, adaln = embeddings(sd, timesteps, context.to(dtype), rep(image_clip_features), dtype=dtype)
    ref_len, seq_len, pose_len = segment_lengths(t, h, w)
    T, H, W = t, h // 2, w // 2
    d = sd["mixins.final_layer.linear.weight"].shape[1]
    cos, sin = rope_tables(d // n_heads, T, H, W, max_frames, max_h, max_w)
    cos, sin = cos.to(device=dev, dtype=dtype), sin.to(device=dev, dtype=dtype)  # :553-554 `.to(t.dtype)`
    hid = patch_embed(sd, images, ref, pose)
    hiddens = [hid]
    for l in range(n_layers):
        hid = block(sd, l, hid, adaln, n_heads, cos, sin, text, clip)
        hiddens.append(hid)
    out = unpatchify(final_layer(sd, hid, emb), ref_len, seq_len, T, H, W)
    return (out, hiddens) if return_hidden else out


# ----------------------------------------------------------------------------
# sampler pieces (fp32; tiny) — sgm/modules/diffusionmodules
# ----------------------------------------------------------------------------


def make_flow_timesteps(num_steps=50, shift_scale=5.0, t_start=0.0):
    """sampling.py:888-903, mode='normal': s = linspace(t_start, 1, n+1) in fp64,
    s / (shift + s - shift*s) -> fp32, sigma = 1 - that."""
    import numpy as np

    s = np.linspace(t_start, 1.0, num_steps + 1, endpoint=True)
    s = s / (shift_scale + s - shift_scale * s)
    return 1 - torch.tensor(s, dtype=torch.float32)


def cfg_euler_step(x, v_uncond, v_cond, sigma, sigma_next, scale=4.0):
    """guiders.py:41-45 + sampling_utils.py:7-10 + sampling.py:960-963."""
    v = v_uncond + scale * (v_cond - v_uncond)
    return x + (sigma_next - sigma) * v
