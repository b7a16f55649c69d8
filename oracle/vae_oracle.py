"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32) of the Wan2.1 VAE decode path
(sgm/models/wan_vae.py:544-568 WanVAE_.decode and everything it calls).  Only tests/, smoke() and bench.py's
CPU-baseline legs may import it.

Parity pin: checked against tests/golden/vae_small.pt, which tests/golden/gen_golden.py produced by running
the UNMODIFIED reference (chunked decode with the 2-frame feature cache).

The reference decodes latent frame by latent frame with a feature cache (CACHE_T=2).  Restated here as
WHOLE-SEQUENCE causal convolutions (left zero pad of 2 frames), which is mathematically identical, with the one
exception the reference makes: in `upsample3d` the first latent frame skips `time_conv` ('Rep' sentinel,
wan_vae.py:105-108) and the temporal history of later frames starts from zeros at frame 1 (:120-131).
`sd` uses the reference's parameter names (decoder.*, conv2.*), values fp32.
"""
import torch
import torch.nn.functional as F

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
        0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]  # wan_vae.py:630-633
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
       3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]  # :634-637


def causal_conv3d(x, w, b):
    """CausalConv3d.forward (wan_vae.py:17-36): pad (w,w,h,h,2*pt,0) then a plain conv."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    return F.conv3d(x, w, b)


def rms_norm(x, gamma):
    """RMS_norm.forward (:39-54), channel_first."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def residual_block(sd, p, x):
    """ResidualBlock.forward (:186-220)."""
    h = causal_conv3d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"]) if p + ".shortcut.weight" in sd else x
    y = causal_conv3d(F.silu(rms_norm(x, sd[p + ".residual.0.gamma"])), sd[p + ".residual.2.weight"], sd[p + ".residual.2.bias"])
    y = causal_conv3d(F.silu(rms_norm(y, sd[p + ".residual.3.gamma"])), sd[p + ".residual.6.weight"], sd[p + ".residual.6.bias"])
    return y + h


def attention_block(sd, p, x):
    """AttentionBlock.forward (:223-262): per-frame single-head attention over h*w tokens."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[p + ".norm.gamma"])
    qkv = F.conv2d(y, sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"]).reshape(b * t, 1, 3 * c, h * w).permute(0, 1, 3, 2)
    q, k, v = qkv.chunk(3, -1)
    o = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    o = F.conv2d(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    return o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def resample(sd, p, x, mode):
    """Resample.forward (:101-160) for upsample2d / upsample3d, whole sequence."""
    b, c, t, h, w = x.shape
    if mode == "upsample3d":
        if t > 1:
            y = causal_conv3d(x[:, :, 1:], sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"])  # frames 1.. only
            y = y.reshape(b, 2, c, t - 1, h, w)
            y = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, 2 * (t - 1), h, w)  # :134-137
            x = torch.cat([x[:, :, :1], y], 2)
        t = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(y, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], padding=1)
    return y.reshape(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def decode(sd, z, temperal_upsample=(True, True, False)):
    """WanVAE_.decode (:544-568) + WanVAE.decode's .float().clamp_(-1,1) (:659-666).  z [b,16,T,h,w]."""
    sd = {k: v.float() for k, v in sd.items()}
    mean, std = torch.tensor(MEAN).view(1, 16, 1, 1, 1), torch.tensor(STD).view(1, 16, 1, 1, 1)
    x = z.float() / (1.0 / std) + mean
    x = causal_conv3d(x, sd["conv2.weight"], sd["conv2.bias"])
    x = causal_conv3d(x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"])
    x = residual_block(sd, "decoder.middle.0", x)
    x = attention_block(sd, "decoder.middle.1", x)
    x = residual_block(sd, "decoder.middle.2", x)
    idx = 0
    for stage in range(4):
        for _ in range(3):
            x = residual_block(sd, f"decoder.upsamples.{idx}", x)
            idx += 1
        if stage < 3:
            x = resample(sd, f"decoder.upsamples.{idx}", x, "upsample3d" if temperal_upsample[stage] else "upsample2d")
            idx += 1
    x = F.silu(rms_norm(x, sd["decoder.head.0.gamma"]))
    x = causal_conv3d(x, sd["decoder.head.2.weight"], sd["decoder.head.2.bias"])
    return x.float().clamp_(-1, 1)


def downsample(sd, p, x, mode):
    """Resample.forward (:101-160) for downsample2d / downsample3d, whole sequence.
    Spatial: ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) per frame (:87-96).  Temporal (downsample3d): the first frame
    bypasses time_conv (feat_cache None -> stored, :146-148); every later chunk convolves [last frame of the previous
    chunk] + its own frames with a 3x1x1 kernel at stride 2 and no padding (:151-159), i.e. output k >= 1 is the window
    of frames (2k-2, 2k-1, 2k) of the whole sequence."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.conv2d(F.pad(y, (0, 1, 0, 1)), sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], stride=2)
    y = y.reshape(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
    if mode == "downsample3d" and t > 1:
        z = F.conv3d(y, sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"], stride=(2, 1, 1))
        y = torch.cat([y[:, :, :1], z], 2)
    return y


def encode(sd, x, temperal_downsample=(False, True, True)):
    """WanVAE_.encode (:516-542): x [b,3,T,H,W] (T = 1 + 4k) -> mu [b,16,1+k,H/8,W/8], scaled by (mu - mean) / std."""
    sd = {k: v.float() for k, v in sd.items()}
    x = causal_conv3d(x.float(), sd["encoder.conv1.weight"], sd["encoder.conv1.bias"])
    idx = 0
    for stage in range(4):
        for _ in range(2):
            x = residual_block(sd, f"encoder.downsamples.{idx}", x)
            idx += 1
        if stage < 3:
            x = downsample(sd, f"encoder.downsamples.{idx}", x, "downsample3d" if temperal_downsample[stage] else "downsample2d")
            idx += 1
    x = residual_block(sd, "encoder.middle.0", x)
    x = attention_block(sd, "encoder.middle.1", x)
    x = residual_block(sd, "encoder.middle.2", x)
    x = F.silu(rms_norm(x, sd["encoder.head.0.gamma"]))
    x = causal_conv3d(x, sd["encoder.head.2.weight"], sd["encoder.head.2.bias"])
    mu = causal_conv3d(x, sd["conv1.weight"], sd["conv1.bias"])[:, :16]
    mean, std = torch.tensor(MEAN).view(1, 16, 1, 1, 1), torch.tensor(STD).view(1, 16, 1, 1, 1)
    return (mu - mean) * (1.0 / std)
