"""The bench's library-path arm (baseline/torchlib.py: the reference's op chain on cuBLAS / SDPA / cuDNN) must compute the
SAME function as the product on the same weights — otherwise `library_baseline` / `vae_decode.library_ms` would compare
different work.  Both are bf16 chains, so they agree to a few bf16 ulps per layer."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm())


def test_torchlib_dit_forward_matches_product():
    from baseline import torchlib
    from scail_b200.dit import DiffusionTransformer
    torch.manual_seed(0)
    m = DiffusionTransformer(hidden_size=256, num_attention_heads=2, inner_hidden_size=512, num_layers=2, text_dim=64,
                             time_embed_dim=256).to(torch.bfloat16).cuda().eval()
    with torch.no_grad():
        for _, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    g = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).cuda()
    x, ref, pose = r(2, 3, 16, 8, 12), r(1, 1, 16, 8, 12), r(1, 3, 16, 4, 6)
    ctx, clip, ts = r(2, 24, 64), r(1, 257, 1280), torch.tensor([250.0, 250.0]).cuda()
    with torch.no_grad():
        ours = m(x, timesteps=ts, context=ctx, ref_concat=ref, concat_smpl_render=pose, image_clip_features=clip, concat_images=x)
        lib = torchlib.dit_forward(m, x, ts, ctx, ref, pose, clip)
    e = rel(ours, lib)
    print("DiT: product vs library chain relL2 %.3e" % e)
    assert ours.shape == lib.shape and e < 1.5e-2


def test_torchlib_vae_decode_matches_product():
    from baseline import torchlib
    from scail_b200.wan_vae import WanVAE
    torch.manual_seed(7)
    vae = WanVAE(dim=16)
    with torch.no_grad():
        for _, p in vae.model.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    z = torch.randn(16, 3, 8, 8, device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        ours = vae.decode([z])
        lib = torchlib.vae_decode(vae, z)
    e = rel(ours, lib)
    print("VAE: product vs library chain relL2 %.3e" % e)
    assert ours.shape == lib.shape == (1, 3, 9, 64, 64) and e < 3e-2
