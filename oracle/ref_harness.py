"""TEST INFRASTRUCTURE ONLY — harness that imports the UNMODIFIED reference
(zai-org/SCAIL, mounted read-only at /root/reference) on CPU so that
`tests/golden/gen_golden.py` can generate golden vectors from it.

Nothing in the product (`scail_b200/`) may import this module.  It only works
where /root/reference exists (the build container); the GPU box never runs it.

Recipe follows SURVEY.md §8(c):
  * stub `pytorch_lightning` / `omegaconf` (sgm/__init__.py:1 imports them eagerly),
  * patch torch.cuda.get_device_name (sat/mpu/ulysses_attn_layer.py:36-37) and
    Tensor.cuda (dit_video_crossattn_sc_xc.py:510-513) so the ctor runs on CPU,
  * gloo world-size-1 process group + mpu.initialize_model_parallel(1, 1) before
    model construction (sat/model/base_model.py:88-89).
"""
import argparse
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SCAIL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dit_video_crossattn_sc_xc.py"))


_READY = False


def setup():
    """Make `import dit_video_crossattn_sc_xc` / `sgm.models.wan_vae` work on CPU."""
    global _READY
    if _READY:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            pass

        class LightningDataModule:
            pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules["pytorch_lightning"] = pl
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class ListConfig(list):
            pass

        class DictConfig(dict):
            pass

        class OmegaConf:
            @staticmethod
            def to_container(x, **kw):
                return x

        oc.ListConfig, oc.DictConfig, oc.OmegaConf = ListConfig, DictConfig, OmegaConf
        lc = types.ModuleType("omegaconf.listconfig")
        lc.ListConfig = ListConfig
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc

    if not torch.cuda.is_available():
        torch.cuda.get_device_name = lambda *a, **k: "cpu"
        torch.Tensor.cuda = lambda self, *a, **k: self

    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("gloo", rank=0, world_size=1)
    from sat import mpu

    try:
        mpu.get_model_parallel_world_size()
    except Exception:
        mpu.initialize_model_parallel(1, 1)
    _READY = True


def dit_config(hidden=256, heads=2, inner=512, layers=1, text_dim=64, mixin_module="dit_video_crossattn_sc_xc",
               latent_hw=300, num_frames=81):
    """kwargs for the reference DiffusionTransformer mirroring
    configs/video_model/Wan2.1-i2v-14Bsc-pose-xc-latent.yaml:22-75 at reduced width."""
    targs = argparse.Namespace(
        checkpoint_activations=False, vocab_size=1, max_sequence_length=64,
        layernorm_order="pre", skip_init=False, model_parallel_size=1, is_decoder=True,
    )
    modules = {
        "pos_embed_config": {"target": f"{mixin_module}.Rotary3DPositionEmbeddingMixin",
                             "params": {"hidden_size_head": hidden // heads, "interleaved_rope": True}},
        "patch_embed_config": {"target": f"{mixin_module}.ImagePatchEmbeddingMixin",
                               "params": {"use_conv": True}},
        "adaln_layer_config": {"target": f"{mixin_module}.AdaLNMixin",
                               "params": {"qk_ln": True, "qk_ln_affine": True, "hidden_size_head": hidden}},
        "final_layer_config": {"target": f"{mixin_module}.FinalLayerMixin"},
    }
    return dict(
        transformer_args=targs, time_freq_dim=256, time_embed_dim=hidden, share_adaln=True,
        elementwise_affine=False, num_frames=num_frames, time_compressed_rate=4,
        latent_width=latent_hw, latent_height=latent_hw, num_layers=layers, patch_size=[1, 2, 2],
        in_channels=20, out_channels=16, text_dim=text_dim, hidden_size=hidden,
        inner_hidden_size=inner, num_attention_heads=heads, use_SwiGLU=False, use_RMSNorm=False,
        layernorm_epsilon=1e-6, modules=modules, dtype="fp32", use_i2v_clip=True,
    )


def build_reference_dit(seed=1234, **cfg_kw):
    """Construct the reference DiffusionTransformer (fp32, CPU) with seeded random
    weights whose values are bf16-representable (SURVEY §8c parity metric)."""
    setup()
    import dit_video_crossattn_sc_xc as ref

    torch.manual_seed(seed)
    model = ref.DiffusionTransformer(**dit_config(**cfg_kw)).eval()
    randomize_(model, seed)
    return model


@torch.no_grad()
def randomize_(model, seed):
    """Give every parameter a non-degenerate, bf16-representable value (the
    reference zero-inits several biases / out-projections which would hide bugs)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() >= 2 and "adaLN" not in name:
            fan_in = p[0].numel()
            std = min(0.05, 1.0 / fan_in ** 0.5)
            p.copy_(torch.randn(p.shape, generator=g) * std)
        elif "layernorm" in name and name.endswith("weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        elif "adaLN" in name:
            p.copy_(torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5)
        else:
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
        p.copy_(p.to(torch.bfloat16).float())


def build_reference_vae(seed=7, dim=96):
    setup()
    from sgm.models import wan_vae as ref

    torch.manual_seed(seed)
    vae = ref.WanVAE_(dim=dim, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                      temperal_downsample=[False, True, True]).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in vae.named_parameters():
            if p.dim() >= 2 and p.numel() > p.shape[0]:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif "gamma" in name:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            p.copy_(p.to(torch.bfloat16).float())
    return vae


VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
            0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
           3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]
