"""TEST INFRASTRUCTURE — a restatement of HOW the reference drives the SAT hooks, so the drop-in mixins can be exercised on
the GPU box where /root/reference (and SAT) is absent.

  DiffusionTransformer.forward   dit_video_crossattn_sc_xc.py:1452-1587   assembles the kwargs (20-channel images/ref/pose
                                                                          with the mask channels appended, emb, lengths, rope_*)
  BaseModel.forward              sat/model/base_model.py:133-138          -> transformer(*args, **kwargs)
  BaseTransformer.forward        sat/model/transformer.py:572-746         word_embedding_forward -> position_embedding_forward
                                                                          -> layer_forward x L (layer_id=torch.tensor(i), position_ids,
                                                                          output_this_layer, output_cross_layer) -> final_forward
                                                                          (parallel_output)
Only the text / CLIP / time embeddings are taken from the product model's own `_embeddings` (in the reference they are
nn.Sequential library modules: not hooks, not part of what this test pins).
"""
from functools import reduce
from operator import mul

import torch


def drive_hooks(model, x, timesteps, context, ref_concat, concat_smpl_render, image_clip_features, chunk_dim=None, sp_rank=0):
    b, t, _, h, w = x.shape
    dt = torch.bfloat16
    x = x.to(dt)
    kwargs = dict(concat_images=x, chunk_dim=chunk_dim)
    # ---- dit_video_crossattn_sc_xc.py:1457-1503 (concat_images present -> mask channels are appended) ----
    mask = torch.zeros(b, t, 4, h, w, device=x.device, dtype=dt)
    ref = ref_concat.repeat(b // ref_concat.shape[0], 1, 1, 1, 1)
    kwargs["ref_concat"] = torch.cat([ref, torch.ones(b, 1, 4, h, w, device=x.device, dtype=dt)], dim=2)
    pose = concat_smpl_render.repeat(b // concat_smpl_render.shape[0], 1, 1, 1, 1)
    kwargs["concat_smpl_render"] = torch.cat([pose, torch.ones(b, t, 4, h // 2, w // 2, device=x.device, dtype=dt)], dim=2)
    x = torch.cat([x, mask], dim=2)
    # ---- :1505-1555 embeddings (library modules in the reference; the product's kernels here) ----
    text, clip, emb, adaln = model._embeddings(timesteps, context, image_clip_features, b)
    kwargs["image_clip_features"] = clip
    kwargs["final_layer_emb"] = emb
    pp = reduce(mul, model.patch_size)
    kwargs["seq_length"] = t * h * w // pp
    kwargs["pose_length"] = t * (h // 2) * (w // 2) // pp
    kwargs["ref_length"] = 1 * h * w // pp
    kwargs["images"] = x
    kwargs["emb"] = adaln
    kwargs["encoder_outputs"] = text
    kwargs["cross_attention_mask"] = torch.ones(context.shape[:2], dtype=x.dtype)
    kwargs["text_length"] = context.shape[1]
    kwargs["rope_T"] = t // model.patch_size[0]
    kwargs["rope_H"] = h // model.patch_size[1]
    kwargs["rope_W"] = w // model.patch_size[2]
    kwargs["global_rope_H"] = 0
    kwargs["global_rope_W"] = 120
    input_ids = position_ids = attention_mask = torch.ones((1, 1)).to(x.dtype)
    kwargs["rope_H_shift"] = 0
    kwargs["rope_W_shift"] = 0
    if chunk_dim is not None:  # :1578-1585
        if chunk_dim == 3:
            kwargs["rope_H_shift"] = sp_rank * (h // model.patch_size[1])
        elif chunk_dim == 4:
            kwargs["rope_W_shift"] = (w // model.patch_size[2]) * sp_rank
        else:
            raise NotImplementedError
    # ---- sat/model/transformer.py:572-746 ----
    hooks = {}
    for name in ("patch_embed", "pos_embed", "adaln_layer", "final_layer"):
        mx = model.mixins[name]
        for hook in ("word_embedding_forward", "position_embedding_forward", "layer_forward", "final_forward"):
            if hasattr(mx, hook):
                assert hook not in hooks, f"hook conflict on {hook}"
                hooks[hook] = getattr(mx, hook)
    attention_mask = attention_mask.type_as(next(model.parameters()))
    output_cross_layer = {}
    hidden_states = hooks["word_embedding_forward"](input_ids, output_cross_layer=output_cross_layer, **kwargs)
    position_embeddings = hooks["position_embedding_forward"](position_ids, output_cross_layer=output_cross_layer, **kwargs)
    assert position_embeddings is None
    for i in range(len(model.transformer.layers)):
        output_this_layer_obj, output_cross_layer_obj = {}, {}
        layer_ret = hooks["layer_forward"](hidden_states, attention_mask, layer_id=torch.tensor(i), **kwargs,
                                           position_ids=position_ids, **output_cross_layer,
                                           output_this_layer=output_this_layer_obj, output_cross_layer=output_cross_layer_obj)
        if isinstance(layer_ret, tuple):
            layer_ret = layer_ret[0]
        hidden_states, output_cross_layer = layer_ret, output_cross_layer_obj
    return hooks["final_forward"](hidden_states, **kwargs, parallel_output=True)
