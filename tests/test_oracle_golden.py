"""CPU tests: the oracle restatements (oracle/dit_oracle.py, oracle/vae_oracle.py) against the golden vectors
generated from the UNMODIFIED reference by tests/golden/gen_golden.py.  fp32 vs fp32: agreement to rounding."""
import os

import torch

from oracle import dit_oracle as O
from oracle import vae_oracle as V

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def test_dit_forward_matches_reference():
    g = torch.load(os.path.join(GOLD, "dit_c1.pt"))
    sd = {k: v.float() for k, v in g["state_dict"].items()}
    i, cfg = g["inputs"], g["cfg"]
    out, hid = O.dit_forward(sd, i["x"], i["timesteps"], i["context"], i["ref_concat"], i["concat_smpl_render"],
                             i["image_clip_features"], cfg["heads"], cfg["layers"], return_hidden=True)
    assert rel(hid[0], g["tokens"]) < 1e-6
    assert rel(hid[1], g["block_out"]) < 1e-6
    assert rel(out, g["out"]) < 1e-6
    text, clip, emb, adaln = O.embeddings(sd, i["timesteps"], i["context"], i["image_clip_features"].repeat(2, 1, 1))
    assert rel(text, g["text_ctx"]) < 1e-6 and rel(clip, g["clip_ctx"]) < 1e-6
    assert rel(emb, g["time_emb"]) < 1e-6 and rel(adaln, g["adaln_emb"]) < 1e-6


def test_index_maps_bit_exact():
    ix = torch.load(os.path.join(GOLD, "dit_index.pt"))
    t, h, w = ix["geom"]["t"], ix["geom"]["h"], ix["geom"]["w"]
    # patch order: feed integer ids through the oracle's patch_embed with one-hot weights
    def ids(tt, hh, ww, base):
        return (base + torch.arange(tt * hh * ww, dtype=torch.float32)).reshape(1, tt, 1, hh, ww)
    wsel = torch.zeros(4, 1, 1, 2, 2)
    for o in range(4):
        wsel[o, 0, 0, o // 2, o % 2] = 1
    sd = {"mixins.patch_embed.proj.weight": wsel, "mixins.patch_embed.proj.bias": torch.zeros(4),
          "mixins.patch_embed.proj_pose.weight": wsel, "mixins.patch_embed.proj_pose.bias": torch.zeros(4)}
    tok = O.patch_embed(sd, ids(t, h, w, 0), ids(1, h, w, 100000), ids(t, h // 2, w // 2, 200000))
    assert torch.equal(tok.to(torch.int64), ix["tok_ids"])
    n_ref, n_seq, n_pose = O.segment_lengths(t, h, w)
    N = n_ref + n_seq + n_pose
    code = (torch.arange(N)[:, None] * 64 + torch.arange(64)[None]).float()[None]
    assert torch.equal(O.unpatchify(code, n_ref, n_seq, t, h // 2, w // 2).to(torch.int64), ix["unpatchify"])
    cos, sin = O.rope_tables(128, t, h // 2, w // 2, 21, 150, 150)
    # golden pairs were recovered as 0.5*(a+b) / 0.5*(b-a) from the reference's rotated all-ones vector
    assert float((cos[:, 0::2] - ix["rope_cos_pairs"]).abs().max()) < 2e-7
    assert float((sin[:, 0::2] - ix["rope_sin_pairs"]).abs().max()) < 2e-7
    assert torch.equal(cos[:, 0::2], cos[:, 1::2]) and torch.equal(sin[:, 0::2], sin[:, 1::2])
    coords = O.token_coords(t, h, w)
    assert coords.shape == (N, 4) and int(coords[n_ref, 0]) == 1 and int(coords[-1, 0]) == 2


def test_sampler_pieces():
    s = torch.load(os.path.join(GOLD, "sampler.pt"))
    assert torch.equal(O.make_flow_timesteps(50, 5.0), s["sigmas"])
    i = s["i"]
    got = O.cfg_euler_step(s["x"], s["v"][:1], s["v"][1:], s["sigmas"][i], s["sigmas"][i + 1], 4.0)
    assert torch.equal(got, s["x_next"])


def test_vae_decode_matches_reference():
    g = torch.load(os.path.join(GOLD, "vae_small.pt"))
    out = V.decode(g["state_dict"], g["z"])
    assert out.shape == g["out"].shape
    assert rel(out, g["out"]) < 1e-5


def test_vae_encode_matches_reference():
    g = torch.load(os.path.join(GOLD, "vae_encode_small.pt"))
    mu = V.encode(g["state_dict"], g["video"])
    assert mu.shape == g["mu"].shape
    assert rel(mu, g["mu"]) < 1e-5
