"""Checkpoint loading with the reference's on-disk layout and error behaviour (inference mode of
sat/training/model_io.py:233-356 `get_checkpoint_iteration` / `get_checkpoint_name` / `load_checkpoint`):

    <load>/latest                                    text file: an iteration number, or the word "release"
    <load>/<iteration | release>/mp_rank_00_model_states.pt      torch.save({"module": state_dict, ...})

Inside the reference's engine the DiT parameters carry the prefix `model.diffusion_model.`; the names after
the prefix are exactly the ones `scail_b200.dit.DiffusionTransformer` registers, so a SCAIL-Preview checkpoint loads
without any key translation.  Host-side only (no kernels)."""
import os
import warnings

import torch

DIT_PREFIX = "model.diffusion_model."


def get_checkpoint_iteration(load_path):
    """model_io.py:233-257: returns (iteration, release)."""
    tracker = os.path.join(load_path, "latest")
    if not os.path.isfile(tracker):
        raise ValueError(f"could not find the metadata file {tracker}, please check --load")
    meta = open(tracker).read().strip()
    try:
        iteration, release = int(meta), False
    except ValueError:
        iteration, release = 0, meta == "release"
        if not release:
            raise ValueError(f"Invalid metadata file {tracker}: {meta!r}")
    if not (iteration > 0 or release):
        raise ValueError(f"error parsing metadata file {tracker}")
    return iteration, release


def get_checkpoint_name(load_path, iteration, release=False, mp_rank=0):
    """model_io.py:36-44 (no ZeRO shards at inference)."""
    return os.path.join(load_path, "release" if release else f"{iteration:d}", f"mp_rank_{mp_rank:02d}_model_states.pt")


def load_checkpoint(model, load_path, prefix=DIT_PREFIX, specific_iteration=None, force_inference=False):
    """model_io.py:260-356, mode == 'inference': load `sd['module']` entries that start with `prefix` (prefix stripped) into
    `model` with strict=False; unexpected keys warn, missing keys raise unless force_inference; model.eval().
    Returns the iteration."""
    iteration, release = get_checkpoint_iteration(load_path)
    if specific_iteration is not None:
        if not (isinstance(specific_iteration, int) and specific_iteration > 0):
            raise ValueError("specific_iteration must be a positive int")
        iteration, release = specific_iteration, False
    name = get_checkpoint_name(load_path, iteration, release)
    sd = torch.load(name, map_location="cpu")
    if "module" not in sd:
        raise ValueError(f"{name} has no 'module' entry (not a SAT checkpoint)")
    module_sd = {k[len(prefix):]: v for k, v in sd["module"].items() if k.startswith(prefix)}
    missing, unexpected = model.load_state_dict(module_sd, strict=False)
    if unexpected:
        warnings.warn(f"Will continue but found unexpected_keys! Check whether you are loading correct checkpoints: {unexpected}.")
    if missing:
        if force_inference:
            warnings.warn(f"Warning: Missing keys for inference: {missing}.")
        else:
            raise ValueError(f"Missing keys for inference: {missing}.\nIf you still want to inference anyway, pass force_inference=True.")
    model.eval()
    return iteration


def save_checkpoint(model, save_path, iteration, prefix=DIT_PREFIX):
    """Writes the layout load_checkpoint reads (model_io.py:199-229, model states only) — used by tests and to re-export
    weights; optimizer / RNG states are training-side and out of scope."""
    d = os.path.join(save_path, f"{iteration:d}")
    os.makedirs(d, exist_ok=True)
    torch.save({"module": {prefix + k: v.detach().cpu() for k, v in model.state_dict().items()}, "iteration": iteration},
               os.path.join(d, "mp_rank_00_model_states.pt"))
    with open(os.path.join(save_path, "latest"), "w") as f:
        f.write(str(iteration))
