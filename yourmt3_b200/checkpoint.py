"""Checkpoint adapter: a reference checkpoint's state dict -> this package's modules.

Upstream trains with PyTorch Lightning and saves ``{"state_dict": {...}}`` whose keys carry the LightningModule
attribute path (``model.`` prefix [RECALL]) in front of the module keys this package mirrors
(``encoder.block.0.layer.0.SelfAttention.q.weight`` ...).  The adapter strips wrapper prefixes, drops keys that
belong to training only (loss weights, metric buffers), reports what did not line up instead of guessing, and
loads with ``strict`` semantics by default.  Tensors may arrive in fp16/bf16/fp32; parameters are kept fp32 on the
host side (the native handles re-pack them into the compute precision)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Tuple

import torch

WRAPPER_PREFIXES = ("state_dict.", "model.", "module.", "_orig_mod.", "net.")
TRAINING_ONLY = ("loss", "criterion", "metric", "ema_", "optimizer", "lr_scheduler")
# upstream buffers that are recomputed here (never an error when a checkpoint carries them) [RECALL names]:
# persistent sinusoidal / rotary tables, mel filterbank + window copies under other attribute paths, pitch-shift layer
IGNORABLE_BUFFERS = ("inv_freq", "pos_emb.pe", "positional_encoding", "cached_cos", "cached_sin", "pshifters",
                     "spectrogram.mel_stft", "spectrogram.stft", "num_batches_tracked")


def strip_prefixes(state: Mapping[str, torch.Tensor], target_keys: Iterable[str],
                   prefixes: Tuple[str, ...] = WRAPPER_PREFIXES) -> Dict[str, torch.Tensor]:
    """Remove leading wrapper prefixes (repeatedly) from every key until it matches a target key or nothing is left
    to strip.  A key that already matches is left alone, so module names that happen to start with a prefix string
    are not damaged."""
    target = set(target_keys)
    out: Dict[str, torch.Tensor] = {}
    for k, v in state.items():
        name = k
        while name not in target:
            for p in prefixes:
                if name.startswith(p):
                    name = name[len(p):]
                    break
            else:
                break
        if name in out and name != k:
            raise ValueError(f"checkpoint keys collide after prefix stripping: {k!r} -> {name!r}")
        out[name] = v
    return out


def adapt_state_dict(model: torch.nn.Module, checkpoint: Mapping) -> Tuple[Dict[str, torch.Tensor], List[str], List[str]]:
    """-> (state dict restricted to the model's keys, missing keys, unexpected keys).  ``checkpoint`` is either a
    plain state dict or a Lightning checkpoint dict with a ``state_dict`` entry."""
    state = checkpoint["state_dict"] if "state_dict" in checkpoint and isinstance(checkpoint["state_dict"], Mapping) \
        else checkpoint
    own = model.state_dict()
    state = strip_prefixes(state, own.keys())
    adapted, unexpected = {}, []
    for k, v in state.items():
        if k in own:
            if tuple(v.shape) != tuple(own[k].shape):
                raise ValueError(f"checkpoint tensor {k!r} has shape {tuple(v.shape)}, the model expects {tuple(own[k].shape)}")
            adapted[k] = v.detach().to(own[k].dtype)
        elif not any(t in k.lower() for t in TRAINING_ONLY) and not any(t in k for t in IGNORABLE_BUFFERS):
            unexpected.append(k)
    # tied LM head: the reference stores the embedding once
    if "lm_head.lm_head.weight" in own and "lm_head.lm_head.weight" not in adapted and "embed_tokens.weight" in adapted \
            and getattr(model, "tie_word_embeddings", False):
        adapted["lm_head.lm_head.weight"] = adapted["embed_tokens.weight"]
    missing = [k for k in own if k not in adapted]
    return adapted, missing, unexpected


def load_checkpoint(model: torch.nn.Module, checkpoint, strict: bool = True, map_location="cpu",
                    trusted: bool = False) -> Tuple[List[str], List[str]]:
    """Load a reference checkpoint (path or dict) into ``model``; returns (missing, unexpected).  ``strict`` raises
    if anything is missing or unexpected (after dropping training-only entries and known-ignorable buffers).

    Files are read with ``torch.load(weights_only=True)``.  Lightning ``.ckpt`` files usually also pickle
    non-tensor objects (hyper-parameters as argparse / AttributeDict, callbacks); that load then fails and this
    function says so instead of surfacing a bare UnpicklingError.  ``trusted=True`` opts in to a full unpickle -
    only for files whose origin you trust, unpickling executes code."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        path = checkpoint
        try:
            checkpoint = torch.load(path, map_location=map_location, weights_only=not trusted)
        except Exception as e:   # pickle.UnpicklingError and friends
            if trusted:
                raise
            raise RuntimeError(
                f"could not read {path!r} with weights_only=True ({type(e).__name__}: {str(e)[:200]}). Lightning "
                "checkpoints often carry pickled hyper-parameters / callbacks; pass trusted=True to unpickle a file "
                "you trust, or re-save it as {'state_dict': ...} with tensors only.") from e
    adapted, missing, unexpected = adapt_state_dict(model, checkpoint)
    if strict and (missing or unexpected):
        raise KeyError(f"checkpoint does not match the model: missing {missing[:8]}{'...' if len(missing) > 8 else ''}, "
                       f"unexpected {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}")
    model.load_state_dict(adapted, strict=False)
    return missing, unexpected
