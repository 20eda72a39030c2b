"""Audio segmentation with the reference's semantics (upstream amt/src/utils/audio.py
``slice_padded_array`` [RECALL]): cut a waveform into fixed-length segments with hop =
segment length (no overlap), zero-padding the last one."""
from __future__ import annotations

import numpy as np
import torch


def slice_padded_array(x, slice_length: int = 32767, slice_hop: int = 32767, pad: bool = True):
    """x: (1, n_samples) or (n_samples,) numpy / torch -> (n_seg, 1, slice_length) of the same kind.
    The tail shorter than slice_length is zero-padded when ``pad`` (else dropped)."""
    is_torch = isinstance(x, torch.Tensor)
    a = x.reshape(-1)
    n = a.shape[0]
    if pad:
        n_seg = max(1, -(-max(n - slice_length, 0) // slice_hop) + 1)
        total = (n_seg - 1) * slice_hop + slice_length
        if total > n:
            z = torch.zeros(total - n, dtype=a.dtype, device=a.device) if is_torch else np.zeros(total - n, a.dtype)
            a = torch.cat([a, z]) if is_torch else np.concatenate([a, z])
    else:
        n_seg = (n - slice_length) // slice_hop + 1 if n >= slice_length else 0
    segs = [a[i * slice_hop: i * slice_hop + slice_length] for i in range(n_seg)]
    if not segs:
        return (torch.zeros((0, 1, slice_length), dtype=a.dtype) if is_torch
                else np.zeros((0, 1, slice_length), a.dtype))
    out = torch.stack(segs) if is_torch else np.stack(segs)
    return out[:, None, :]
