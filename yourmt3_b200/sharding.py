"""Long-form audio sharded by independent 2.048 s segments across the GPUs of one box
(SURVEY.md 8e).  Segments share nothing, so ranks take contiguous ranges (output stays ordered),
weights are replicated, and exactly ONE collective runs: an all-gather of equally sized, padded
int32 token blocks (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).  The block shape is known
from the model configuration (``token_shape``), so no shape exchange precedes the gather."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_seg: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous ranges of ceil(n_seg / world) segments: returns (start, stop, per_rank)."""
    per = -(-n_seg // world) if n_seg > 0 else 0
    start = min(rank * per, n_seg)
    return start, min(start + per, n_seg), per


@torch.no_grad()
def transcribe_sharded(infer_fn: Callable[[torch.Tensor], torch.Tensor], audio_segments: torch.Tensor, bsz: int,
                       device: torch.device, pad_id: int = 0, group: Optional[dist.ProcessGroup] = None,
                       token_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Every rank holds (or can index) the same ``audio_segments`` (n_seg, 1, L); rank r transcribes
    segments [start_r, stop_r) in batches of ``bsz`` with ``infer_fn`` (-> (b, ...) integer tokens) and all ranks
    return the full ordered token tensor (n_seg, ...) as int32 on ``device``.

    ``token_shape``: per-segment token shape ((L,) or (C, L)), known from the model configuration.  Without it the
    shape is taken from the local output, which requires every rank to own at least one segment."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_seg = audio_segments.shape[0]
    start, stop, per = shard_range(n_seg, world, rank)
    outs = []
    for i in range(start, stop, bsz):
        x = audio_segments[i:min(i + bsz, stop)].to(device, torch.float32, non_blocking=True)
        outs.append(infer_fn(x).to(torch.int32))
    if token_shape is not None:
        tail = tuple(int(v) for v in token_shape)
        if outs and tuple(outs[0].shape[1:]) != tail:
            raise ValueError(f"infer_fn returned per-segment shape {tuple(outs[0].shape[1:])}, token_shape says {tail}")
    elif outs:
        tail = tuple(outs[0].shape[1:])
    else:
        raise ValueError("transcribe_sharded: this rank owns no segment and token_shape was not given "
                         "(the block shape must be known without a collective)")
    block = torch.full((per,) + tail, pad_id, dtype=torch.int32, device=device)
    if outs:
        local = torch.cat(outs, 0)
        block[: local.shape[0]] = local
    if world == 1:
        return block[:n_seg]
    gathered = torch.empty((world * per,) + tail, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, block, group=group)      # the one collective of the path
    return gathered[:n_seg]
