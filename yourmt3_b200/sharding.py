"""Long-form audio sharded by independent 2.048 s segments across the GPUs of one box
(SURVEY.md 8e).  Segments share nothing, so ranks take contiguous ranges (output stays ordered),
weights are replicated, and exactly ONE collective runs: an all-gather of equally sized, padded
int32 token blocks (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_seg: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous ranges of ceil(n_seg / world) segments: returns (start, stop, per_rank)."""
    per = -(-n_seg // world) if n_seg > 0 else 0
    start = min(rank * per, n_seg)
    return start, min(start + per, n_seg), per


@torch.no_grad()
def transcribe_sharded(infer_fn: Callable[[torch.Tensor], torch.Tensor], audio_segments: torch.Tensor, bsz: int,
                       device: torch.device, pad_id: int = 0, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Every rank holds (or can index) the same ``audio_segments`` (n_seg, 1, L); rank r transcribes
    segments [start_r, stop_r) in batches of ``bsz`` with ``infer_fn`` (-> (b, ...) integer tokens) and all ranks
    return the full ordered token tensor (n_seg, ...) as int32 on ``device``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_seg = audio_segments.shape[0]
    start, stop, per = shard_range(n_seg, world, rank)
    outs = []
    for i in range(start, stop, bsz):
        x = audio_segments[i:min(i + bsz, stop)].to(device, torch.float32, non_blocking=True)
        outs.append(infer_fn(x).to(torch.int32))
    # token shape: learned from local output, or from any other rank when this shard is empty
    shape = torch.zeros(8, dtype=torch.int64, device=device)
    if outs:
        tail = outs[0].shape[1:]
        shape[0] = len(tail)
        for j, d in enumerate(tail):
            shape[1 + j] = d
    if world > 1:
        dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    tail = tuple(int(v) for v in shape[1:1 + int(shape[0])].tolist())
    block = torch.full((per,) + tail, pad_id, dtype=torch.int32, device=device)
    if outs:
        local = torch.cat(outs, 0)
        block[: local.shape[0]] = local
    if world == 1:
        return block[:n_seg]
    gathered = torch.empty((world * per,) + tail, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, block, group=group)      # the one collective of the path
    return gathered[:n_seg]
