"""BASELINE.json configs[4]: 1 hour of synthetic 16 kHz mono audio, YPTF.MoE+Multi, sharded by independent 2.048 s
segments across the GPUs of one box; only the final int32 token all-gather crosses NVLink (NCCL).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      tools/transcribe_hour.py [--minutes 60] [--bsz 256] [--preset yptf_moe_multi]
(or plain `python tools/transcribe_hour.py` for 1 GPU).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import yourmt3_b200 as ymt3  # noqa: E402
from yourmt3_b200.audio_utils import slice_padded_array  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--minutes", type=float, default=60.0)
ap.add_argument("--bsz", type=int, default=256)
ap.add_argument("--preset", default="yptf_moe_multi")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
audio = {"codec": "spec", "hop_length": 300} if args.preset.startswith("yptf") else {}
m = ymt3.init_nondegenerate_(ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**audio), model_cfg=ymt3.get_model_cfg(args.preset),
                                          precision="bf16"), 0).to(dev)
n_samples = int(args.minutes * 60 * 16000)
g = torch.Generator().manual_seed(1234)
wave = torch.randn(n_samples, generator=g) * 0.1            # same waveform on every rank (host memory)
segs = slice_padded_array(wave, 32767, 32767).pin_memory()   # (n_seg, 1, 32767)


def run():
    toks = m.inference_file_sharded(args.bsz, segs, stop_at_eos=True)
    torch.cuda.synchronize()
    return toks


run()   # warm-up (workspace allocation, CUDA graph capture)
times = []
for _ in range(args.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    toks = run()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
t = torch.tensor([min(times)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    secs = n_samples / 16000.0
    print(json.dumps({"workload": "1h_sharded" if args.minutes == 60 else f"{args.minutes}min_sharded", "preset": args.preset,
                      "n_gpus": world, "audio_seconds": secs, "segments": int(segs.shape[0]), "tokens_shape": list(toks.shape),
                      "wall_s": float(t.item()), "audio_s_per_wall_s": secs / float(t.item()), "bsz": args.bsz,
                      "includes": "H2D of this rank's segments from pinned host memory, frontend, encode, decode, NCCL all-gather"}),
          flush=True)
if world > 1:
    dist.destroy_process_group()
