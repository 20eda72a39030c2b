#!/usr/bin/env python
"""Benchmark of the YourMT3 inference hot path on B200 (driver contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl native|reference]

A "step" is one pass of the hot path over one batch of synthetic 16 kHz mono segments
(2.048 s each).  Metric: audio-seconds transcribed per wall-second (whole job, all ranks).
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU path
(installed torchaudio / transformers modules = what upstream calls) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEG_SAMPLES = 32767
SEG_SECONDS = SEG_SAMPLES / 16000.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
class FrontendWorkload:
    """BASELINE.json configs[1]: log-mel frontend (n_fft 2048, hop 128, 512 mels) on a big batch."""
    name = "logmel_frontend_hop128_mel512"
    dtype = "f32"

    def __init__(self, batch):
        self.batch = batch or 4096
        self.bytes_per_seg = SEG_SAMPLES * 4 + 256 * 512 * 4     # algorithmic: read audio + write log-mel

    def setup_native(self, dev):
        import torch
        import yourmt3_b200 as ymt3
        self.layer, _ = ymt3.get_spectrogram_layer_from_audio_cfg(ymt3.get_audio_cfg())
        g = torch.Generator().manual_seed(1234 + dev.index)
        self.host_in = (torch.randn(self.batch, SEG_SAMPLES, generator=g) * 0.1).pin_memory()
        self.dev_in = self.host_in.to(dev)
        self.layer(self.dev_in[:2])
        self.e2e_batch = min(self.batch, 512)

    def step(self):
        return self.layer(self.dev_in)          # 2.7 GB in+out per step  >> 126 MB L2

    def step_e2e(self):
        return self.layer.forward_host(self.host_in[: self.e2e_batch])

    def e2e_bytes(self):
        return self.e2e_batch * SEG_SAMPLES * 4, self.e2e_batch * 256 * 512 * 4

    launches_per_step = 1
    roofline_bound = "hbm"

    @property
    def ncu_traffic_bytes(self):
        # dram__bytes_read + dram__bytes_write of ymt3_logmel_kernel from `ncu --set full`
        # (profiles/r02_logmel_v7_mel_b512_ncu_full.txt: 67.2 MB + 211.8 MB per launch of 512 segments), scaled to this batch
        return int((67.21e6 + 211.82e6) / 512 * self.batch)

    def roofline_units(self):
        return self.batch * self.bytes_per_seg      # algorithmic bytes per launch

    def config(self):
        return {"workload": self.name, "segments_per_step_per_gpu": self.batch, "segment_samples": SEG_SAMPLES,
                "n_fft": 2048, "hop": 128, "n_mels": 512, "l2_policy": "inputs+outputs (2.7 GB/step) larger than L2",
                "e2e_segments_per_step": getattr(self, "e2e_batch", None)}

    # reference arm: the installed dependency the upstream wrapper calls, on CPU
    def setup_reference(self):
        import torch
        import torchaudio
        self.ref_batch = 64
        self.ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                                                       f_max=8000.0, n_mels=512, power=1.0)
        g = torch.Generator().manual_seed(1234)
        self.ref_in = torch.randn(self.ref_batch, SEG_SAMPLES, generator=g) * 0.1

    def step_reference(self):
        import torch
        with torch.no_grad():
            return torch.log(torch.clamp(self.ms(self.ref_in), min=1e-5)).transpose(1, 2).contiguous()

    def reference_sample(self):
        return f"torchaudio MelSpectrogram+log on CPU, {self.ref_batch} segments per step"


def get_workload(name, batch):
    if name in ("frontend", "logmel"):
        return FrontendWorkload(batch)
    try:
        from yourmt3_b200 import bench_workloads
    except ImportError:
        bench_workloads = None
    if bench_workloads is not None:
        return bench_workloads.get(name, batch)
    raise SystemExit(f"unknown workload {name}")


DEFAULT_WORKLOAD = "frontend"


def default_workload():
    try:
        from yourmt3_b200 import bench_workloads
        return bench_workloads.DEFAULT
    except ImportError:
        return DEFAULT_WORKLOAD


# ----------------------------------------------------------------------------------------------
def cpu_baseline(wl, budget_s=12.0):
    import torch
    wl.setup_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    wl.step_reference()
    t0 = time.perf_counter()
    n = 0
    while True:
        wl.step_reference()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    scale = wl.reference_scale() if hasattr(wl, "reference_scale") else 1.0
    return {"value": wl.ref_batch * SEG_SECONDS * n / dt * scale, "unit": "audio-s/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": wl.reference_sample() + f", {n} steps in {dt:.1f}s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = get_workload(args.workload or default_workload(), args.batch)
    wl.setup_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(1, min(args.warmup, 2))):
        wl.step_reference()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.step_reference()
    dt = time.perf_counter() - t0
    val = wl.ref_batch * SEG_SECONDS * args.steps / dt * (wl.reference_scale() if hasattr(wl, "reference_scale") else 1.0)
    line = {"metric": "audio_seconds_per_wall_second", "value": val, "unit": "audio-s/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": torch.get_num_threads(),
                             "kind": "port", "sample": wl.reference_sample()},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_native(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = get_workload(args.workload or default_workload(), args.batch)
    wl.setup_native(dev)
    for _ in range(max(3, args.warmup)):
        wl.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        wl.step()
        ev[i + 1].record()
    barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())

    # end-to-end through the public host-buffer API (H2D + kernels + D2H inside the timed region)
    for _ in range(2):
        wl.step_e2e()
    barrier()
    e0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        wl.step_e2e()
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - e0
    et = torch.tensor([e_dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e_dt = float(et.item())
    h2d, d2h = wl.e2e_bytes()

    if rank == 0:
        peaks = measured_peaks()
        units = wl.units(world) if hasattr(wl, "units") else wl.batch * world
        value = units * SEG_SECONDS * args.steps / (total_ms * 1e-3)
        if hasattr(wl, "rooflines"):
            roofs = wl.rooflines(peaks)     # model workloads: dominant decode kernel + the log-mel frontend kernel
        else:
            # frontend workload: the step IS one launch of the log-mel kernel
            kern_ms = sorted(per_step)[len(per_step) // 2]
            achieved = wl.roofline_units() / (kern_ms * 1e-3) / 1e9
            roofs = {"roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": achieved / peaks["hbm_gbs"], "traffic": getattr(wl, "ncu_traffic_bytes", None),
                                  "peak_source": peaks["source"], "kernel": "ymt3_logmel_kernel", "kernel_ms": kern_ms}}
            try:
                from yourmt3_b200.bench_workloads import frontend_compute_bound
                roofs["roofline"]["compute"] = frontend_compute_bound(wl.batch * 256, kern_ms)
            except ImportError:
                pass
        line = {
            "metric": "audio_seconds_per_wall_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": wl.config(),
            **roofs,
            "e2e": {"value": (units if getattr(wl, "scaling", "weak") == "strong" else wl.e2e_batch * world)
                    * SEG_SECONDS * n_e2e / e_dt, "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": wl.launches_per_step * args.steps,
            "gpu_launches_note": "estimated from the structure of the step (kernels per decode step x steps + encoder), "
                                 "not counted by a profiler",
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        if world == 1 and not args.no_gpu_eager_baseline and hasattr(wl, "gpu_eager_baseline"):
            line["gpu_eager_baseline"] = wl.gpu_eager_baseline(dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--batch", type=int, default=None, help="segments per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true",
                    help="skip the eager-torch-on-the-same-GPU bar (model workloads, N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
