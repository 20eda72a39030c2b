"""Model workloads for bench.py (kept in the package so bench.py stays a thin driver).

A step = ``YourMT3.inference`` on one batch of synthetic 2.048 s segments resident in HBM
(frontend -> encoder -> device-resident greedy decode to max length; random-init weights never
emit EOS, so every row decodes the full ``event_length`` -- the worst case).
The e2e leg goes through ``inference_file`` with PINNED HOST audio and reads the tokens back.
"""
from __future__ import annotations

import time

SEG_SAMPLES = 32767

PRESETS = {
    # name: (model preset, audio overrides, default batch, precision)
    "t5_small": ("mt3_t5_small", {}, 512, "bf16"),   # 626x at 256, 834x at 512 (profiles/r01_ab_decode_attn_split_length.txt)
    "t5_small_f32": ("mt3_t5_small", {}, 64, "f32"),
    "yptf": ("yptf", {"codec": "spec", "hop_length": 300}, 64, "bf16"),
    # batch: 728 segments x 13 channels = 9464 rows = 74 M-tiles of 128, so the tile counts of the decode-step GEMMs
    # are (near-)multiples of the 148 SMs; measured 1024x at 364, 1041x at 512, 1072x at 728, 1062x at 768, 1080x at
    # 1024 per GPU (profiles/r01_sweep_batch.txt) -- 728 keeps one step under 1.5 s
    "yptf_moe_multi": ("yptf_moe_multi", {"codec": "spec", "hop_length": 300}, 728, "bf16"),
}
DEFAULT = "yptf_moe_multi"   # the model BASELINE.json quotes the target on


class ModelWorkload:
    roofline_bound = "hbm"
    dominant_kernel = "decode_attn_kernel"
    # dram__bytes_read + dram__bytes_write of ONE decode_attn_kernel launch at cache length 128 for 3328 x 6 (sequence,
    # head) pairs, from `ncu --set full` (profiles/r01_cross_absorbed_v3_and_decode_attn_len128_ncu_full.txt, launch 2:
    # 657.4 MB read + 11.3 MB written; algorithmic 669.6 MB)
    NCU_DECODE_ATTN_TRAFFIC_3328x6_L128 = 657.416704e6 + 11.324672e6

    def __init__(self, name, batch):
        self.name = name
        self.preset, self.audio_over, dbatch, self.precision = PRESETS[name]
        self.batch = batch or dbatch
        self.dtype = self.precision

    def setup_native(self, dev):
        import torch
        import yourmt3_b200 as ymt3
        self.torch = torch
        self.model = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**self.audio_over),
                                  model_cfg=ymt3.get_model_cfg(self.preset), precision=self.precision)
        ymt3.init_nondegenerate_(self.model, seed=0)
        self.model = self.model.to(dev)
        self.model.decode_lanes = int(__import__("os").environ.get("YMT3_LANES", "1"))
        g = torch.Generator().manual_seed(1234 + (dev.index or 0))
        self.host_in = (torch.randn(self.batch, 1, SEG_SAMPLES, generator=g) * 0.1).pin_memory()
        self.dev_in = self.host_in.to(dev)
        self.e2e_batch = self.batch
        self.max_len = self.model.max_token_length
        self.channels = getattr(self.model.decoder, "num_channels", 1)
        self.launches_per_step = self._count_launches()
        self.bytes_per_seg = SEG_SAMPLES * 4 + self.model.feat_length * self.model.feat_dim * 4
        # frontend kernel timed alone (CUDA events) for the roofline object
        self._frontend_ms = None

    def _count_launches(self):
        dec_layers = self.model.model_cfg["decoder"][self.model.decoder_type]["num_layers"]
        # per layer: qkv, self-attention, o, cross q, cross-attention, cross o, wi, wo (+ 3 RMSNorm kernels on the fp32
        # path; bf16 fuses them into the GEMMs); per step: embed, (final norm,) lm head with the fused arg-max,
        # select + advance
        per_step = dec_layers * (8 if self.precision == "bf16" else 11) + (3 if self.precision == "bf16" else 4)
        return per_step * self.max_len + 64          # + frontend/encoder launches (lower bound)

    def step(self):
        toks = self.model.inference(self.dev_in, stop_at_eos=True)
        dist = self.torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # the one collective of the sharded path: all-gather of the int32 token blocks over NCCL
            t32 = toks.to(self.torch.int32).contiguous()
            out = self.torch.empty((dist.get_world_size() * t32.shape[0],) + tuple(t32.shape[1:]), dtype=t32.dtype,
                                   device=t32.device)
            dist.all_gather_into_tensor(out, t32)
            return out
        return toks

    def step_e2e(self):
        outs = self.model.inference_file(self.batch, self.host_in, stop_at_eos=True)
        return outs

    def e2e_bytes(self):
        return self.batch * SEG_SAMPLES * 4, self.batch * self.channels * self.max_len * 8

    def rooflines(self, peaks):
        """{"roofline": dominant kernel of the step, "roofline_frontend": the log-mel kernel (metric (ii))}.

        The dominant kernel of every model workload is ``decode_attn_kernel`` (self-attention of one decoder layer
        over the device-resident KV cache; 8 launches per step, ~40-50 % of the step, see tools/time_phases.py).  It is
        timed ALONE here through the C ABI (``ymt3_op_decode_attention``) at the workload's own shape
        (N = batch x channels sequences, 6 heads x 64, bf16/f32 cache) over a uniform sample of the cache lengths the
        decode loop visits (8, 24, ..., 248 of 256 -> mean 128), CUDA events, median of 5 per length.  Algorithmic
        bytes per launch = N*H*64*es*(2*len + 6): K and V rows [0, len) read once, q / new k / new v read, new k / v
        appended, out written."""
        import ctypes  # noqa: F401
        from . import _lib
        torch = self.torch
        lib = _lib.load()
        dev = self.dev_in.device
        cfg = self.model.model_cfg["decoder"][self.model.decoder_type]
        H, dk = cfg["num_heads"], cfg.get("d_kv", 64)
        N, Lcap = self.batch * self.channels, self.max_len
        td = _lib.torch_dtype(self.model._prec)
        es = 2 if td == torch.bfloat16 else 4
        q = torch.randn(N, H * dk, device=dev).to(td)
        kn, vn, out = torch.randn_like(q), torch.randn_like(q), torch.empty_like(q)
        Kc = torch.randn(N, H, Lcap, dk, device=dev).to(td)
        Vc = torch.randn(N, H, Lcap, dk, device=dev).to(td)
        step = torch.zeros(1, dtype=torch.int32, device=dev)
        s = _lib.current_stream_ptr()
        tot_ms, tot_bytes, n = 0.0, 0.0, 0
        for ln in range(8, Lcap + 1, 16):
            step.fill_(ln - 1)
            ts = []
            for i in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(lib.ymt3_op_decode_attention(self.model._prec, q.data_ptr(), kn.data_ptr(), vn.data_ptr(),
                                                        Kc.data_ptr(), Vc.data_ptr(), step.data_ptr(), 0, out.data_ptr(),
                                                        N, H, Lcap, s), "decode_attention")
                b.record()
                b.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(b))
            ts.sort()
            tot_ms += ts[len(ts) // 2]
            tot_bytes += N * H * dk * es * (2 * ln + 6)
            n += 1
        kern_ms, alg = tot_ms / n, tot_bytes / n
        achieved = alg / (kern_ms * 1e-3) / 1e9
        traffic = self.NCU_DECODE_ATTN_TRAFFIC_3328x6_L128
        if traffic is not None:
            traffic = int(traffic * (N * H) / (3328.0 * 6) * (Lcap / 256.0) * (es / 2.0))
        main = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
                "kernel": "decode_attn_kernel", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg,
                "launches_per_step": self.max_len * cfg["num_layers"],
                "note": "self-attention over the KV cache, timed alone at this workload's shape, mean over cache lengths 8..%d" % Lcap}
        f_ms = self.dominant_kernel_ms()
        f_ach = self.roofline_units() / (f_ms * 1e-3) / 1e9
        front = {"bound": "hbm", "achieved": f_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": f_ach / peaks["hbm_gbs"],
                 "traffic": self.ncu_traffic_bytes, "peak_source": peaks["source"], "kernel": "ymt3_logmel_kernel",
                 "kernel_ms": f_ms, "algorithmic_bytes_per_launch": self.roofline_units(), "launches_per_step": 1,
                 "note": "compute-bound on the fp32 pipes (2048-point FFT per 128/300 new samples), see DESIGN.md 3.1"}
        return {"roofline": main, "roofline_frontend": front}

    def dominant_kernel_ms(self):
        """log-mel kernel of this workload timed alone with CUDA events (median of 10)."""
        torch = self.torch
        x = self.dev_in
        ts = []
        for _ in range(3):
            self.model.spectrogram(x)
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.model.spectrogram(x)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    @property
    def ncu_traffic_bytes(self):
        """dram__bytes_read + dram__bytes_write of the log-mel kernel from `ncu --set full`, scaled to this batch:
        spec/hop-300 codec: profiles/r01_logmel_default_workload_ncu_full.txt (33.7 MB + 60.2 MB per 256 segments;
        below the algorithmic 148.9 MB because part of the output is still resident in the 126 MB L2 when the
        kernel ends); melspec/hop-128: profiles/r01_logmel_v1_ncu_full.txt (67.2 MB + 213.7 MB per 512 segments)."""
        if self.model.audio_cfg["codec"] == "spec":
            return int((33.69e6 + 60.15e6) / 256 * self.batch)
        return int((67.20e6 + 213.67e6) / 512 * self.batch)

    def roofline_units(self):
        return self.batch * self.bytes_per_seg

    def config(self):
        m = getattr(self, "model", None) or self.ref_model
        self.channels = getattr(m.decoder, "num_channels", 1)
        self.max_len = m.max_token_length
        return {"workload": self.name, "preset": self.preset, "segments_per_step_per_gpu": self.batch,
                "segment_samples": SEG_SAMPLES, "codec": m.audio_cfg["codec"], "hop": m.audio_cfg["hop_length"],
                "encoder": m.encoder_type, "decoder": m.decoder_type, "channels": self.channels,
                "decode_steps": self.max_len, "vocab": m.vocab_size, "weights": "random-init (non-degenerate, seed 0)",
                "l2_policy": "per-step working set (activations + KV cache) larger than L2",
                "cross_attention": "absorbed (latent)" if (getattr(self, "model", None) is not None
                                                           and self.model._absorbed()) else "k/v",
                "roofline_kernel_note": "roofline = decode self-attention kernel (dominant); roofline_frontend = log-mel"}

    # ---- reference arm: same architecture through the CPU oracle (HF-pinned torch restatement) ----
    def setup_reference(self):
        import numpy as np
        import torch
        import yourmt3_b200 as ymt3
        from oracle import pipeline as OP
        self.OP = OP
        self.ref_batch = 1 if self.preset != "mt3_t5_small" else 2
        m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**self.audio_over), model_cfg=ymt3.get_model_cfg(self.preset),
                         precision="f32")
        ymt3.init_nondegenerate_(m, seed=0)
        self.ref_model = m
        self.ref_sd = {k: v.detach() for k, v in m.state_dict().items()}
        g = torch.Generator().manual_seed(1234)
        self.ref_audio = (torch.randn(self.ref_batch, SEG_SAMPLES, generator=g) * 0.1).numpy().astype(np.float32)
        # bounded sample: decode fewer steps on the CPU and extrapolate linearly in the step count
        self.ref_steps = min(64, m.max_token_length)

    def step_reference(self):
        """One bounded CPU sample: the full frontend + encoder + ONE decode step, then the same with ref_steps decode
        steps; the per-step decode cost is their difference and the full-length time is extrapolated linearly in the
        step count (the CPU attention cost grows with the cache length, so this is an UPPER bound on the CPU speed)."""
        m = self.ref_model
        kw = dict(n_pos=m.decoder.pos_table.shape[0], stop_at_eos=False)
        t0 = time.perf_counter()
        self.OP.transcribe(self.ref_sd, self.ref_audio, m.audio_cfg, m.model_cfg, max_length=1, **kw)
        t1 = time.perf_counter()
        out = self.OP.transcribe(self.ref_sd, self.ref_audio, m.audio_cfg, m.model_cfg, max_length=self.ref_steps, **kw)
        t2 = time.perf_counter()
        per_step = max((t2 - t1) - (t1 - t0), 0.0) / max(self.ref_steps - 1, 1)
        self._ref_spent = getattr(self, "_ref_spent", 0.0) + (t2 - t0)
        self._ref_full = getattr(self, "_ref_full", 0.0) + (t1 - t0) + per_step * (m.max_token_length - 1)
        return out

    def reference_scale(self):
        """bench.py divides the audio seconds of the sample by the time SPENT in step_reference; multiplying by
        spent / estimated-full-length time turns that into audio seconds per full-length wall second."""
        return self._ref_spent / self._ref_full if getattr(self, "_ref_full", 0.0) > 0 else 1.0

    def reference_sample(self):
        return (f"CPU oracle (HF-T5-pinned torch restatement) of {self.preset}, {self.ref_batch} segment(s): frontend + "
                f"encoder + 1 decode step, then {self.ref_steps} of {self.ref_model.max_token_length} decode steps; "
                f"full-length time = first + per-step cost x remaining steps")


def get(name, batch):
    if name not in PRESETS:
        raise SystemExit(f"unknown workload {name!r}; choose from frontend, {', '.join(PRESETS)}")
    return ModelWorkload(name, batch)
