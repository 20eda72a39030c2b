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
    scaling = "weak"
    roofline_bound = "hbm"
    dominant_kernel = "decode_attn_kernel"
    # dram__bytes_read + dram__bytes_write of ONE decode_attn_kernel launch AT THE BENCH SHAPE (9464 x 6 (sequence, head)
    # pairs = 728 segments x 13 channels, bf16 cache, length 128 = the mean of a 256-step decode), from `ncu --set full`
    # (profiles/r02_decode_attn_9464x6_len128_ncu_full.txt: 1868.5 MB read + 52.5 MB written; algorithmic 1904.3 MB)
    NCU_DECODE_ATTN_TRAFFIC_9464x6_L128 = 1868.480e6 + 52.546304e6

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

    def units(self, world):
        """segments transcribed per step by the whole job"""
        return self.batch * world

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
        traffic = self.NCU_DECODE_ATTN_TRAFFIC_9464x6_L128
        if traffic is not None:   # measured at 9464 x 6 / length 128 / bf16: scale factor 1 for the default workload
            traffic = int(traffic * (N * H) / (9464.0 * 6) * (Lcap / 256.0) * (es / 2.0))
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
                 "note": "issue / fp32-pipe bound (2048-point FFT per 128/300 new samples), see DESIGN.md 3.1 and the "
                         "`compute` object beside this one"}
        front["compute"] = frontend_compute_bound(self.batch * self.model.feat_length, f_ms)
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
        spec/hop-300 codec: profiles/r02_logmel_v7_spec_b728_ncu_full.txt (102.3 MB + 271.3 MB for 728 segments, captured
        at the default batch; below the algorithmic 423.4 MB because part of the output is still resident in the
        126 MB L2 when the kernel ends); melspec/hop-128: profiles/r02_logmel_v7_mel_b512_ncu_full.txt (67.2 MB + 211.8 MB
        per 512 segments)."""
        if self.model.audio_cfg["codec"] == "spec":
            return int((102.326e6 + 271.327e6) / 728 * self.batch)
        return int((67.21e6 + 211.82e6) / 512 * self.batch)

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
                "cross_attention": "absorbed (latent)" if (self.precision == "bf16" and m.decoder_type == "multi-t5"
                                                           and m.absorb_cross_attention) else "k/v",
                "roofline_kernel_note": "roofline = decode self-attention kernel (dominant); roofline_frontend = log-mel"}

    # ---- reference arm: same architecture on the host cores: torchaudio frontend (the library upstream wraps) +
    # the HF-pinned torch restatement of the model (oracle/), B = 8 segments ----
    REF_BATCH = 8

    def setup_reference(self):
        import torch
        import yourmt3_b200 as ymt3
        from oracle import pipeline as OP
        self.OP = OP
        self.ref_batch = self.REF_BATCH
        m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**self.audio_over), model_cfg=ymt3.get_model_cfg(self.preset),
                         precision="f32")
        ymt3.init_nondegenerate_(m, seed=0)
        self.ref_model = m
        self.ref_sd = {k: v.detach() for k, v in m.state_dict().items()}
        self.ref_frontend = reference_frontend(m.audio_cfg)
        g = torch.Generator().manual_seed(1234)
        self.ref_audio = torch.randn(self.ref_batch, SEG_SAMPLES, generator=g) * 0.1
        # bounded sample: decode fewer steps on the CPU and extrapolate linearly in the step count
        self.ref_steps = min(64, m.max_token_length)

    def step_reference(self):
        """One bounded CPU sample: frontend (torchaudio) + encoder ONCE, then a 1-step decode and a ref_steps-step
        decode on the same encoder states; the per-step decode cost is their difference and the full-length time is
        frontend + encoder + first step + per-step cost x remaining steps (the CPU attention cost grows with the
        cache length, so this is an UPPER bound on the CPU speed)."""
        import torch
        m, OP = self.ref_model, self.OP
        n_pos = m.decoder.pos_table.shape[0]
        with torch.no_grad():
            t0 = time.perf_counter()
            feats = self.ref_frontend(self.ref_audio)
            if m.encoder_type == "t5":
                enc = OP.t5_encode(self.ref_sd, feats, m.model_cfg, n_pos)
            else:
                from oracle import perceiver_tf as OPTF
                enc = OPTF.encode(self.ref_sd, feats, m.model_cfg)
            t1 = time.perf_counter()
            OP.t5_generate(self.ref_sd, enc, m.model_cfg, n_pos, 1, stop_at_eos=False)
            t2 = time.perf_counter()
            out = OP.t5_generate(self.ref_sd, enc, m.model_cfg, n_pos, self.ref_steps, stop_at_eos=False)
            t3 = time.perf_counter()
        per_step = max((t3 - t2) - (t2 - t1), 0.0) / max(self.ref_steps - 1, 1)
        self._ref_spent = getattr(self, "_ref_spent", 0.0) + (t3 - t0)
        self._ref_full = getattr(self, "_ref_full", 0.0) + (t2 - t0) + per_step * (m.max_token_length - 1)
        return out

    def reference_scale(self):
        """bench.py divides the audio seconds of the sample by the time SPENT in step_reference; multiplying by
        spent / estimated-full-length time turns that into audio seconds per full-length wall second."""
        return self._ref_spent / self._ref_full if getattr(self, "_ref_full", 0.0) > 0 else 1.0

    def reference_sample(self):
        return (f"torchaudio frontend + CPU oracle (HF-pinned torch restatement) of {self.preset}, {self.ref_batch} "
                f"segments per step: frontend + encoder once, a 1-step and a {self.ref_steps}-step decode "
                f"(of {self.ref_model.max_token_length}); full-length time = frontend + encoder + first step + per-step "
                f"cost x remaining steps")

    # ---- the "bar to beat" of SURVEY.md 2c: the SAME eager torch modules on the SAME GPU (cuFFT / cuBLAS / ATen),
    # bf16 autocast, full decode length, host loop with the reference's per-step structure ----
    def gpu_eager_baseline(self, dev, segments=None, budget_s=60.0):
        import torch
        import yourmt3_b200 as ymt3
        from oracle import pipeline as OP
        B = int(segments or min(self.batch, 64 if self.preset == "mt3_t5_small" else self.batch))
        m = ymt3.YourMT3(audio_cfg=ymt3.get_audio_cfg(**self.audio_over), model_cfg=ymt3.get_model_cfg(self.preset),
                         precision="f32")
        ymt3.init_nondegenerate_(m, seed=0)
        sd = {k: v.detach().to(dev) for k, v in m.state_dict().items()}
        fe = reference_frontend(m.audio_cfg).to(dev)
        x = self.dev_in[:B, 0] if getattr(self, "dev_in", None) is not None and self.dev_in.shape[0] >= B \
            else torch.randn(B, SEG_SAMPLES, device=dev) * 0.1
        n_pos, L = m.decoder.pos_table.shape[0], m.max_token_length
        enc_chunk = 64     # the eager conv pre-encoder materialises (b, 64..128, T, F) fp32 activations
        old = OP.DEVICE
        OP.DEVICE = str(dev)

        def run(steps):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                if m.encoder_type == "t5":
                    enc = torch.cat([OP.t5_encode(sd, fe(x[i:i + enc_chunk]), m.model_cfg, n_pos)
                                     for i in range(0, B, enc_chunk)], 0)
                else:
                    from oracle import perceiver_tf as OPTF
                    enc = torch.cat([OPTF.encode(sd, fe(x[i:i + enc_chunk]), m.model_cfg) for i in range(0, B, enc_chunk)], 0)
                tok = OP.t5_generate(sd, enc.float(), m.model_cfg, n_pos, steps, stop_at_eos=True)
            torch.cuda.synchronize()
            return tok
        try:
            run(2)                                         # warm-up (cuBLAS / cuDNN heuristics, allocator)
            t0 = time.perf_counter()
            run(L)
            dt = time.perf_counter() - t0
            out = {"value": B * SEG_SAMPLES / 16000.0 / dt, "unit": "audio-s/s", "segments": B, "decode_steps": L,
                   "seconds": dt, "dtype": "bf16 autocast (fp32 softmax / norms as in the reference)",
                   "what": "torchaudio frontend + eager torch modules of the oracle on the SAME GPU (cuFFT / cuBLAS / "
                           "cuDNN / ATen), host greedy loop with the reference's per-step EOS sync"}
        except Exception as e:   # OOM etc.: report, never fail the bench
            out = {"value": None, "error": f"{type(e).__name__}: {str(e)[:160]}", "segments": B}
        finally:
            OP.DEVICE = old
            del sd
            torch.cuda.empty_cache()
        return out


def frontend_compute_bound(frames, kernel_ms, n_fft=2048, sm_count=148, sm_mhz=1965.0):
    """the honest bound of the log-mel kernel beside the HBM one.  Algorithmic fp32 flops per frame of a
    real n_fft-point transform (split-radix-class count 2.5 N log2 N for the real-input FFT) + window (N) + power
    spectrum (3 per bin); peak = SMs x 128 fp32 lanes x 2 (FMA) x the maximum SM clock - an upper bound no SIMT FFT
    reaches because butterflies are add/sub-heavy (at most half of the issue slots are FMAs)."""
    import math
    per_frame = 2.5 * n_fft * math.log2(n_fft) + n_fft + 3 * (n_fft // 2 + 1)
    peak = sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    ach = frames * per_frame / (kernel_ms * 1e-3) / 1e12
    return {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "flops_per_frame": per_frame, "peak_source": f"{sm_count} SMs x 128 lanes x 2 x {sm_mhz:.0f} MHz"}


def reference_frontend(audio_cfg):
    """The reference frontend itself: the torchaudio transform upstream's spectrogram.py wraps + log(clamp)
    (SP/torchaudio/transforms/_transforms.py:621-631 / :101-123), as an nn.Module: (B, L) -> (B, T, F)."""
    import torch
    import torchaudio

    class Ref(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.eps = float(audio_cfg.get("log_eps", 1e-5))
            if audio_cfg["codec"] == "melspec":
                self.t = torchaudio.transforms.MelSpectrogram(
                    sample_rate=audio_cfg["sample_rate"], n_fft=audio_cfg["n_fft"], hop_length=audio_cfg["hop_length"],
                    f_min=audio_cfg["f_min"], f_max=audio_cfg["f_max"], n_mels=audio_cfg["n_mels"],
                    power=audio_cfg.get("power", 1.0))
                self.lo, self.hi = 0, None
            else:
                self.t = torchaudio.transforms.Spectrogram(n_fft=audio_cfg["n_fft"], hop_length=audio_cfg["hop_length"],
                                                           power=audio_cfg.get("power", 1.0))
                self.lo = 1 if audio_cfg.get("spec_drop_dc", True) else 0
                self.hi = audio_cfg["n_fft"] // 2 + 1

        def forward(self, x):
            with torch.no_grad():
                y = self.t(x.float())[:, self.lo:self.hi]
                return torch.log(torch.clamp(y, min=self.eps)).transpose(1, 2).contiguous()
    return Ref()


class HourWorkload(ModelWorkload):
    """BASELINE.json configs[4]: ONE HOUR of synthetic 16 kHz audio (1758 segments of 2.048 s), YPTF.MoE+Multi,
    sharded by contiguous segment ranges over the ranks = STRONG scaling (total work fixed).  A step transcribes the
    whole hour: every rank runs its shard in near-equal batches, then the one NCCL all-gather of the int32 tokens."""
    scaling = "strong"
    TOTAL_SEGMENTS = 1758

    def __init__(self, name, batch):
        super().__init__("yptf_moe_multi", batch)
        self.name = "hour_sharded"
        self.max_batch = batch or PRESETS["yptf_moe_multi"][2]

    def setup_native(self, dev):
        import torch
        from .sharding import shard_range
        dist = torch.distributed
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        start, stop, _ = shard_range(self.TOTAL_SEGMENTS, self.world, self.rank)
        n_local = stop - start
        n_batches = max(1, -(-n_local // self.max_batch))
        self.batch = -(-n_local // n_batches)              # near-equal batches (no ragged tail batch)
        super().setup_native(dev)
        g = torch.Generator().manual_seed(1234)
        # the same hour of audio on every rank (host, pinned); the rank's shard also resident on the device
        self.host_full = (torch.randn(self.TOTAL_SEGMENTS, 1, SEG_SAMPLES, generator=g) * 0.1).pin_memory()
        self.dev_full_shard = self.host_full[start:stop].to(dev)
        self.start, self.stop = start, stop
        self.dev = dev
        self.e2e_batch = self.batch

    def units(self, world):
        return self.TOTAL_SEGMENTS

    def _shard_view(self):
        # transcribe_sharded indexes [start, stop) of a full-length tensor: give it a device-resident view of the shard
        class _View:
            def __init__(v, t, start, n):
                v.t, v.start, v.shape = t, start, (n,) + tuple(t.shape[1:])

            def __getitem__(v, sl):
                return v.t[sl.start - v.start: sl.stop - v.start]
        return _View(self.dev_full_shard, self.start, self.TOTAL_SEGMENTS)

    def step(self):
        return self.model.inference_file_sharded(self.batch, self._shard_view(), stop_at_eos=True)

    def step_e2e(self):
        return self.model.inference_file_sharded(self.batch, self.host_full, stop_at_eos=True).cpu()

    def e2e_bytes(self):
        n_local = self.stop - self.start
        return n_local * SEG_SAMPLES * 4, self.TOTAL_SEGMENTS * self.channels * self.max_len * 4

    def config(self):
        c = super().config()
        c.update(workload=self.name, total_segments=self.TOTAL_SEGMENTS, audio_seconds=self.TOTAL_SEGMENTS * SEG_SAMPLES / 16000.0,
                 segments_per_batch=self.batch, sharding="contiguous segment ranges, replicated weights, one int32 all-gather")
        return c


def get(name, batch):
    if name in ("hour", "hour_sharded"):
        return HourWorkload(name, batch)
    if name not in PRESETS:
        raise SystemExit(f"unknown workload {name!r}; choose from frontend, hour, {', '.join(PRESETS)}")
    return ModelWorkload(name, batch)
