"""Generate the committed golden fixtures under tests/golden/ from the INSTALLED
reference dependencies (torchaudio 2.11 / transformers 5.5 on CPU).

The mounted /root/reference has no code (README + LICENSE only), so the golden
vectors are produced by the libraries the upstream path calls, through the same
call sites upstream uses.  Re-run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def frontend():
    import torchaudio
    rng = np.random.default_rng(20240704)
    # (a) melspec, hop 128, short segment exercising both reflect edges + odd frame count
    L = 4100
    x = (rng.standard_normal((2, L)) * 0.1).astype(np.float32)
    x[1] *= np.linspace(0, 1, L, dtype=np.float32) ** 2          # dynamic range
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                                              f_max=8000.0, n_mels=512, power=1.0)
    y = torch.log(torch.clamp(ms(torch.from_numpy(x)), min=1e-5)).transpose(1, 2).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "logmel_melspec_hop128.npz"), audio=x, out=y.astype(np.float32),
                        torchaudio=np.array(torchaudio.__version__))
    # (b) spec, hop 300, drop DC
    L = 5000
    x = (rng.standard_normal((2, L)) * 0.1).astype(np.float32)
    sp = torchaudio.transforms.Spectrogram(n_fft=2048, hop_length=300, power=1.0)
    y = torch.log(torch.clamp(sp(torch.from_numpy(x)), min=1e-5))[:, 1:, :].transpose(1, 2).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "logmel_spec_hop300.npz"), audio=x, out=y.astype(np.float32))
    # (c) filterbank in sparse form
    fb = torchaudio.functional.melscale_fbanks(1025, 50.0, 8000.0, 512, 16000).numpy()
    idx = np.nonzero(fb)
    np.savez_compressed(os.path.join(OUT, "melscale_fbanks_512.npz"), rows=idx[0].astype(np.int32),
                        cols=idx[1].astype(np.int32), vals=fb[idx])
    # (d) power=2 variant, tiny
    L = 2200
    x = (rng.standard_normal((1, L)) * 0.3).astype(np.float32)
    ms2 = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=128, f_min=50.0,
                                               f_max=8000.0, n_mels=512, power=2.0)
    y = torch.log(torch.clamp(ms2(torch.from_numpy(x)), min=1e-5)).transpose(1, 2).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "logmel_melspec_power2.npz"), audio=x, out=y.astype(np.float32))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["frontend", "t5"]
    if "frontend" in which:
        frontend()
    if "t5" in which:
        try:
            from tools import make_golden_t5
            make_golden_t5.main(OUT)
        except ImportError:
            pass
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
