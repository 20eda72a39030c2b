"""B200-native (sm_100a) implementation of the YourMT3 inference hot path.

Host-side modules mirror the upstream ``amt/src/model`` interfaces; all device
work goes through the C ABI in ``include/ymt3_b200.h`` (``csrc/libymt3_b200.so``).
"""
from .config import audio_cfg, model_cfg, get_audio_cfg, get_model_cfg  # noqa: F401
from .spectrogram import (Melspectrogram, Spectrogram,  # noqa: F401
                          get_spectrogram_layer_from_audio_cfg, mel_filterbank)

__version__ = "0.1.0"
