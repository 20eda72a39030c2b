"""B200-native (sm_100a) implementation of the YourMT3 inference hot path.

Host-side modules mirror the upstream ``amt/src/model`` interfaces; all device
work goes through the C ABI in ``include/ymt3_b200.h`` (``csrc/libymt3_b200.so``).
"""
from .config import audio_cfg, model_cfg, get_audio_cfg, get_model_cfg  # noqa: F401
from .spectrogram import (Melspectrogram, Spectrogram,  # noqa: F401
                          get_spectrogram_layer_from_audio_cfg, mel_filterbank)

from .t5mod import T5EncoderYMT3, T5DecoderYMT3, MultiChannelT5Decoder  # noqa: F401
from .t5mod_helper import task_cond_dec_generate  # noqa: F401
from .lm_head import LMHead  # noqa: F401
from .ymt3 import YourMT3  # noqa: F401
from .init_utils import init_nondegenerate_  # noqa: F401
from .checkpoint import load_checkpoint  # noqa: F401
from .midi import write_midi  # noqa: F401

__version__ = "0.1.0"
