"""recnet_b200: B200-native (sm_100a) hot path of RecNet -- decoder + reconstructors, forward and backward.

Import name: ``recnet_b200`` (see recnet_b200.py at the repo root; this directory's name,
``reconstruction-network-for-video-captioning_b200``, is not a valid Python identifier).
"""
from . import _lib
from .config import EvalConfig, TrainConfig
from .models import Decoder, GlobalReconstructor, LocalReconstructor

__all__ = ["Decoder", "GlobalReconstructor", "LocalReconstructor", "TrainConfig", "EvalConfig", "_lib"]
