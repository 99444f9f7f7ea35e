"""vad_b200 -- B200-native (sm_100a) MFCC + feed-forward VAD hot path of nameofuser1/vad.

Python call surface of the reference (mfcc.py functions, Analyser.load_init_inactive_frames /
feed_frame, process_file / split_into_frames / scale_features) over a thin C ABI
(include/vadb200.h, vad_b200/libvadb200.so) that launches hand-written CUDA kernels.
There is no CPU fallback: importing the numeric modules without the built library raises.
"""
from . import config  # noqa: F401

__all__ = ["config", "mfcc", "analyser", "batch", "runtime", "synth"]
__version__ = "0.1.0"
