"""multibox_b200 -- B200-native (sm_100a) implementation of the data-parallel hot
path of the Multibox detector: GT->prior optimal matching, multibox loss
forward/backward, detection-time decode / filter / top-k / NMS.

Module names mirror the reference's files (priors.py, loss.py, detect.py); the
arithmetic lives in hand-written CUDA behind the C ABI of
include/multibox_b200.h (multibox_b200/csrc/).  There is no CPU fallback.
"""
from . import priors  # noqa: F401  (host-side, numpy only)

__all__ = ["priors", "loss", "detect", "dist", "synth"]
__version__ = "0.1.0"


def __getattr__(name):
    # loss / detect / dist import torch and bind the CUDA library lazily
    if name in ("loss", "detect", "dist", "synth", "_lib", "_build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
