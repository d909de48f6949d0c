"""Process-wide defaults shared by the reference-shaped classes (device selection, lazily
created library-only helpers)."""
import os

_default_device = None


def default_device():
    """CUDA device index used by classes that the reference constructs without a device
    argument: $RCED_DEVICE, else $LOCAL_RANK (torchrun), else 0."""
    global _default_device
    if _default_device is None:
        _default_device = int(os.environ.get("RCED_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return _default_device


def set_default_device(index):
    global _default_device
    _default_device = int(index)
