"""Import the UNMODIFIED reference numpy code from /root/reference (authoring
container only -- that path does not exist on the GPU box, and nothing in the
``-m gpu`` tests, smoke() or bench.py may call this).

Two shims are needed: numpy 2 removed ``np.mat`` (used at
data_utils/audio_feature.py:76) and librosa / pypesq / pystoi are not installed
(imported at model_utils/utils.py:7-10, data_utils/data_loader.py:9).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "data_utils"))


def load():
    """Returns (AudioFeature, AudioReBuild, DataLoader, AudioParser) classes of the reference."""
    import numpy as np
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    for name in ("librosa", "pypesq", "pystoi", "soundfile"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.pesq = m.stoi = m.load = lambda *a, **k: None
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference package names (data_utils, model_utils) are imported under their own
    # names; the product mirrors live inside fullycnnspeechenhancement_b200/, so no clash.
    from data_utils.audio_feature import AudioFeature
    from data_utils.data_loader import DataLoader, AudioParser
    from model_utils.utils import AudioReBuild
    return AudioFeature, AudioReBuild, DataLoader, AudioParser


def load_models():
    """The reference's own model classes {name: class} (model_utils/model.py), imported UNMODIFIED with
    oracle/tf_standin.py standing in for TensorFlow: `cls(is_training=False)(x)` runs the reference's wiring on the
    variables given to tf_standin.set_variables()."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    from oracle import tf_standin
    tf_standin.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    model = importlib.import_module("model_utils.model")
    return {"FullyCNN": model.FullyCNNSEModel, "FullyCNNV2": model.FullyCNNSEModelV2, "FullyCNNV3": model.FullyCNNSEModelV3}
