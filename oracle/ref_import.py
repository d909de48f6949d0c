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


class _SeqParallel(object):
    """joblib.Parallel run in the calling process: the reference's worker processes could not import the stand-in modules."""

    def __init__(self, n_jobs=None, **kwargs):
        pass

    def __call__(self, tasks):
        return [f(*a, **k) for f, a, k in tasks]


WRITTEN = {}   # soundfile.write stand-in: path -> (samples, sample rate)


def load_test_entry():
    """The reference's entry points, UNMODIFIED: returns (the module of /root/reference/test.py -- main, FullyCNNTester,
    DataSet, DataLoader; its attribute `infer` is the module of /root/reference/infer.py with InferenceEngine --, the
    reference's load_conf_info) with stand-ins for what is not installed --
      tensorflow            oracle/tf_standin.py (graph mode; Saver.restore reads <checkpoint>.standin.npz)
      librosa.load          16-bit PCM wav via scipy.io.wavfile, samples / 32768 as float32 (no resampling: the file's
                            rate must be the requested one)
      soundfile.write       records the arrays in ref_import.WRITTEN instead of writing files
      pypesq.pesq           0.0 (ITU-T P.862 is not available)
      pystoi.stoi           oracle/stoi_ref.py
      joblib.Parallel       sequential, in process (in model_utils/tester.py and data_utils/data_loader.py)
    """
    import importlib
    import importlib.util
    import numpy as np
    from scipy.io import wavfile
    from oracle import stoi_ref
    load_models()                                  # tensorflow stand-in + sys.path
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix

    def librosa_load(path, sr=None):
        rate, data = wavfile.read(path)
        if data.dtype != np.int16 or (sr is not None and rate != sr):
            raise NotImplementedError("librosa.load stand-in: 16-bit PCM at the requested rate only")
        return data.astype(np.float32) / 32768.0, rate

    def sf_write(path, data, samplerate=None, **kwargs):
        WRITTEN[path] = (np.array(data), samplerate)

    mods = {"librosa": dict(load=librosa_load), "soundfile": dict(write=sf_write), "pypesq": dict(pesq=lambda a, b, sr: 0.0),
            "pystoi": dict(stoi=lambda a, b, sr: stoi_ref.stoi(a, b, sr))}
    for name, attrs in mods.items():
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    for name in ("model_utils.utils", "model_utils.tester", "data_utils.data_loader"):   # bind the stand-ins above
        if name in sys.modules:
            importlib.reload(sys.modules[name])
    tester = importlib.import_module("model_utils.tester")
    loader = importlib.import_module("data_utils.data_loader")
    tester.Parallel = _SeqParallel
    loader.Parallel = _SeqParallel
    def entry_module(filename, modname):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, filename))
        module = importlib.util.module_from_spec(spec)
        saved = sys.modules.get("config")
        sys.modules.pop("config", None)                # `from config import load_conf_info`: the reference's own config.py
        try:
            spec.loader.exec_module(module)
        finally:
            sys.modules.pop("config", None)
            if saved is not None:
                sys.modules["config"] = saved
        return module
    entry = entry_module("test.py", "rced_reference_test_entry")
    entry.infer = entry_module("infer.py", "rced_reference_infer_entry")   # InferenceEngine (infer.py:19-78)
    return entry, entry.load_conf_info
