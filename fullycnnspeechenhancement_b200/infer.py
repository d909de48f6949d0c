"""Single-file enhancement with the reference's InferenceEngine interface (infer.py of the
reference):  python -m fullycnnspeechenhancement_b200.infer --cfg CFG --audio-file X.wav"""
import argparse
import os

import numpy as np

from . import audio_io
from .config import load_conf_info
from .data_utils.data_loader import AudioParser
from .model_utils.tester import BaseTester
from .model_utils.utils import AudioReBuild


def _as_reference_feeds(spec, extractor):
    """(mag [1,T,F,1], phase [1,T,F]) built the way infer.py:57-61 of the reference builds them: the
    [F,T] arrays are *reshaped*, not transposed, so time and frequency are scrambled (SURVEY.md
    section 0).  Kept because a drop-in must return what the reference returns."""
    n_bins, n_frames = spec.shape
    mag = extractor.power_spectrum(spec)
    phase = extractor.divide_phase(spec)
    return np.reshape(mag, (1, n_frames, n_bins, 1)), np.reshape(phase, (1, n_frames, n_bins))


class InferenceEngine(BaseTester):
    """Single-file enhancement (infer.py:19-77 of the reference).  ``layout="reshape"`` reproduces the
    reference literally; ``layout="transpose"`` is the layout test.py uses and runs K1 -> K2 -> K3 in
    one GPU pass."""

    def __init__(self, infer_config, layout="reshape"):
        super(InferenceEngine, self).__init__(infer_config)
        data = lambda key, cast=str: cast(infer_config.get("data", key))
        self.sample_rate = data("sample_rate", int)
        self.feature_dim = data("feature_dim", int)
        self.audio_save_path = data("audio_save_path")
        self.window_ms = data("window_ms", int)
        self.stride_ms = data("stride_ms", int)
        if layout not in ("reshape", "transpose"):
            raise ValueError("layout must be 'reshape' or 'transpose'")
        self.layout = layout
        for step in (self.creat_graph, self._init_session, self._load_checkpoint, self.param_count):
            step()
        self.audio_parser = AudioParser(self.sample_rate, self.window_ms, self.stride_ms, use_complex=True)
        self.audio_rebuilder = AudioReBuild()

    def enhance_signal(self, sig):
        """float waveform -> enhanced float64 waveform of the same length."""
        if self.layout == "transpose":
            return self.model.engine().enhance([sig])[0].astype(np.float64)
        spec = self.audio_parser.parse_audio(sig)                                  # [F,T] complex
        mag, phase = _as_reference_feeds(spec, self.audio_parser.extractor)
        pred = self.test_step(mag)
        rebuilt = self.audio_rebuilder.rebuild_audio([len(sig)], pred.squeeze(-1), phase, self.sample_rate,
                                                     self.window_ms, self.stride_ms)
        return rebuilt[0]

    def denoise(self, audio_file):
        """Enhances one wav file and writes ``<audio_save_path>/<name>_de.wav`` (infer.py:72-77)."""
        sig, _ = self.audio_parser.load_audio(audio_file)
        enhanced = self.enhance_signal(sig)
        os.makedirs(self.audio_save_path, exist_ok=True)
        target = os.path.join(self.audio_save_path, os.path.basename(audio_file).replace(".wav", "_de.wav"))
        audio_io.write_wav(target, enhanced, self.sample_rate)
        print("Saving denoise file to {}.".format(target))
        return target


def main(argv=None):
    ap = argparse.ArgumentParser(description="Inference")
    ap.add_argument("--cfg", default="", type=str, help="cfg file for infer")
    ap.add_argument("--audio-file", default="", type=str, help="audio to denoise")
    ap.add_argument("--layout", default="reshape", choices=["reshape", "transpose"])
    args = ap.parse_args(argv)
    InferenceEngine(load_conf_info(args.cfg), layout=args.layout).denoise(args.audio_file)


if __name__ == "__main__":
    main()
