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


class InferenceEngine(BaseTester):
    def __init__(self, infer_config, layout="reshape"):
        """``layout``: "reshape" reproduces infer.py:59-61 of the reference literally -- the
        [F,T] spectrogram is *reshaped* (not transposed) to (1,T,F,1), which scrambles time and
        frequency; "transpose" is the layout test.py uses."""
        super(InferenceEngine, self).__init__(infer_config)
        self.sample_rate = int(infer_config.get("data", "sample_rate"))
        self.feature_dim = int(infer_config.get("data", "feature_dim"))
        self.audio_save_path = infer_config.get("data", "audio_save_path")
        self.window_ms = int(infer_config.get("data", "window_ms"))
        self.stride_ms = int(infer_config.get("data", "stride_ms"))
        self.layout = layout
        self.creat_graph()
        self._init_session()
        self._load_checkpoint()
        self.param_count()
        self.audio_parser = AudioParser(self.sample_rate, self.window_ms, self.stride_ms, use_complex=True)
        self.audio_rebuilder = AudioReBuild()

    def enhance_signal(self, sig):
        sig_length = len(sig)
        if self.layout == "transpose":
            return self.model.engine().enhance([sig])[0].astype(np.float64)
        spec = self.audio_parser.parse_audio(sig)                                  # [F,T] complex
        mag = self.audio_parser.extractor.power_spectrum(spec)
        mag = np.reshape(mag, (1, mag.shape[1], mag.shape[0], 1))
        phase = self.audio_parser.extractor.divide_phase(spec)
        phase = np.reshape(phase, (1, phase.shape[1], phase.shape[0]))
        pred = self.test_step(mag)
        return self.audio_rebuilder.rebuild_audio([sig_length], pred.squeeze(-1), phase, self.sample_rate,
                                                  self.window_ms, self.stride_ms)[0]

    def denoise(self, audio_file):
        sig, _ = self.audio_parser.load_audio(audio_file)
        out = self.enhance_signal(sig)
        if not os.path.exists(self.audio_save_path):
            os.makedirs(self.audio_save_path)
        path = os.path.join(self.audio_save_path, os.path.basename(audio_file).replace(".wav", "_de.wav"))
        audio_io.write_wav(path, out, self.sample_rate)
        print("Saving denoise file to {}.".format(path))
        return path


def main(argv=None):
    ap = argparse.ArgumentParser(description="Inference")
    ap.add_argument("--cfg", default="", type=str, help="cfg file for infer")
    ap.add_argument("--audio-file", default="", type=str, help="audio to denoise")
    ap.add_argument("--layout", default="reshape", choices=["reshape", "transpose"])
    args = ap.parse_args(argv)
    InferenceEngine(load_conf_info(args.cfg), layout=args.layout).denoise(args.audio_file)


if __name__ == "__main__":
    main()
