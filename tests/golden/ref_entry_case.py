"""The small evaluation job behind tests/golden/reference_test_entry.npz: three clean / noisy wav pairs of different
lengths, a paired manifest, a test.cfg with the reference's keys.  Used by make_golden.py (which runs the reference's own
test.py on it) and by the tests (which run the drop-in and the oracle on the same files)."""
import json
import os

import numpy as np
from scipy.io import wavfile

LENGTHS = [9000, 6100, 4000]          # 70, 48 and 31 frames: a batch of two (ragged) and a batch of one
SEEDS = [170, 171, 172]
WEIGHT_SEED = 11
ARCHS = ["FullyCNN", "FullyCNNV2", "FullyCNNV3"]


def write_pcm16(path, x, rate=8000):
    x = np.clip(np.asarray(x, np.float64), -1.0, 1.0 - 1.0 / 32768.0)
    wavfile.write(path, rate, np.round(x * 32768.0).astype(np.int16))


def build(directory, net_work, checkpoint_prefix):
    """Writes the wavs, the manifest and the cfg into ``directory``; returns (cfg path, [(clean path, noisy path, L)])."""
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    os.makedirs(directory, exist_ok=True)
    items = []
    manifest = os.path.join(directory, "manifest.dev")
    with open(manifest, "w") as f:
        for i, (seed, L) in enumerate(zip(SEEDS, LENGTHS)):
            noisy, clean = noisy_utterance(seed, L, return_clean=True)
            pc, pm = os.path.join(directory, "utt%d.wav" % i), os.path.join(directory, "utt%d_noisy.wav" % i)
            write_pcm16(pc, clean)
            write_pcm16(pm, noisy)
            items.append((pc, pm, L))
            f.write(json.dumps({"audio_filepath": pc, "clean_audio_filepath": pc, "mix_audio_filepath": pm,
                                "duration": L / 8000.0}) + "\n")
    cfg = os.path.join(directory, "test.cfg")
    with open(cfg, "w") as f:
        f.write("\n".join(["[testing]", "batch_size=2", "checkpoint_filepath=%s" % checkpoint_prefix, "", "[model]",
                           "net_arch=RCED", "net_work=%s" % net_work, "", "[data]", "snr=0", "sample_rate=8000", "nfft=256",
                           "feature_dim=129", "window_ms=32", "stride_ms=16", "windows=hanning",
                           "audio_save_path=%s" % os.path.join(directory, "out"), "test_manifest_path=%s" % manifest]) + "\n")
    return cfg, items


def read_pcm16(path):
    rate, data = wavfile.read(path)
    assert rate == 8000 and data.dtype == np.int16
    return data.astype(np.float32) / 32768.0
