"""The reference's own evaluation entry point against the oracle and the drop-in.

tests/golden/reference_test_entry.npz was written by tests/golden/make_golden.py running /root/reference/test.py main()
UNMODIFIED -- its DataSet / DataLoader, AudioFeature, FullyCNNTester (graph creation, checkpoint restore, param_count, the
batch loop, rebuild_audio, the scores) -- on the job of tests/golden/ref_entry_case.py, with stand-ins only for the
packages that are not installed (oracle/ref_import.load_test_entry; TensorFlow's three operations: oracle/tf_standin.py).
What it hands to soundfile.write is what a user of the reference gets; the oracle chain (CPU) and the drop-in's test.main
(GPU) must deliver the same waveforms and scores from the same files."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_entry_case as case                                              # noqa: E402
from oracle import cpu_path, network, rebuild                             # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "reference_test_entry.npz"))


def _sdr(clean, de):                                                      # model_utils/utils.py:64-90 of the reference
    clean, de = np.asarray(clean, np.float64), np.asarray(de, np.float64)
    return 10 * np.log10(np.power(clean, 2).sum() / (np.power(de - clean, 2).sum() + np.finfo(np.float32).eps))


@pytest.mark.parametrize("arch", case.ARCHS)
def test_oracle_chain_matches_the_reference_entry_point(arch, golden, tmp_path):
    cfg, items = case.build(str(tmp_path), arch, "unused")
    w = network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True)
    assert golden["param_total_" + arch][0] == network.trainable_param_count(arch)     # what the reference's param_count printed
    sdrs = []
    for batch in ([0, 1], [2]):                                            # batch_size 2, no sampler: [0, 1], [2]
        mix = [case.read_pcm16(items[i][1]) for i in batch]
        for i, m in zip(batch, mix):
            assert abs(float(np.abs(m.astype(np.float64)).sum()) - golden["mixsum_%s_%d" % (arch, i)][0]) < 1e-6
        out = cpu_path.enhance_batch_cpu(mix, arch, w, net_dtype="float64")
        for i, o in zip(batch, out):
            ref = golden["de_%s_%d" % (arch, i)]
            clean = case.read_pcm16(items[i][0])
            assert len(o) == len(ref) == len(clean) == items[i][2]
            assert rebuild.sdr_db(ref, o) >= 100.0
            sdrs.append(_sdr(clean, o))
    assert abs(np.mean(sdrs) - golden["sdr_avg_" + arch][0]) < 1e-3


@pytest.mark.parametrize("arch", case.ARCHS)
def test_oracle_matches_the_reference_inference_engine(arch, golden, tmp_path):
    """infer.py:54-71 of the reference, as executed: the [F, T] magnitude and phase are RESHAPED (not transposed) to
    [1, T, F, 1] before the network, and the result goes through rebuild_audio."""
    from oracle import stft
    cfg, items = case.build(str(tmp_path), arch, "unused")
    w = network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True)
    sig = case.read_pcm16(items[1][1])
    X = stft.compute_spectrogram(sig, 8000, 0.032, 0.016, 256, True)
    mag = np.reshape(stft.power_spectrum(X), (1, X.shape[1], X.shape[0], 1))
    ph = np.reshape(stft.divide_phase(X), (1, X.shape[1], X.shape[0]))
    pred = network.forward(arch, w, mag.astype(np.float32), np.float64).astype(np.float32)
    out = rebuild.rebuild_audio([len(sig)], pred.squeeze(-1), ph, 8000, 32, 16)[0]
    ref = golden["infer_de_" + arch]
    assert len(out) == len(ref) == items[1][2]
    assert rebuild.sdr_db(ref, out) >= 100.0
    # ... and it is NOT what the transposed layout of test.py gives for the same file
    assert rebuild.sdr_db(golden["de_%s_1" % arch], ref) < 20.0


@pytest.mark.parametrize("arch", ["FullyCNNV2"])
def test_reference_entry_point_live(arch, golden, tmp_path, capsys):
    """Where /root/reference exists: run its test.py main() again and compare with the committed vectors."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from oracle import tf_standin
    entry, load_conf = ref_import.load_test_entry()
    prefix = str(tmp_path / "RCED.ckpt")
    np.savez(prefix + ".standin.npz", **network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True))
    cfg, items = case.build(str(tmp_path), arch, prefix)
    ref_import.WRITTEN.clear()
    tf_standin.reset_graph()
    entry.main(load_conf(cfg), 1)
    assert "Total number of Parameters: %d" % golden["param_total_" + arch][0] in capsys.readouterr().out
    for i in range(len(items)):
        de, rate = ref_import.WRITTEN[os.path.join(str(tmp_path), "out", "utt%d_de.wav" % i)]
        assert rate == 8000 and rebuild.sdr_db(golden["de_%s_%d" % (arch, i)], de) >= 120.0


@pytest.mark.gpu
@pytest.mark.parametrize("arch", case.ARCHS)
def test_dropin_entry_point_matches_the_reference_entry_point(arch, golden, tmp_path, monkeypatch, capsys):
    """fullycnnspeechenhancement_b200/test.py main() on the same files, a TensorFlow-format checkpoint of the same weights
    (model_utils/ckpt.py): waveforms within 60 dB of what the reference's entry point wrote, the same scores, the same
    parameter total."""
    pytest.importorskip("torch")
    from fullycnnspeechenhancement_b200 import test as dropin_entry
    from fullycnnspeechenhancement_b200.config import load_conf_info
    from fullycnnspeechenhancement_b200.model_utils import ckpt, tester as dropin_tester
    prefix = str(tmp_path / "ckpt" / ("RCED_%s_0_9.ckpt" % arch))
    ckpt.write_checkpoint(prefix, network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True))
    cfg, items = case.build(str(tmp_path), arch, prefix)
    written = {}
    monkeypatch.setattr(dropin_tester.audio_io, "write_wav", lambda path, data, rate: written.__setitem__(path, np.array(data)))
    made = []
    real_tester = dropin_entry.FullyCNNTester

    def recording(cfg_):
        made.append(real_tester(cfg_))
        return made[-1]
    monkeypatch.setattr(dropin_entry, "FullyCNNTester", recording)
    sdr_avg = dropin_entry.main(load_conf_info(cfg), 1)
    assert "Total number of Parameters: %d" % golden["param_total_" + arch][0] in capsys.readouterr().out
    for i, (pc, pm, L) in enumerate(items):
        de = written[os.path.join(str(tmp_path), "out", "utt%d_de.wav" % i)]
        ref = golden["de_%s_%d" % (arch, i)]
        assert len(de) == len(ref) == L
        assert rebuild.sdr_db(ref, de) >= 60.0
        assert np.array_equal(np.asarray(written[os.path.join(str(tmp_path), "out", "utt%d_mix.wav" % i)], np.float32),
                              case.read_pcm16(pm))
    assert abs(sdr_avg - golden["sdr_avg_" + arch][0]) < 1e-2
    assert abs(made[0].stoi_score.avg - golden["stoi_avg_" + arch][0]) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("arch", case.ARCHS)
def test_dropin_inference_engine_matches_the_reference_inference_engine(arch, golden, tmp_path, monkeypatch):
    """fullycnnspeechenhancement_b200/infer.py InferenceEngine(cfg).denoise(file) against what the reference's own
    InferenceEngine wrote for the same file and weights."""
    pytest.importorskip("torch")
    from fullycnnspeechenhancement_b200 import infer as dropin_infer
    from fullycnnspeechenhancement_b200.config import load_conf_info
    from fullycnnspeechenhancement_b200.model_utils import ckpt
    prefix = str(tmp_path / "ckpt" / ("RCED_%s_0_9.ckpt" % arch))
    ckpt.write_checkpoint(prefix, network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True))
    cfg, items = case.build(str(tmp_path), arch, prefix)
    written = {}
    monkeypatch.setattr(dropin_infer.audio_io, "write_wav", lambda path, data, rate: written.__setitem__(path, np.array(data)))
    dropin_infer.InferenceEngine(load_conf_info(cfg)).denoise(items[1][1])
    (path, de), = written.items()
    assert path == os.path.join(str(tmp_path), "out", "utt1_noisy_de.wav")
    ref = golden["infer_de_" + arch]
    assert len(de) == len(ref) and rebuild.sdr_db(ref, de) >= 60.0
