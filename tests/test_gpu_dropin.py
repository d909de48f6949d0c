"""GPU: the reference-shaped classes (config -> model construction -> checkpoint restore ->
test_step / denoise / test) drive the CUDA path and agree with the oracle."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from fullycnnspeechenhancement_b200 import audio_io                      # noqa: E402
from fullycnnspeechenhancement_b200.config import load_conf_info         # noqa: E402
from fullycnnspeechenhancement_b200.data_utils.data_loader import AudioParser, DataLoader, DataSet  # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import ckpt              # noqa: E402
from fullycnnspeechenhancement_b200.model_utils.model import FullyCNNSEModelV3  # noqa: E402
from fullycnnspeechenhancement_b200.model_utils.tester import FullyCNNTester    # noqa: E402
from fullycnnspeechenhancement_b200.model_utils.utils import AudioReBuild       # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance         # noqa: E402
from oracle import network, rebuild, stft                                # noqa: E402


def _write_cfg(path, section, ckpt_path, net_work, save, manifest=None):
    lines = ["[%s]" % section, "batch_size=2", "checkpoint_filepath=%s" % ckpt_path, "", "[model]", "net_arch=RCED",
             "net_work=%s" % net_work, "", "[data]", "snr=0", "sample_rate=8000", "nfft=256", "feature_dim=129",
             "window_ms=32", "stride_ms=16", "windows=hanning", "audio_save_path=%s" % save]
    if manifest:
        lines.append("test_manifest_path=%s" % manifest)
    open(path, "w").write("\n".join(lines) + "\n")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    w = network.random_weights("FullyCNNV2", seed=11, randomize_bn=True)
    prefix = str(d / "ckpt" / "RCED_FullyCNNV2_0_9.ckpt")
    ckpt.write_checkpoint(prefix, w)
    wavs = []
    for i, L in enumerate([16000, 12345, 8000]):
        p = str(d / ("utt%d.wav" % i))
        audio_io.write_wav(p, noisy_utterance(70 + i, L), 8000)
        wavs.append(p)
    # paired manifest (no noise manifest): the keys DataSet.__getitem__ reads at data_loader.py:113-118 of
    # the reference, plus audio_filepath, which its tester uses to name the outputs (tester.py:148)
    manifest = str(d / "manifest.dev")
    with open(manifest, "w") as f:
        for p, L in zip(wavs, [16000, 12345, 8000]):
            f.write(json.dumps({"audio_filepath": p, "clean_audio_filepath": p, "mix_audio_filepath": p,
                                "duration": L / 8000.0}) + "\n")
    # speech + noise manifests (mixing at snr dB inside the loader)
    speech_manifest, noise_manifest = str(d / "manifest.speech"), str(d / "manifest.noise")
    noise_wav = str(d / "noise0.wav")
    audio_io.write_wav(noise_wav, 0.3 * np.random.default_rng(5).normal(size=5000).clip(-3, 3) / 3, 8000)
    with open(speech_manifest, "w") as f:
        for p, L in zip(wavs, [16000, 12345, 8000]):
            f.write(json.dumps({"audio_filepath": p, "duration": L / 8000.0}) + "\n")
        f.write(json.dumps({"audio_filepath": wavs[0], "duration": 0.1}) + "\n")        # below min_duration: dropped
    with open(noise_manifest, "w") as f:
        f.write(json.dumps({"audio_filepath": noise_wav, "duration": 5000 / 8000.0}) + "\n")
    return dict(dir=d, weights=w, prefix=prefix, wavs=wavs, manifest=manifest, speech_manifest=speech_manifest,
                noise_manifest=noise_manifest)


def test_parser_extractor_rebuilder_match_oracle():
    parser = AudioParser(8000, 32, 16, use_complex=True)
    x = noisy_utterance(5, 9000)
    spec = parser.parse_audio(x)                                   # [129, T] complex128
    ref = stft.compute_spectrogram(x, 8000, 0.032, 0.016, 256, True)
    assert spec.shape == ref.shape and spec.dtype == np.complex128
    assert np.abs(spec - ref).max() / np.abs(ref).max() <= 1e-4
    mag = parser.extractor.power_spectrum(ref)
    ph = parser.extractor.divide_phase(ref)
    assert np.abs(mag - stft.power_spectrum(ref)).max() / np.abs(ref).max() <= 1e-6
    assert np.abs(ph - stft.divide_phase(ref)).max() <= 1e-5
    z = parser.extractor.divide_phase(np.zeros((2, 3), np.complex128))
    assert np.all(z == 1 + 0j)
    with pytest.raises(ValueError):                                # audio_feature.py:29-30
        parser.extractor.compute_spectrogram(x, 8000, 0.016, 0.032, 256)
    # reconstruction of a batch with the reference call signature
    X = stft.padding_batch([ref, stft.compute_spectrogram(x[:4000], 8000, 0.032, 0.016, 256, True)])
    spec3 = (stft.power_spectrum(X).squeeze(-1) * 0.7).astype(np.float32)
    ph3 = stft.divide_phase(X).squeeze(-1)
    for nfft in (512, 256):
        got = AudioReBuild(nfft=nfft).rebuild_audio([9000, 4000], spec3, ph3, 8000, 32.0, 16.0)
        want = rebuild.rebuild_audio([9000, 4000], spec3, ph3, 8000, 32.0, 16.0, nfft=nfft)
        for g, w_ in zip(got, want):
            assert g.dtype == np.float64 and len(g) == len(w_)
            assert rebuild.sdr_db(w_, g) >= 60.0


def test_tester_from_config_and_checkpoint(workdir, capsys):
    cfg_path = str(workdir["dir"] / "test.cfg")
    _write_cfg(cfg_path, "testing", workdir["prefix"], "FullyCNNV2", str(workdir["dir"] / "out_test"), workdir["manifest"])
    cfg = load_conf_info(cfg_path)
    tester = FullyCNNTester(cfg)
    assert "Total number of Parameters: 32192" in capsys.readouterr().out      # readme.md:66
    x = np.abs(np.random.default_rng(0).normal(0, 2, (2, 20, 129, 1)))          # float64 feed, like the reference
    y = tester.test_step(x)
    assert y.shape == (2, 20, 129, 1) and y.dtype == np.float32
    ref = network.forward("FullyCNNV2", workdir["weights"], x.astype(np.float32), np.float64)
    assert np.abs(y - ref).max() / np.abs(ref).max() <= 1e-4
    # the whole evaluation loop (test.py main)
    ds = DataSet(workdir["manifest"], None, sample_rate=8000, window_ms=32, stride_ms=16, use_complex=True)
    loader = DataLoader(ds, 2, sampler=None, num_works=1)
    tester.test(loader)
    out_dir = str(workdir["dir"] / "out_test")
    assert sorted(os.listdir(out_dir)) == sorted(
        [n for i in range(3) for n in ("utt%d.wav" % i, "utt%d_mix.wav" % i, "utt%d_de.wav" % i)])
    de, _ = audio_io.load_wav(os.path.join(out_dir, "utt1_de.wav"), 8000)
    assert len(de) == 12345
    # speech + noise manifests: item i is mixed with noise item i at snr dB, short items are dropped
    np.random.seed(3)
    ds = DataSet(workdir["speech_manifest"], workdir["noise_manifest"], sample_rate=8000, window_ms=32, stride_ms=16,
                 snr=5.0, use_complex=True)
    assert len(ds) == 3 and len(ds.noise_list) == 3
    (mix, speech), (mix_spec, speech_spec) = ds[1]
    assert len(mix) == len(speech) == 12345 and mix_spec.shape == speech_spec.shape == (129, 96)
    snr = 10 * np.log10(np.sum(speech.astype(np.float64) ** 2) / np.sum((mix - speech).astype(np.float64) ** 2))
    assert abs(snr - 5.0) < 1e-3
    assert np.abs(mix_spec - stft.compute_spectrogram(mix.astype(np.float32), 8000, 0.032, 0.016, 256, True)).max() \
        / np.abs(mix_spec).max() <= 1e-4
    loader = DataLoader(ds, 2, sampler=None, num_works=4)
    shapes = [(m.shape, len(ms)) for m, c, ms, cs in loader]
    assert shapes == [((2, 124, 129, 1), 2), ((1, 62, 129, 1), 1)]


def test_inference_engine_reshape_quirk_and_transpose(workdir):
    from fullycnnspeechenhancement_b200.infer import InferenceEngine
    cfg_path = str(workdir["dir"] / "infer.cfg")
    _write_cfg(cfg_path, "inference", workdir["prefix"], "FullyCNNV2", str(workdir["dir"] / "out_infer"))
    eng = InferenceEngine(load_conf_info(cfg_path))                  # [inference] section accepted
    sig, _ = audio_io.load_wav(workdir["wavs"][0], 8000)
    out = eng.enhance_signal(sig)
    # oracle of infer.py:54-71 including the reshape (not transpose) of the [F,T] arrays
    X = stft.compute_spectrogram(sig, 8000, 0.032, 0.016, 256, True)
    mag = np.reshape(stft.power_spectrum(X), (1, X.shape[1], X.shape[0], 1))
    ph = np.reshape(stft.divide_phase(X), (1, X.shape[1], X.shape[0]))
    pred = network.forward("FullyCNNV2", workdir["weights"], mag.astype(np.float32), np.float64).astype(np.float32)
    ref = rebuild.rebuild_audio([len(sig)], pred.squeeze(-1), ph, 8000, 32, 16)[0]
    assert len(out) == len(sig) and rebuild.sdr_db(ref, out) >= 60.0
    path = eng.denoise(workdir["wavs"][0])
    assert path.endswith("utt0_de.wav") and os.path.exists(path)
    # the layout test.py uses
    eng.layout = "transpose"
    out_t = eng.enhance_signal(sig)
    Xt = np.transpose(X)[None, :, :, None]
    pred_t = network.forward("FullyCNNV2", workdir["weights"], stft.power_spectrum(Xt).astype(np.float32), np.float64)
    ref_t = rebuild.rebuild_audio([len(sig)], pred_t.astype(np.float32)[..., 0], stft.divide_phase(Xt)[..., 0], 8000, 32, 16)[0]
    assert rebuild.sdr_db(ref_t, out_t) >= 60.0


def test_model_objects_and_frozen_graph(workdir, tmp_path):
    from fullycnnspeechenhancement_b200.freeze import FreezeEngine
    w3 = network.random_weights("FullyCNNV3", seed=4, randomize_bn=True)
    prefix = str(tmp_path / "v3.ckpt")
    ckpt.write_checkpoint(prefix, w3)
    pb = str(tmp_path / "v3.pb")
    assert FreezeEngine("FullyCNNV3").freeze_graph(prefix, pb) == "decode_final/BiasAdd"
    m = FullyCNNSEModelV3(is_training=False).restore(pb)             # frozen graph as weight source
    assert m.param_count() == 32653
    x = np.abs(np.random.default_rng(3).normal(0, 2, (1, 9, 129, 1))).astype(np.float32)
    ref = network.forward("FullyCNNV3", w3, x, np.float64)
    assert np.abs(m(x) - ref).max() / np.abs(ref).max() <= 1e-4
    xt = torch.from_numpy(x).cuda()
    yt = m(xt)                                                       # CUDA tensor in -> CUDA tensor out
    assert yt.is_cuda and np.abs(yt.cpu().numpy() - ref).max() / np.abs(ref).max() <= 1e-4
    with pytest.raises(RuntimeError):
        FullyCNNSEModelV3(is_training=False)(x)                      # no weights loaded
    with pytest.raises(KeyError):
        FullyCNNSEModelV3(is_training=False).set_weights({"decode_final/kernel": x})


def test_sdr_batch_matches_the_reference_formula():
    """rced_sdr_sums (float64 sums on the GPU) against SDR.sdr (model_utils/utils.py:68-78 of the reference)."""
    from fullycnnspeechenhancement_b200.model_utils.utils import SDR, sdr_batch
    rng = np.random.default_rng(3)
    refs = [rng.normal(0, 0.3, n).astype(np.float32) for n in (1, 255, 4096, 32000, 100001)]
    ests = [r + rng.normal(0, s, len(r)).astype(np.float32) for r, s in zip(refs, (0.1, 0.01, 0.3, 1e-3, 0.05))]
    ests[1] = refs[1].copy()                                     # perfect estimate: finite thanks to the epsilon
    got = sdr_batch(refs, ests)
    want = np.array([SDR()(r.astype(np.float64), e.astype(np.float64)) for r, e in zip(refs, ests)])
    assert got.shape == (5,) and np.all(np.isfinite(got))
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    assert sdr_batch([], []).shape == (0,)


# ---------------------------------------------------------------------------------------------------------------
# host-buffer C entry points (rced_enhance_host / rced_enhance_host_async / rced_host_sync)
# ---------------------------------------------------------------------------------------------------------------
def _host_eng(seed=77):
    from fullycnnspeechenhancement_b200.engine import Enhancer
    from oracle import network
    w = network.random_weights("FullyCNNV2", seed=seed, randomize_bn=True)
    return Enhancer("FullyCNNV2", w, device=0), w


def test_host_entry_point_equals_the_device_path():
    """rced_enhance_host (numpy in, numpy out, chunked over the library's streams) against the device-pointer path
    (rced_stft -> rced_forward -> rced_istft on caller-owned tensors), for ragged lengths, several chunks per call,
    asynchronous calls queued behind each other, and out_len shorter than the input.  The two paths pack the frames
    differently, and the tensor-core kernel's output layer sums its 32-tap diagonals in an order that depends on a
    row's position in its 32-row block: equal to FP32 rounding (>= 120 dB), not bit for bit; the same packing twice is
    bit-identical."""
    import torch
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    eng, _ = _host_eng()
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(200, 9000, 37)] + [1, 255, 256, 257]
    waves = [noisy_utterance(500 + i, n) for i, n in enumerate(lens)]
    eng.host_config(chunk_rows=300)                          # many chunks per call
    outs = eng.enhance(waves)
    # device path on the same packed batch
    plan = eng.plan(np.array(lens))
    d_wav = torch.from_numpy(np.concatenate(waves)).to(eng.device)
    d_out = torch.zeros_like(d_wav)
    eng.run_plan_device(plan, d_wav, d_out)
    torch.cuda.synchronize()
    ref = d_out.cpu().numpy()
    for o, off, n in zip(outs, plan["wav_off_host"], lens):
        r = ref[off:off + n].astype(np.float64)
        assert len(o) == n and np.sum((o - r) ** 2) <= 1e-12 * max(np.sum(r ** 2), 1e-30)
    # asynchronous calls on two buffer sets, one synchronisation
    t = eng.host_tables(np.array(lens))
    bufs = []
    for k in range(3):
        h_in = np.zeros(t["total"], np.float32)
        for w, o in zip(waves, t["wav_off"]):
            h_in[o:o + len(w)] = w
        h_out = np.full(t["total"], -7.0, np.float32)
        bufs.append((h_in, h_out))
        eng.enhance_host(h_in, h_out, t, sync=False)
    eng.host_sync()
    for h_in, h_out in bufs:
        for o, n, r in zip(t["out_off"], lens, outs):        # (the synchronous call cuts its chunks differently)
            assert np.array_equal(h_out[o:o + n], bufs[0][1][o:o + n])     # the same packing: bit-identical
            assert np.sum((h_out[o:o + n].astype(np.float64) - r) ** 2) <= 1e-12 * max(np.sum(r.astype(np.float64) ** 2), 1e-30)
    # truncated outputs (the reference cuts to len(clean_sig)): nothing behind out_len is written when the gap is large
    cut = [max(1, n - 40) for n in lens]
    t2 = eng.host_tables(np.array(lens), out_lens=cut)
    h_out = np.full(t2["total"], -7.0, np.float32)
    eng.enhance_host(bufs[0][0], h_out, t2, sync=True)
    for o, n, c, r in zip(t2["out_off"], lens, cut, outs):
        assert np.array_equal(h_out[o:o + c], r[:c])         # (same call mode and chunking as `outs`)
        if n - c >= 16 + 3:
            assert np.all(h_out[o + c + 16:o + n] == -7.0)
    eng.close()


def test_host_entry_point_recomputes_a_tripped_chunk_with_the_fp32_kernel():
    """The host pipeline checks the tensor-core launches' guard words when it synchronises and recomputes a tripped chunk
    from the caller's buffers with the FP32 kernel: the result equals the FP32 engine's."""
    from fullycnnspeechenhancement_b200.engine import Enhancer
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    from oracle import network
    w = network.random_weights("FullyCNNV2", seed=4321, randomize_bn=True)
    for L in network.layer_table("FullyCNNV2")[1:9]:
        w[L["scope"] + "/kernel"] = w[L["scope"] + "/kernel"] * np.float32(8.0)
    waves = [noisy_utterance(900 + i, 3000 + 517 * i) for i in range(6)]
    tc = Enhancer("FullyCNNV2", w, device=0)
    fp32 = Enhancer("FullyCNNV2", w, device=0, variant="ffma")
    assert tc.variant == "tc" and fp32.variant == "ffma"
    tc.host_config(chunk_rows=60)
    a = tc.enhance(waves)
    assert not tc.tc_status()[0] < 65504          # the guard did trip
    b = fp32.enhance(waves)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    tc.close()
    fp32.close()


def test_host_entry_point_through_a_relay_device():
    """rced_host_set_relay: the waveforms travel host -> peer GPU -> NVLink -> the handle's GPU and back; results are
    bit-identical to the direct route.  Needs two GPUs that are peers."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from fullycnnspeechenhancement_b200 import _lib
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    eng, _ = _host_eng(seed=78)
    waves = [noisy_utterance(700 + i, 2000 + 811 * i) for i in range(9)]
    eng.host_config(chunk_rows=100)
    direct = eng.enhance(waves)
    try:
        eng.host_set_relay(1)
    except _lib.RcedError as exc:
        if exc.code == _lib.ERR_STATE:
            pytest.skip("GPUs 0 and 1 are not peers")
        raise
    via = eng.enhance(waves)
    again = eng.enhance(waves)
    eng.host_set_relay(-1)
    back = eng.enhance(waves)
    for a, b, c, d in zip(direct, via, again, back):
        assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)
    eng.close()


def test_outputs_longer_than_the_noisy_input_follow_rebuild_audio():
    """FullyCNNTester.test truncates what rebuild_audio produced -- (T+1)*128 samples -- to len(clean_sig[i])
    (model_utils/tester.py:107-113, utils.py:181-182): a clean signal a little longer than the noisy one gets the samples
    rebuilt from the zero-padded last frames, not a cut at len(mix)."""
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    from oracle import network, rebuild, stft
    eng, w = _host_eng(seed=79)
    waves = [noisy_utterance(40, 1000), noisy_utterance(41, 5000), noisy_utterance(42, 300)]
    clean_lens = [1010, 5000, 5000]                      # 1010 <= 1024 rebuilt; 5000 <= 5120; 5000 > (2+1)*128 = 384 rebuilt
    outs = eng.enhance(waves, out_lens=clean_lens)
    assert [len(o) for o in outs] == [1010, 5000, 384]
    for wv, o, n in zip(waves, outs, clean_lens):
        X = stft.compute_spectrogram(wv, 8000, 0.032, 0.016, 256, True).T[None, :, :, None]
        mag = stft.power_spectrum(X).astype(np.float32)
        pred = network.forward("FullyCNNV2", w, mag, np.float64).astype(np.float32)
        ref = rebuild.rebuild_audio([n], pred[..., 0], stft.divide_phase(X)[..., 0], 8000, 32.0, 16.0)[0]
        assert len(ref) == len(o) and rebuild.sdr_db(ref, o) >= 60.0
    eng.close()


def test_host_entry_point_with_scattered_utterances_and_bad_arguments():
    """Utterances anywhere in the caller's buffers (any order, large gaps): copied one by one into a compact device layout;
    nothing outside the outputs is touched.  Argument errors come back as RCED_ERR_ARG, not as CUDA faults."""
    import ctypes
    from fullycnnspeechenhancement_b200 import _lib
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    eng, _ = _host_eng(seed=80)
    waves = [noisy_utterance(60 + i, n) for i, n in enumerate([3000, 1, 777, 4096, 12000])]
    want = eng.enhance(waves)
    lens = np.array([len(w) for w in waves], np.int32)
    off = np.array([900000, 5, 400000, 2000000, 100000], np.int64)          # arbitrary order, gaps of ~10^5..10^6 samples
    buf = np.zeros(2100000, np.float32)
    for w, o in zip(waves, off):
        buf[o:o + len(w)] = w
    out = np.full(2100000, -3.0, np.float32)
    t = {"n": len(waves), "wav_off": off, "wav_len": lens, "out_off": off, "out_len": lens}
    eng.enhance_host(buf, out, t, sync=True)
    mask = np.ones(len(out), bool)
    for w, o, r in zip(waves, off, want):
        assert np.array_equal(out[o:o + len(w)], r)
        mask[o:o + len(w)] = False
    assert np.all(out[mask] == -3.0)
    # argument checks
    lib, h = _lib.lib(), eng._h
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    bad_len = lens.copy()
    bad_len[1] = 0
    assert lib.rced_enhance_host(h, p(buf), p(off), p(bad_len), 5, 512, p(out), p(off), p(lens)) == _lib.ERR_ARG
    too_long = lens.copy()
    too_long[2] = 5000                                                          # > (T+1)*128 for a 777-sample utterance
    assert lib.rced_enhance_host(h, p(buf), p(off), p(lens), 5, 512, p(out), p(off), p(too_long)) == _lib.ERR_ARG
    assert lib.rced_enhance_host(h, p(buf), p(off), p(lens), 5, 384, p(out), p(off), p(lens)) == _lib.ERR_ARG   # irfft length
    assert lib.rced_enhance_host(h, p(buf), p(off), p(lens), 0, 512, p(out), p(off), p(lens)) == 0              # nothing to do
    assert lib.rced_enhance_host(h, None, p(off), p(lens), 5, 512, p(out), p(off), p(lens)) == _lib.ERR_ARG
    eng.close()


def test_host_pipeline_stress_short():
    """Five seconds of tools/host_stress.py: random batches, chunk sizes and call modes with several calls in flight; every
    output equals the utterance enhanced alone (FP32 kernel: bit for bit)."""
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "tools", "host_stress.py"), "5", "ffma"], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0 and "host stress ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
