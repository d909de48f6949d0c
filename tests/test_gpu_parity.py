"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the oracle.

Tolerances (BASELINE.md section 4 / north_star):
  * frame counts, padded rows, output lengths: exact;
  * magnitudes (K1) and network outputs (K2): max|d| / max|ref| <= 1e-4 against the float64 oracle;
  * waveforms (K3 and end to end): SNR = 10 log10(sum ref^2 / sum (out-ref)^2) >= 60 dB.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from fullycnnspeechenhancement_b200 import _lib                      # noqa: E402
from fullycnnspeechenhancement_b200.engine import Enhancer, num_frames  # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance       # noqa: E402
from oracle import network, rebuild, stft                              # noqa: E402

ARCHS = ["FullyCNN", "FullyCNNV2", "FullyCNNV3"]
MAG_TOL = 1e-4
SNR_DB = 60.0


def rel_err(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / np.abs(ref).max())


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def engines():
    out = {}
    for a in ARCHS:
        w = network.random_weights(a, seed=1234, randomize_bn=True)
        out[a] = (Enhancer(a, w, device=0, variant="ffma"), w)   # the FP32 FFMA kernel; tests/test_gpu_tc.py covers "tc"
    yield out
    for e, _ in out.values():
        e.close()


def test_tmem_roundtrip():
    _lib.check(_lib.lib().rced_selftest_tmem(0))


def _stft_oracle_rows(waves, rows_per_utt):
    mags, phases = [], []
    for w, rows in zip(waves, rows_per_utt):
        X = stft.compute_spectrogram(w, 8000, 0.032, 0.016, 256, True).T        # [T,129]
        assert X.shape[0] == num_frames(len(w))
        Xp = np.zeros((rows, 129), np.complex128)
        Xp[:X.shape[0]] = X
        mags.append(stft.power_spectrum(Xp))
        phases.append(stft.divide_phase(Xp))
    return np.concatenate(mags), np.concatenate(phases)


def _upload_batch(dev, waves, rows_per_utt):
    lens = np.array([len(w) for w in waves], np.int64)
    wav_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    row_off = np.concatenate([[0], np.cumsum(rows_per_utt)]).astype(np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return (t(np.concatenate(waves).astype(np.float32)), t(wav_off), t(lens.astype(np.int32)), t(row_off),
            int(row_off[-1]))


@pytest.mark.parametrize("pad_rows", [0, 3])
def test_stft_matches_oracle(dev, engines, pad_rows):
    eng = engines["FullyCNNV2"][0]
    lengths = [1, 100, 255, 256, 257, 383, 384, 385, 1000, 4000, 32000]
    waves = [noisy_utterance(100 + i, L) for i, L in enumerate(lengths)]
    rows = [int(num_frames(L)) + pad_rows for L in lengths]
    d_wav, d_off, d_len, d_row, total = _upload_batch(dev, waves, rows)
    mag, phase = eng.stft_device(d_wav, d_off, d_len, d_row, total)
    torch.cuda.synchronize()
    ref_mag, ref_phase = _stft_oracle_rows(waves, rows)
    mag = mag.cpu().numpy()
    ph = phase.cpu().numpy()
    assert mag.shape == ref_mag.shape
    assert rel_err(mag, ref_mag) <= MAG_TOL
    # padding rows are exactly (0, 1+0j)
    pos = 0
    for L, r in zip(lengths, rows):
        T = int(num_frames(L))
        assert np.all(mag[pos + T:pos + r] == 0)
        assert np.all(ph[pos + T:pos + r, :, 0] == 1) and np.all(ph[pos + T:pos + r, :, 1] == 0)
        pos += r
    # phase: compare where the bin is not numerically empty
    strong = ref_mag > 1e-3 * ref_mag.max()
    dphi = np.abs((ph[..., 0] + 1j * ph[..., 1]) - ref_phase)
    assert dphi[strong].max() < 1e-3
    # reconstructed complex spectrum within tolerance everywhere
    X = mag * (ph[..., 0] + 1j * ph[..., 1])
    Xr = ref_mag * ref_phase
    assert np.abs(X - Xr).max() / np.abs(Xr).max() <= MAG_TOL


@pytest.mark.parametrize("arch", ARCHS)
@pytest.mark.parametrize("tmem", [True, False])
def test_network_small_T(dev, engines, arch, tmem):
    eng, w = engines[arch]
    eng.set_skip_in_tmem(tmem)
    rng = np.random.default_rng(77)
    for T in (1, 7, 8, 9, 12):
        x = np.abs(rng.normal(0, 3.0, (2, T, 129, 1))).astype(np.float32)
        ref = network.forward(arch, w, x, np.float64)[..., 0].reshape(2 * T, 129)
        row_off = torch.tensor([0, T, 2 * T], dtype=torch.int64, device=dev)
        pred = eng.forward_device(torch.from_numpy(x.reshape(2 * T, 129)).to(dev), row_off)
        torch.cuda.synchronize()
        assert rel_err(pred.cpu().numpy(), ref) <= MAG_TOL, (arch, T)
    eng.set_skip_in_tmem(True)


@pytest.mark.parametrize("arch", ARCHS)
def test_network_many_frames_ragged(dev, engines, arch):
    """More frames than resident warps (slot reuse, prefetch, halo restoration) and utterance
    boundaries falling anywhere; compared with the float64 oracle run per utterance."""
    eng, w = engines[arch]
    rng = np.random.default_rng(5)
    Ts = [249, 1, 499, 124, 3, 700, 62]
    xs = [np.abs(rng.normal(0, 2.0, (1, T, 129, 1))).astype(np.float32) for T in Ts]
    ref = np.concatenate([network.forward(arch, w, x, np.float64)[0, :, :, 0] for x in xs])
    row_off = torch.from_numpy(np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)).to(dev)
    mag = torch.from_numpy(np.concatenate([x[0, :, :, 0] for x in xs])).to(dev)
    for tmem in (True, False):
        eng.set_skip_in_tmem(tmem)
        pred = eng.forward_device(mag, row_off)
        torch.cuda.synchronize()
        assert rel_err(pred.cpu().numpy(), ref) <= MAG_TOL, (arch, tmem)
    eng.set_skip_in_tmem(True)


def test_network_batch_padding_invariance(dev, engines):
    """Dense [N,T_max,129] layout with zero-padded tails: an utterance's valid frames do not
    depend on its batch mates."""
    arch = "FullyCNNV2"
    eng, w = engines[arch]
    rng = np.random.default_rng(9)
    Tmax, Ts = 40, [40, 17, 5]
    x = np.zeros((3, Tmax, 129, 1), np.float32)
    for i, T in enumerate(Ts):
        x[i, :T] = np.abs(rng.normal(0, 2.0, (T, 129, 1)))
    ref = network.forward(arch, w, x, np.float64)[..., 0]
    row_off = torch.arange(4, dtype=torch.int64, device=dev) * Tmax
    pred = eng.forward_device(torch.from_numpy(x.reshape(3 * Tmax, 129)).to(dev), row_off)
    torch.cuda.synchronize()
    pred = pred.cpu().numpy().reshape(3, Tmax, 129)
    assert rel_err(pred, ref) <= MAG_TOL
    for i, T in enumerate(Ts):      # alone == in batch (bitwise: same kernel, same data)
        alone = eng.forward_device(torch.from_numpy(x[i].reshape(Tmax, 129)).to(dev),
                                   torch.tensor([0, Tmax], dtype=torch.int64, device=dev))
        torch.cuda.synchronize()
        assert np.array_equal(alone.cpu().numpy(), pred[i])


@pytest.mark.parametrize("irfft_n", [512, 256])
def test_istft_matches_oracle(dev, engines, irfft_n):
    eng = engines["FullyCNNV2"][0]
    rng = np.random.default_rng(3)
    lengths = [100, 256, 1000, 4000, 32000, 70001]
    extra = [0, 2, 0, 1, 0, 0]                       # batch-padding rows
    waves = [noisy_utterance(200 + i, L) for i, L in enumerate(lengths)]
    rows = [int(num_frames(L)) + e for L, e in zip(lengths, extra)]
    ref_mag, ref_phase = _stft_oracle_rows(waves, rows)
    pred = (ref_mag * rng.uniform(0.2, 1.2, (1, 129)) - 0.05).astype(np.float32)
    row_off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    out_off = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ph = np.stack([ref_phase.real, ref_phase.imag], -1).astype(np.float32)
    out = torch.full((int(sum(lengths)) + 7,), 123.0, dtype=torch.float32, device=dev)
    eng.istft_device(t(pred), t(ph), t(row_off), max(rows), out, t(out_off), t(np.array(lengths, np.int32)),
                     irfft_n=irfft_n)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert np.all(out[-7:] == 123.0)                 # nothing written past the last utterance
    pos = 0
    for i, (L, r) in enumerate(zip(lengths, rows)):
        sl = slice(int(row_off[i]), int(row_off[i + 1]))
        # float32 phase is what the kernel was given: feed the oracle the same values
        ph64 = ph[sl, :, 0].astype(np.float64) + 1j * ph[sl, :, 1].astype(np.float64)
        ref = rebuild.rebuild_audio([L], pred[sl][None], ph64[None], 8000, 32.0, 16.0, nfft=irfft_n)[0]
        assert len(ref) == L
        snr = rebuild.sdr_db(ref, out[pos:pos + L])
        assert snr >= SNR_DB, (L, irfft_n, snr)
        pos += L


@pytest.mark.parametrize("arch", ARCHS)
def test_end_to_end_vs_oracle_chain(dev, engines, arch):
    """waveform -> enhanced waveform through Enhancer.enhance (host API, chunked over streams)
    against the oracle chain STFT -> network(float64) -> rebuild."""
    eng, w = engines[arch]
    lengths = [32000, 16000, 24001, 100, 4000, 32000, 9000]
    waves = [noisy_utterance(300 + i, L) for i, L in enumerate(lengths)]
    outs = eng.enhance(waves, chunk_utts=3)
    assert [len(o) for o in outs] == lengths
    for wv, o in zip(waves, outs):
        X = stft.compute_spectrogram(wv, 8000, 0.032, 0.016, 256, True).T[None, :, :, None]
        mag = stft.power_spectrum(X).astype(np.float32)       # TF feed casts to float32 (tester.py:69)
        phase = stft.divide_phase(X)
        pred = network.forward(arch, w, mag, np.float64).astype(np.float32)
        ref = rebuild.rebuild_audio([len(wv)], pred[..., 0], phase[..., 0], 8000, 32.0, 16.0)[0]
        snr = rebuild.sdr_db(ref, o)
        assert snr >= SNR_DB, (arch, len(wv), snr)


def test_identity_roundtrip_nfft256(dev, engines):
    """Known-answer test: K1 -> identity network -> K3 with
    irfft_n = 256 returns the input waveform (a physical check independent of the oracle)."""
    eng = engines["FullyCNNV2"][0]
    L = 32000
    wv = noisy_utterance(41, L)
    T = int(num_frames(L))
    d_wav, d_off, d_len, d_row, total = _upload_batch(dev, [wv], [T])
    mag, phase = eng.stft_device(d_wav, d_off, d_len, d_row, total)
    out = torch.zeros(L, dtype=torch.float32, device=dev)
    eng.istft_device(mag, phase, d_row, T, out, d_off, d_len, irfft_n=256)
    torch.cuda.synchronize()
    assert rebuild.sdr_db(wv.astype(np.float64), out.cpu().numpy()) > 80.0


def test_stream_chunked_equals_whole(dev, engines):
    """BASELINE config 4 (shortened): chunked enhancement with halo equals the un-chunked one."""
    eng = engines["FullyCNNV2"][0]
    L = 8000 * 60
    wv = noisy_utterance(55, L)
    whole = eng.enhance([wv])[0]
    chunked = eng.enhance_stream(wv, chunk_seconds=4.0)
    assert len(chunked) == L
    assert rebuild.sdr_db(whole, chunked) >= 100.0


def test_online_streaming_equals_whole(dev, engines):
    """Block-by-block enhancement (StreamingEnhancer: arbitrary push sizes, 13-hop look-back, 6-hop
    look-ahead) returns the whole-file result to float32 rounding, sample for sample."""
    from fullycnnspeechenhancement_b200.streaming import StreamingEnhancer
    eng = engines["FullyCNNV3"][0]
    L = 8000 * 20 + 77
    wv = noisy_utterance(56, L)
    whole = eng.enhance([wv])[0]
    st = StreamingEnhancer(eng, block=4096)
    rng = np.random.default_rng(1)
    outs, pos = [], 0
    while pos < L:
        n = int(rng.integers(100, 9000))
        outs.append(st.push(wv[pos:pos + n]))
        pos += n
    outs.append(st.flush())
    got = np.concatenate(outs)
    assert len(got) == L
    assert rebuild.sdr_db(whole, got) >= 100.0


def test_golden_fixtures_on_gpu(dev, engines, golden_dir):
    """Committed vectors: reference-generated STFT/rebuild, oracle-generated network outputs."""
    import os
    g = np.load(os.path.join(golden_dir, "stft_rebuild_ref.npz"))
    eng = engines["FullyCNNV2"][0]
    for seed, L in zip(g["case_seeds"], g["case_lengths"]):
        wv = g["wav_%d" % seed]
        T = int(num_frames(L))
        d_wav, d_off, d_len, d_row, total = _upload_batch(dev, [wv], [T])
        mag, phase = eng.stft_device(d_wav, d_off, d_len, d_row, total)
        torch.cuda.synchronize()
        ref = g["mag_%d" % seed].T
        assert mag.shape[0] == ref.shape[0]
        assert rel_err(mag.cpu().numpy(), ref) <= MAG_TOL
    # rebuild goldens: batch of the three longest cases, dense padded layout
    pb = g["padded_batch"][..., 0]                       # [3,Tmax,129] complex
    pred = g["pred"]
    lens = [int(x) for x in g["case_lengths"][3:]]
    N, Tmax = pb.shape[0], pb.shape[1]
    ph = np.exp(1j * np.angle(pb))
    ph32 = np.stack([ph.real, ph.imag], -1).astype(np.float32).reshape(N * Tmax, 129, 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    for nfft in (512, 256):
        out = torch.zeros(int(sum(lens)), dtype=torch.float32, device=dev)
        eng.istft_device(t(pred.reshape(N * Tmax, 129)), t(ph32), t(np.arange(N + 1, dtype=np.int64) * Tmax), Tmax,
                         out, t(out_off), t(np.array(lens, np.int32)), irfft_n=nfft)
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        for i, L in enumerate(lens):
            ref = g["rebuild%d_%d" % (nfft, i)]
            assert len(ref) == L
            assert rebuild.sdr_db(ref, o[out_off[i]:out_off[i] + L]) >= SNR_DB
    n = np.load(os.path.join(golden_dir, "network_oracle.npz"))
    for arch in ARCHS:
        e, w = engines[arch]
        for T in (1, 7, 8, 9, 12):
            x = n["x_%s_%d" % (arch, T)]
            y = n["y_%s_%d" % (arch, T)][..., 0].reshape(2 * T, 129)
            pred = e.forward_device(t(x.reshape(2 * T, 129)), torch.tensor([0, T, 2 * T], dtype=torch.int64, device=dev))
            torch.cuda.synchronize()
            assert rel_err(pred.cpu().numpy(), y) <= MAG_TOL
    # the same inputs through the reference's own model classes (tests/golden/make_golden.py, oracle/tf_standin.py)
    gm = np.load(os.path.join(golden_dir, "network_ref_model.npz"))
    for arch in ARCHS:
        e, w = engines[arch]
        for T in (1, 8, 12):
            x = n["x_%s_%d" % (arch, T)]
            y = gm["y_%s_%d" % (arch, T)][..., 0].reshape(2 * T, 129)
            pred = e.forward_device(t(x.reshape(2 * T, 129)), torch.tensor([0, T, 2 * T], dtype=torch.int64, device=dev))
            torch.cuda.synchronize()
            assert rel_err(pred.cpu().numpy(), y) <= MAG_TOL
