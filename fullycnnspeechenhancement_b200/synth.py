"""Synthetic 8 kHz noisy-speech generator (SURVEY.md section 8d "Synthetic signal").

Speech-like voices are a few chirped harmonics of a random f0 under a slow amplitude
envelope plus one or two pure tones; noise is white Gaussian plus a "babble" of six
such voices, mixed at the requested SNR with the reference's own mixing formula
(data_utils/data_loader.py:47-51).  Output is float32 peak-normalised to 0.9, the
dtype ``librosa.load`` hands the reference.
"""
import numpy as np


def _voice(rng, n, sr):
    t = np.arange(n, dtype=np.float64) / sr
    dur = max(n / sr, 1e-3)
    f0 = rng.uniform(80.0, 300.0)
    chirp = rng.uniform(-0.2, 0.2)
    inst = f0 * (1.0 + chirp * t / dur)
    phase = 2.0 * np.pi * np.cumsum(inst) / sr
    env = 0.5 * (1.0 + np.sin(2.0 * np.pi * rng.uniform(3.0, 5.0) * t + rng.uniform(0, 2 * np.pi)))
    sig = np.zeros(n)
    for h in range(1, int(rng.integers(3, 7)) + 1):
        sig += np.sin(h * phase + rng.uniform(0, 2 * np.pi)) / h
    sig *= env
    for _ in range(int(rng.integers(1, 3))):
        sig += 0.3 * np.sin(2.0 * np.pi * rng.uniform(300.0, 3400.0) * t + rng.uniform(0, 2 * np.pi))
    return sig


def noisy_utterance(seed, n_samples, sample_rate=8000, snr_db=None, return_clean=False):
    rng = np.random.default_rng(seed)
    speech = _voice(rng, n_samples, sample_rate)
    noise = rng.normal(0.0, 1.0, n_samples)
    babble = np.zeros(n_samples)
    for _ in range(6):
        babble += _voice(rng, n_samples, sample_rate)
    noise = noise / np.sqrt(np.mean(noise ** 2)) + babble / np.sqrt(np.mean(babble ** 2) + 1e-12)
    if snr_db is None:
        snr_db = float(rng.choice([0.0, 5.0, 10.0]))
    p_sig = np.sum(np.abs(speech) ** 2)                       # data_loader.py:47-51
    background_volume = p_sig / (10 ** (snr_db / 10))
    p_back = np.sum(np.abs(noise) ** 2)
    mix = speech + np.sqrt(background_volume / p_back) * noise
    scale = 0.9 / max(np.max(np.abs(mix)), 1e-12)
    mix32 = (mix * scale).astype(np.float32)
    if return_clean:
        return mix32, (speech * scale).astype(np.float32)
    return mix32


def noisy_batch(base_seed, lengths, sample_rate=8000):
    """One float32 waveform per entry of ``lengths``; per-utterance seed = base_seed + index."""
    return [noisy_utterance(base_seed + i, int(n), sample_rate) for i, n in enumerate(lengths)]
