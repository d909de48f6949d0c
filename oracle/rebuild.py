"""float64 numpy restatement of the reference waveform reconstruction (test oracle).

Follows /root/reference/model_utils/utils.py:93-183 (class AudioReBuild).
"""
import numpy as np
from scipy.signal import lfilter

PRE_EMPHASIS = 0.97  # utils.py:106


def de_emphasis_loop(signal):
    """utils.py:104-113 verbatim in structure: a per-sample python loop, float64."""
    out = np.empty_like(np.asarray(signal, dtype=np.float64))
    for r, sig in enumerate(signal):
        acc = [sig[0]]
        for i in range(1, len(sig)):
            acc.append(sig[i] + acc[i - 1] * PRE_EMPHASIS)
        out[r] = np.array(acc)
    return out


def de_emphasis(signal):
    """Same recurrence y[i] = x[i] + 0.97*y[i-1] through scipy's direct-form filter
    (identical float64 operation order; checked bit-exact against the loop in
    tests/test_oracle_golden.py)."""
    return lfilter([1.0], [1.0, -PRE_EMPHASIS], np.asarray(signal, dtype=np.float64), axis=1)


def rebuild_audio(sig_length_list, spec, phase, sample_rate=8000, windows_ms=32, stride_ms=16,
                  nfft=512, window=np.hamming, faithful_loop=False):
    """utils.py:171-183.  spec [N,T,F] real, phase [N,T,F] complex.

    nfft=512 is the shipped default (utils.py:94; constructed without arguments at
    infer.py:34, tester.py:93).  Returns a list of float64 arrays of the given lengths.
    """
    n_window = int((windows_ms * sample_rate) / 1000)
    n_stride = int((stride_ms * sample_rate) / 1000)
    n_overlap = n_window - n_stride
    stft = spec * phase                                            # utils.py:119-126
    frames = np.fft.irfft(stft, nfft)[:, :, :n_window]             # utils.py:115-117,176
    frames = frames / window(n_window)                             # utils.py:128-137
    main = frames[:, :, n_overlap:].reshape(frames.shape[0], -1)   # utils.py:139-147
    sig = np.append(frames[:, 0, :n_overlap], main, axis=1)
    sig = de_emphasis_loop(sig) if faithful_loop else de_emphasis(sig)
    return [sig[i][:sig_length_list[i]] for i in range(len(sig))]


def sdr_db(ref, est):
    """utils.py:76-78 (class SDR): 10*log10(sum(ref^2) / sum((est-ref)^2))."""
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    den = np.sum((est - ref) ** 2)
    if den == 0:
        return np.inf
    return 10.0 * np.log10(np.sum(ref ** 2) / den)
