"""Minimal wav file I/O standing in for librosa.load / soundfile.write (neither is installed).

``load_wav`` returns float32 mono in [-1, 1] at the requested rate (polyphase resampling when
the file's rate differs -- librosa would use resampy's kaiser_best; results differ slightly in
that case).  ``write_wav`` writes 16-bit PCM, soundfile's default subtype for float input."""
import numpy as np
from scipy.io import wavfile
from scipy.signal import resample_poly


def load_wav(path, sample_rate):
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    if sample_rate is not None and sr != sample_rate:
        g = np.gcd(int(sr), int(sample_rate))
        x = resample_poly(x, sample_rate // g, sr // g).astype(np.float32)
        sr = sample_rate
    return np.ascontiguousarray(x, dtype=np.float32), sr


def write_wav(path, data, sample_rate):
    x = np.clip(np.asarray(data, dtype=np.float64), -1.0, 1.0 - 1.0 / 32768.0)
    wavfile.write(path, int(sample_rate), np.round(x * 32768.0).astype(np.int16))
