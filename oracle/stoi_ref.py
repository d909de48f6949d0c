"""TEST INFRASTRUCTURE: a second, loop-by-loop restatement of STOI (Taal et al. 2011, with the constants
and framing conventions of the `pystoi` package the reference imports at model_utils/utils.py:10 and
calls at :61) used to check fullycnnspeechenhancement_b200/model_utils/stoi.py.  pystoi is a third-party
dependency that is neither vendored in /root/reference nor installed here (requriements.txt: `pystoi`,
no version pin), so this file follows the published algorithm; parity with pystoi itself is unpinned.

Written without looking at the product module's vectorised code: frames, bands and segments are plain
Python loops, the resampler builds its own polyphase filter and applies it by direct convolution."""
import math

import numpy as np

FS, N_FRAME, NFFT, NUMBAND, MINFREQ, N_SEG, BETA, DYN_RANGE = 10000, 256, 512, 15, 150, 30, -15.0, 40
EPS = np.finfo("float").eps


def _kaiser_sinc(p, q):
    g = math.gcd(p, q)
    p, q = p // g, q // g
    fc = 1.0 / (2 * max(p, q))
    half = int(math.ceil((60.0 - 8.0) / (28.714 * fc / 10.0)))
    t = np.arange(-half, half + 1)
    h = np.kaiser(2 * half + 1, 0.1102 * (60.0 - 8.7)) * (2 * p * fc * np.sinc(2 * fc * t))
    return h / h.sum(), p, q


def resample(x, fs_to, fs_from):
    """Polyphase resampling by p/q as zero-stuffing + FIR + decimation (the definition scipy.signal.resample_poly
    implements; output length ceil(len * p / q), filter centred)."""
    if fs_to == fs_from:
        return np.asarray(x, dtype=np.float64)
    h, p, q = _kaiser_sinc(fs_to, fs_from)
    h = h * p
    x = np.asarray(x, dtype=np.float64)
    up = np.zeros(len(x) * p)
    up[::p] = x
    full = np.convolve(up, h)
    half = (len(h) - 1) // 2
    n_out = -(-len(x) * p // q)
    return full[half:half + n_out * q:q][:n_out]


def _hann(n):
    return np.hanning(n + 2)[1:-1]


def stoi(x, y, fs):
    x = resample(x, FS, fs)
    y = resample(y, FS, fs)
    w = _hann(N_FRAME)
    hop = N_FRAME // 2
    # 1. drop frames whose clean energy is more than 40 dB below the loudest clean frame
    starts = list(range(0, len(x) - N_FRAME, hop))
    xf = [w * x[i:i + N_FRAME] for i in starts]
    yf = [w * y[i:i + N_FRAME] for i in starts]
    en = [20 * math.log10(np.linalg.norm(f) + EPS) for f in xf]
    keep = [i for i in range(len(xf)) if (max(en) - DYN_RANGE - en[i]) < 0]
    xs = np.zeros((len(keep) - 1) * hop + N_FRAME)
    ys = np.zeros_like(xs)
    for j, i in enumerate(keep):
        xs[j * hop:j * hop + N_FRAME] += xf[i]
        ys[j * hop:j * hop + N_FRAME] += yf[i]
    # 2. one-third-octave band envelopes
    f = np.linspace(0, FS, NFFT + 1)[:NFFT // 2 + 1]
    bands = []
    for k in range(NUMBAND):
        lo = MINFREQ * 2.0 ** ((2 * k - 1) / 6.0)
        hi = MINFREQ * 2.0 ** ((2 * k + 1) / 6.0)
        bands.append((int(np.argmin((f - lo) ** 2)), int(np.argmin((f - hi) ** 2))))
    starts = list(range(0, len(xs) - N_FRAME, hop))
    if len(starts) < N_SEG:
        return 1e-5
    X = np.zeros((NUMBAND, len(starts)))
    Y = np.zeros((NUMBAND, len(starts)))
    for m, i in enumerate(starts):
        sx = np.abs(np.fft.rfft(w * xs[i:i + N_FRAME], NFFT)) ** 2
        sy = np.abs(np.fft.rfft(w * ys[i:i + N_FRAME], NFFT)) ** 2
        for k, (a, b) in enumerate(bands):
            X[k, m] = math.sqrt(sx[a:b].sum())
            Y[k, m] = math.sqrt(sy[a:b].sum())
    # 3. intermediate intelligibility of every band and 30-frame segment
    clip = 10 ** (-BETA / 20)
    total, count = 0.0, 0
    for m in range(N_SEG, len(starts) + 1):
        for k in range(NUMBAND):
            xv = X[k, m - N_SEG:m]
            yv = Y[k, m - N_SEG:m]
            yv = yv * (np.linalg.norm(xv) / (np.linalg.norm(yv) + EPS))
            yv = np.minimum(yv, xv * (1 + clip))
            yv = yv - yv.mean()
            xv = xv - xv.mean()
            total += float(np.dot(yv / (np.linalg.norm(yv) + EPS), xv / (np.linalg.norm(xv) + EPS)))
            count += 1
    return total / count
