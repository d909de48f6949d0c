"""Short-time objective intelligibility (STOI) on the host, in numpy / scipy.

The reference scores enhanced utterances with ``pystoi.stoi(clean, denoised, sample_rate)``
(model_utils/utils.py:48-62, called at model_utils/tester.py:136-140).  pystoi is a third-party
package that is neither vendored in the reference nor installed here (requriements.txt lists it
without a version), so this module restates the published algorithm it implements -- C. H. Taal,
R. C. Hendriks, R. Heusdens, J. Jensen, "An Algorithm for Intelligibility Prediction of
Time-Frequency Weighted Noisy Speech", IEEE TASLP 2011 -- with pystoi's constants:

  1. resample both signals to 10 kHz (polyphase filter, Kaiser-windowed sinc as in Octave's resample),
  2. drop the frames (256 samples, hop 128, Hann) whose CLEAN energy is more than 40 dB below the
     loudest clean frame, overlap-add the rest back into signals,
  3. 512-point spectra of Hann-windowed 256-sample frames with hop 128, grouped into 15
     one-third-octave bands from 150 Hz,
  4. for every band and every run of N = 30 frames (384 ms): scale the degraded envelope to the clean
     one's energy, clip it at -15 dB signal-to-distortion, and correlate the two envelopes,
  5. the score is the mean correlation over bands and runs.

Only the non-extended measure is implemented (the reference calls stoi() with its defaults)."""
import numpy as np
from scipy.signal import resample_poly

FS = 10000          # sample rate the measure is defined at
N_FRAME = 256       # window
NFFT = 512
NUMBAND = 15        # one-third-octave bands
MINFREQ = 150       # centre frequency of the first band
N = 30              # frames per intermediate-intelligibility segment
BETA = -15.0        # lower signal-to-distortion bound (dB)
DYN_RANGE = 40      # speech dynamic range (dB)
EPS = np.finfo("float").eps


def _resample_filter(p, q):
    """Kaiser-windowed sinc of Octave's resample(): 60 dB rejection, roll-off a tenth of the cut-off."""
    g = np.gcd(p, q)
    p, q = p // g, q // g
    cutoff = 1.0 / (2 * max(p, q))
    rolloff = cutoff / 10.0
    rejection_db = 60.0
    half = int(np.ceil((rejection_db - 8.0) / (28.714 * rolloff)))
    t = np.arange(-half, half + 1)
    ideal = 2 * p * cutoff * np.sinc(2 * cutoff * t)
    beta = 0.1102 * (rejection_db - 8.7)
    return np.kaiser(2 * half + 1, beta) * ideal


def resample(x, fs_to, fs_from):
    if fs_to == fs_from:
        return np.asarray(x, dtype=np.float64)
    h = _resample_filter(fs_to, fs_from)
    return resample_poly(np.asarray(x, dtype=np.float64), fs_to, fs_from, window=h / h.sum())


def _frames(x, framelen, hop):
    """Rows x[i : i + framelen] for i = 0, hop, ... < len(x) - framelen (the last complete frame is NOT taken when it ends
    exactly at the end of the signal: range(0, len - framelen, hop), as in pystoi)."""
    starts = np.arange(0, len(x) - framelen, hop)
    if len(starts) == 0:
        return np.zeros((0, framelen))
    return x[starts[:, None] + np.arange(framelen)[None, :]]


def _hann(framelen):
    return np.hanning(framelen + 2)[1:-1]


def remove_silent_frames(x, y, dyn_range=DYN_RANGE, framelen=N_FRAME, hop=N_FRAME // 2):
    w = _hann(framelen)
    xf = _frames(x, framelen, hop) * w
    yf = _frames(y, framelen, hop) * w
    if len(xf) == 0:
        return np.zeros(0), np.zeros(0)
    energies = 20 * np.log10(np.linalg.norm(xf, axis=1) + EPS)
    keep = (np.max(energies) - dyn_range - energies) < 0
    xf, yf = xf[keep], yf[keep]

    def overlap_add(f):
        out = np.zeros((len(f) - 1) * hop + framelen) if len(f) else np.zeros(0)
        for i in range(len(f)):
            out[i * hop:i * hop + framelen] += f[i]
        return out

    return overlap_add(xf), overlap_add(yf)


def third_octave_matrix(fs=FS, nfft=NFFT, num_bands=NUMBAND, min_freq=MINFREQ):
    f = np.linspace(0, fs, nfft + 1)[:nfft // 2 + 1]
    k = np.arange(num_bands, dtype=np.float64)
    lo = min_freq * np.power(2.0, (2 * k - 1) / 6)
    hi = min_freq * np.power(2.0, (2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    for i in range(num_bands):
        a = int(np.argmin(np.square(f - lo[i])))
        b = int(np.argmin(np.square(f - hi[i])))
        obm[i, a:b] = 1
    return obm


def stoi(x, y, fs_sig, extended=False):
    """STOI of the degraded / processed signal ``y`` against the clean signal ``x`` (same length), sampled at
    ``fs_sig`` Hz.  Returns a float, higher is better (about 0 .. 1)."""
    if extended:
        raise NotImplementedError("the reference uses the non-extended measure")
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if x.shape != y.shape:
        raise Exception("x and y should have the same length, found {} and {}".format(x.shape, y.shape))
    x = resample(x, FS, int(fs_sig))
    y = resample(y, FS, int(fs_sig))
    x, y = remove_silent_frames(x, y)
    w = _hann(N_FRAME)
    xs = np.fft.rfft(_frames(x, N_FRAME, N_FRAME // 2) * w, n=NFFT).T     # [bins, frames]
    ys = np.fft.rfft(_frames(y, N_FRAME, N_FRAME // 2) * w, n=NFFT).T
    if xs.shape[-1] < N:
        # pystoi warns ("Not enough STFT frames to compute intermediate intelligibility measure after removing silent
        # frames") and returns 1e-5
        return 1e-5
    obm = third_octave_matrix()
    xt = np.sqrt(obm @ np.square(np.abs(xs)))                               # [bands, frames]
    yt = np.sqrt(obm @ np.square(np.abs(ys)))
    m = xt.shape[1] - N + 1
    idx = np.arange(N)[None, :] + np.arange(m)[:, None]                     # [segments, N]
    xseg = xt[:, idx].transpose(1, 0, 2)                                    # [segments, bands, N]
    yseg = yt[:, idx].transpose(1, 0, 2)
    norm = np.linalg.norm(xseg, axis=2, keepdims=True) / (np.linalg.norm(yseg, axis=2, keepdims=True) + EPS)
    yn = yseg * norm
    clip = 10 ** (-BETA / 20)
    yp = np.minimum(yn, xseg * (1 + clip))
    yp = yp - yp.mean(axis=2, keepdims=True)
    xc = xseg - xseg.mean(axis=2, keepdims=True)
    yp = yp / (np.linalg.norm(yp, axis=2, keepdims=True) + EPS)
    xc = xc / (np.linalg.norm(xc, axis=2, keepdims=True) + EPS)
    return float(np.sum(yp * xc) / (xseg.shape[0] * xseg.shape[1]))
