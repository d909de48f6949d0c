"""AudioFeature with the reference's interface (data_utils/audio_feature.py:12-115), computed by
the STFT kernel of librced_b200.so.  Arrays come back as numpy in the reference's layout
([F, T], complex128 / float32) so existing callers keep working; values are float32-accurate.

Only the configuration the reference actually runs is implemented on the GPU: Hamming window
(audio_feature.py:13-20 with windows_name=None -- no call site passes anything else), 256-sample
frames, 128-sample hop, nfft 256 (data_loader.py:59).  Anything else raises NotImplementedError
instead of silently taking another path."""
import ctypes

import numpy as np
import torch

from .. import _lib, runtime

_SUPPORTED_WINDOWS = (None, "hamming")


class AudioFeature(object):
    def __init__(self, windows_name=None, device=None):
        if windows_name not in _SUPPORTED_WINDOWS:
            raise NotImplementedError("only the Hamming window (the reference's effective default) has a CUDA path")
        self.window = np.hamming
        self.device_index = runtime.default_device() if device is None else int(device)
        self._eng = None

    # the STFT kernel needs a library handle; it does not depend on any network weights
    def _engine(self):
        if self._eng is None:
            from ..engine import Enhancer
            n = _lib.lib().rced_folded_weight_count(2)
            self._eng = Enhancer("FullyCNNV2", np.zeros(n, np.float32), device=self.device_index, variant="ffma")
        return self._eng

    def compute_spectrogram(self, signal, sample_rate, window_s=0.02, stride_s=0.01, nfft=512, use_complex=False):
        if stride_s > window_s:
            raise ValueError("Stride size must not be greater than window size.")
        frame_length = int(round(window_s * sample_rate))
        frame_step = int(round(stride_s * sample_rate))
        if (frame_length, frame_step, nfft) != (256, 128, 256):
            raise NotImplementedError("the CUDA STFT implements the reference's shipped configuration only: "
                                      "256-sample window, 128-sample stride, nfft 256 (got %d/%d/%d)"
                                      % (frame_length, frame_step, nfft))
        spec = self.spectrogram_batch([signal])[0]
        if use_complex:
            return spec
        return self._mag_of(spec)

    def spectrogram_batch(self, signals):
        """STFT of several signals in one launch; returns a list of [129, T_i] complex128 arrays."""
        eng = self._engine()
        dev = eng.device
        lens = np.array([len(s) for s in signals], dtype=np.int64)
        if np.any(lens < 1):
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # audio_feature.py:54 on empty input
        frames = np.array([_lib.num_frames(n) for n in lens], dtype=np.int64)
        wav = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.float32) for s in signals])).to(dev)
        wav_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)).to(dev)
        wav_len = torch.from_numpy(lens.astype(np.int32)).to(dev)
        row_off_h = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        mag, phase = eng.stft_device(wav, wav_off, wav_len, torch.from_numpy(row_off_h).to(dev), int(row_off_h[-1]))
        torch.cuda.synchronize(dev)
        m = mag.cpu().numpy().astype(np.float64)
        p = phase.cpu().numpy().astype(np.float64)
        X = m * (p[..., 0] + 1j * p[..., 1])
        return [np.transpose(X[row_off_h[i]:row_off_h[i + 1]]) for i in range(len(signals))]

    def _mag_of(self, spec):
        return self.power_spectrum(spec).astype(np.float32)

    def _mag_phase_gpu(self, frames, want_mag, want_phase):
        x = np.ascontiguousarray(np.asarray(frames, dtype=np.complex64))
        dev = torch.device("cuda", self.device_index)
        d = torch.from_numpy(x.view(np.float32).reshape(-1)).to(dev)
        n = x.size
        mag = torch.empty(n, dtype=torch.float32, device=dev) if want_mag else None
        ph = torch.empty(2 * n, dtype=torch.float32, device=dev) if want_phase else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().rced_mag_phase(self.device_index, ctypes.c_void_p(d.data_ptr()), n,
                                             ctypes.c_void_p(mag.data_ptr()) if want_mag else None,
                                             ctypes.c_void_p(ph.data_ptr()) if want_phase else None, stream))
        torch.cuda.synchronize(dev)
        out_mag = mag.cpu().numpy().astype(np.float64).reshape(x.shape) if want_mag else None
        out_ph = ph.cpu().numpy().view(np.complex64).astype(np.complex128).reshape(x.shape) if want_phase else None
        return out_mag, out_ph

    def power_spectrum(self, frames):
        """|X| (linear magnitude; audio_feature.py:101-110)."""
        return self._mag_phase_gpu(frames, True, False)[0]

    def divide_phase(self, fft_frames):
        """exp(j*angle(X)) = X/|X|, 1+0j where X == 0 (audio_feature.py:112-115)."""
        return self._mag_phase_gpu(fft_frames, False, True)[1]
