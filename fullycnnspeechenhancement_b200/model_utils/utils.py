"""AudioReBuild and the scoring helpers with the reference's interface
(model_utils/utils.py of the reference).  rebuild_audio runs the reconstruction kernel (K3)."""
import numpy as np
import torch

from .. import runtime


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val          # the reference adds val once whatever n is (utils.py:26-30)
        self.count += n
        self.avg = self.sum / self.count


class SDR(object):
    """10*log10(||y||^2 / (||y_pred - y||^2 + eps_f32)) (utils.py:64-90 of the reference)."""

    def sdr(self, y, y_pred):
        assert len(y.shape) == 1
        assert len(y) == len(y_pred)
        y_en = np.power(y, 2).sum()
        err_en = np.power(y_pred - y, 2).sum()
        return 10 * np.log10(y_en / (err_en + np.finfo(np.float32).eps))

    def __call__(self, x, y):
        return self.sdr(x, y)


def sdr_batch(refs, ests, device=None):
    """SDR of several (reference, estimate) pairs in one launch of the library's energy-sum kernel;
    the same formula as ``SDR.sdr`` with the sums taken in float64 on the GPU.  Inputs are converted to
    float32 (what the enhancement path produces); returns a float64 array."""
    import ctypes

    from .. import _lib
    dev_index = runtime.default_device() if device is None else int(device)
    dev = torch.device("cuda", dev_index)
    n = len(refs)
    assert n == len(ests)
    lens = np.array([len(r) for r in refs], dtype=np.int64)
    for r, e in zip(refs, ests):
        assert len(r) == len(e)
    if n == 0:
        return np.zeros(0)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    up = lambda xs: torch.from_numpy(np.concatenate([np.asarray(x, dtype=np.float32) for x in xs])).to(dev)
    d_ref, d_est = up(refs), up(ests)
    d_off = torch.from_numpy(off).to(dev)
    d_len = torch.from_numpy(lens.astype(np.int32)).to(dev)
    sums = torch.empty((n, 2), dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().rced_sdr_sums(dev_index, p(d_ref), p(d_off), p(d_est), p(d_off), p(d_len), n, int(lens.max()),
                                        p(sums), stream))
    s = sums.cpu().numpy()
    return 10 * np.log10(s[:, 0] / (s[:, 1] + np.finfo(np.float32).eps))


class PESQ(object):
    """The reference scores with ``pypesq.pesq`` (model_utils/utils.py:32-45): the ITU-T P.862 C implementation behind a
    Python wrapper.  Neither the wrapper nor the ITU code is vendored in the reference or installed here, and the
    standard's reference code cannot be restated from memory, so the score is reported as unavailable (``nan``) --
    never as a number."""
    available = False

    def __init__(self, sr=16000):
        self.sr = sr

    def __call__(self, a, b):
        assert len(a.shape) == 1
        assert len(a) == len(b)
        return float("nan")


class STOI(object):
    """``pystoi.stoi(clean, processed, sr)`` of the reference (model_utils/utils.py:48-62), restated in numpy
    (model_utils/stoi.py; pystoi itself is not installed)."""
    available = True

    def __init__(self, sr=16000):
        self.sr = sr

    def __call__(self, a, b):
        from .stoi import stoi
        assert len(a.shape) == 1
        assert len(a) == len(b)
        return stoi(a, b, self.sr)


class AudioReBuild(object):
    def __init__(self, windows_name=None, nfft=512, device=None):
        if windows_name not in (None, "hamming"):
            raise NotImplementedError("only the Hamming window has a CUDA path")
        if nfft not in (512, 256):
            raise NotImplementedError("irfft length must be 512 (the reference's default) or 256")
        self.nfft = nfft
        self.device_index = runtime.default_device() if device is None else int(device)
        self._eng = None

    def _engine(self):
        if self._eng is None:
            from .. import _lib
            from ..engine import Enhancer
            n = _lib.lib().rced_folded_weight_count(2)
            self._eng = Enhancer("FullyCNNV2", np.zeros(n, np.float32), device=self.device_index, variant="ffma")
        return self._eng

    def rebuild_audio(self, sig_length_list, spec, phase, sample_rate, windows_ms, stride_ms):
        """spec [N,T,F] real, phase [N,T,F] complex -> list of N float64 arrays truncated to
        sig_length_list (utils.py:171-183 of the reference)."""
        n_window = int((windows_ms * sample_rate) / 1000)
        n_stride = int((stride_ms * sample_rate) / 1000)
        if (n_window, n_stride) != (256, 128):
            raise NotImplementedError("the CUDA reconstruction implements 256-sample frames with a 128-sample hop")
        spec = np.asarray(spec)
        phase = np.asarray(phase)
        if spec.ndim != 3 or spec.shape[2] != 129 or phase.shape != spec.shape:
            raise ValueError("spec and phase must both be [N, T, 129]")
        N, T = spec.shape[0], spec.shape[1]
        lens = np.asarray(sig_length_list, dtype=np.int64)
        assert len(lens) == N
        keep = np.minimum(lens, (T + 1) * 128)            # numpy slicing [:L] never extends
        eng = self._engine()
        dev = eng.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        ph = np.stack([phase.real, phase.imag], axis=-1).astype(np.float32).reshape(N * T, 129, 2)
        out_off = np.concatenate([[0], np.cumsum(keep)[:-1]]).astype(np.int64)
        out = torch.zeros(int(keep.sum()) + 1, dtype=torch.float32, device=dev)
        eng.istft_device(t(spec.astype(np.float32).reshape(N * T, 129)), t(ph),
                         torch.arange(N + 1, dtype=torch.int64, device=dev) * T, T, out, t(out_off),
                         t(keep.astype(np.int32)), irfft_n=self.nfft)
        torch.cuda.synchronize(dev)
        o = out.cpu().numpy().astype(np.float64)
        return [o[out_off[i]:out_off[i] + keep[i]] for i in range(N)]
