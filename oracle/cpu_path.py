"""The reference's enhancement path on CPU, end to end (test oracle + CPU baseline leg).

Mirrors the batch loop of FullyCNNTester.test (/root/reference/model_utils/tester.py:92-113):
per-utterance numpy STFT (the reference does this in its DataSet, data_loader.py:54-61),
padding_batch, power_spectrum / divide_phase, the network, rebuild_audio.  TensorFlow 1.14 is
not installable, so the network leg is torch-CPU float32 conv2d (oracle.network.forward_torch)
using every host thread -- the closest available stand-in for TF's Eigen/MKL CPU kernels, which
is what the reference's own launcher uses (Work/*/run_test.sh:5 sets CUDA_VISIBLE_DEVICES='').
"""
import time

import numpy as np

from . import network, rebuild, stft


def enhance_batch_cpu(waves, net_work, weights, sample_rate=8000, window_ms=32, stride_ms=16,
                      nfft_rebuild=512, faithful_loop=True, net_dtype="float32", timings=None):
    t0 = time.perf_counter()
    specs = [stft.compute_spectrogram(w, sample_rate, window_ms / 1000, stride_ms / 1000, 256, True) for w in waves]
    batch = stft.padding_batch(specs)                       # [N,T,129,1] complex128
    mag = stft.power_spectrum(batch)
    phase = stft.divide_phase(batch)
    t1 = time.perf_counter()
    if net_dtype == "float64":
        pred = network.forward(net_work, weights, mag.astype(np.float32), np.float64).astype(np.float32)
    else:
        pred = network.forward_torch(net_work, weights, mag.astype(np.float32), "float32")
    t2 = time.perf_counter()
    out = rebuild.rebuild_audio([len(w) for w in waves], pred.squeeze(-1), phase.squeeze(-1), sample_rate,
                                float(window_ms), float(stride_ms), nfft=nfft_rebuild, faithful_loop=faithful_loop)
    t3 = time.perf_counter()
    if timings is not None:
        timings["stft"] = timings.get("stft", 0.0) + (t1 - t0)
        timings["network"] = timings.get("network", 0.0) + (t2 - t1)
        timings["rebuild"] = timings.get("rebuild", 0.0) + (t3 - t2)
    return out
