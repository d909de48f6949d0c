"""First-light script for a GPU box: self-tests and a quick timing of the three kernels."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fullycnnspeechenhancement_b200 import _lib
from fullycnnspeechenhancement_b200.engine import Enhancer, num_frames
from fullycnnspeechenhancement_b200.model_utils import fold
from fullycnnspeechenhancement_b200.synth import noisy_utterance

L = _lib.lib()
print("device", torch.cuda.get_device_name(0))
print("tmem selftest rc", L.rced_selftest_tmem(0), L.rced_last_error())
tf = ctypes.c_double()
print("ffma peak rc", L.rced_ffma_peak(0, 4096, ctypes.byref(tf)), "TFLOP/s", tf.value)
for name in sys.argv[1:] or ["FullyCNNV2"]:
    eng = Enhancer(name, fold.glorot_weights(name, 0))
    n_utt, Ls = 1024, 32000
    base = [noisy_utterance(i, Ls) for i in range(16)]
    wav = torch.from_numpy(np.concatenate([base[i % 16] for i in range(n_utt)])).cuda()
    plan = eng.plan(np.full(n_utt, Ls))
    out = torch.empty_like(wav)
    T = int(num_frames(Ls)); rows = n_utt * T
    mag = torch.empty((rows, 129), device="cuda"); ph = torch.empty((rows, 129, 2), device="cuda"); pred = torch.empty_like(mag)
    row_off = plan["row_off_all"]
    for tm in (True, False):
        eng.set_skip_in_tmem(tm)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for it in range(3):
            ev[0].record()
            eng.stft_device(wav, plan["wav_off"], plan["wav_len"], row_off, rows, mag, ph)
            ev[1].record()
            eng.forward_device(mag, row_off, pred)
            ev[2].record()
            eng.istft_device(pred, ph, row_off, T, out, plan["wav_off"], plan["wav_len"])
            ev[3].record()
            torch.cuda.synchronize()
        k1, k2, k3 = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
        flops = 2.0 * L.rced_mac_per_frame(eng.arch, 1) * rows
        print("%s tmem=%s  K1 %.3f ms (%.0f GB/s)  K2 %.3f ms (%.2f TFLOP/s valid, %.1f%% of measured FFMA peak)  K3 %.3f ms (%.0f GB/s)  -> %.0f audio-s/s"
              % (name, tm, k1, rows * 2060 / k1 / 1e6, k2, flops / k2 / 1e9, 100 * flops / k2 / 1e9 / tf.value, k3, rows * 2060 / k3 / 1e6,
                 n_utt * Ls / 8000 / ((k1 + k2 + k3) / 1e3)))
    eng.close()
