"""Times K1 (STFT) and K3 (reconstruction) alone on BASELINE.json configs[1] shapes (1024 x 4 s) with CUDA events:
    python tools/k13_time.py [reps]
(development aid for the two HBM-side kernels; the bench line's roofline_k1 / roofline_k3 are the record)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fullycnnspeechenhancement_b200.engine import Enhancer, num_frames      # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import fold                 # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance            # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_utt, L = 1024, 32000
eng = Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2", seed=0), device=0)
pool = [noisy_utterance(1000 + i, L) for i in range(16)]
wav = torch.from_numpy(np.concatenate([pool[i % 16] for i in range(n_utt)])).cuda()
out = torch.empty_like(wav)
plan = eng.plan(np.full(n_utt, L))
T = int(num_frames(L))
rows = n_utt * T
ro = plan["row_off_all"]
mag = torch.empty((rows, 129), device="cuda")
phase = torch.empty((rows, 129, 2), device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return t[len(t) // 2], t[0]


k1 = timed(lambda: eng.stft_device(wav, plan["wav_off"], plan["wav_len"], ro, rows, mag, phase))
k3 = timed(lambda: eng.istft_device(mag, phase, ro, T, out, plan["wav_off"], plan["wav_len"]))
print("K1 median %.4f ms (min %.4f)   K3 median %.4f ms (min %.4f)   checksum %.6e"
      % (k1[0], k1[1], k3[0], k3[1], float(out.double().abs().sum())))
