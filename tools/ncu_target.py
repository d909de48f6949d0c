"""Target of the ncu captures: a few passes of K1 -> K2 -> K3 on BASELINE.json configs[1] shapes (1024 x 4 s, V2),
nothing else (no CPU baseline, no host pipeline), so that `ncu -k regex:... -s N -c 1` finds its kernels quickly.
    ncu --set full --clock-control none --import-source on -k regex:'rced_(stft|net_tc|istft)_kernel' -s 3 -c 3 \
        -o gpurun_out/r02_path python tools/ncu_target.py [tc|ffma] [passes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fullycnnspeechenhancement_b200.engine import Enhancer, num_frames      # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import fold                 # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance            # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "tc"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_utt, L = 1024, 32000
eng = Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2", seed=0), device=0, variant=variant)
pool = [noisy_utterance(1000 + i, L) for i in range(16)]
wav = torch.from_numpy(np.concatenate([pool[i % 16] for i in range(n_utt)])).cuda()
out = torch.empty_like(wav)
plan = eng.plan(np.full(n_utt, L))
T = int(num_frames(L))
rows = n_utt * T
ro = plan["row_off_all"]
mag = torch.empty((rows, 129), device="cuda")
phase = torch.empty((rows, 129, 2), device="cuda")
pred = torch.empty((rows, 129), device="cuda")
for _ in range(passes):
    eng.stft_device(wav, plan["wav_off"], plan["wav_len"], ro, rows, mag, phase)
    eng.forward_device(mag, ro, pred)
    eng.istft_device(pred, phase, ro, T, out, plan["wav_off"], plan["wav_len"])
torch.cuda.synchronize()
print("ncu target done:", variant, passes, "passes; checksum", float(out.double().abs().sum()))
