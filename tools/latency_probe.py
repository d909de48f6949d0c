"""One 4 s utterance and one streaming block through the host API, a few times: target of
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/latency_probe.py
to see which kernels make up the host-to-host latency (bench.py reports the latency itself)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fullycnnspeechenhancement_b200.engine import Enhancer            # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import fold           # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance      # noqa: E402

eng = Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2", seed=0), device=0)
for n in (32000, 2944, 6528):
    w = noisy_utterance(5, n)
    for _ in range(3):
        eng.enhance([w])
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        eng.enhance([w])
        ts.append(time.perf_counter() - t0)
    print("samples %6d: host-to-host %.1f us (median of 20)" % (n, 1e6 * float(np.median(ts))), flush=True)
