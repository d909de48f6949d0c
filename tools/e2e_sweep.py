"""End-to-end (host buffers) step time of the bench workload for several chunk sizes / stream counts of the host
pipeline: python tools/e2e_sweep.py  -> one line per setting (ms per 1024-utterance step over 10 queued steps)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fullycnnspeechenhancement_b200.engine import Enhancer            # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import fold           # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance      # noqa: E402

n_utt, L, steps = 1024, 32000, int(sys.argv[1]) if len(sys.argv) > 1 else 10
eng = Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2", seed=0), device=0)
pool = [noisy_utterance(1000 + i, L) for i in range(16)]
h_in = [torch.from_numpy(np.concatenate([pool[i % 16] for i in range(n_utt)])).pin_memory() for _ in range(2)]
h_out = [torch.empty_like(h_in[0]).pin_memory() for _ in range(2)]
t = eng.host_tables(np.full(n_utt, L))


def run(n):
    for i in range(n):
        eng.enhance_host(h_in[i & 1], h_out[i & 1], t, sync=False)
    eng.host_sync()


for streams in (3,):   # (copy-in, compute, copy-out: fixed)
    for rows in ((32768, 131072) if os.environ.get("E2E_QUICK") else (32768, 65536, 87381, 131072, 180000, 262144)):
        eng.host_config(chunk_rows=rows)
        run(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(steps)
        dt = time.perf_counter() - t0
        print("streams %d chunk_rows %6d: %.3f ms/step  (%.1f k audio-s/s)" % (streams, rows, 1e3 * dt / steps, n_utt * 4 * steps / dt / 1e3), flush=True)
