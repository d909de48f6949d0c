"""Stress of the host-buffer entry points: random batches (1..120 utterances of 1..20,000 samples), random mix of synchronous
and queued asynchronous calls, random chunk sizes, buffers of several calls in flight at once; every output is compared with
the same utterance enhanced alone (>= 100 dB: the tensor-core kernel's output layer sums in a position-dependent order; bit
for bit with the FP32 kernel).  python tools/host_stress.py [seconds] [tc|ffma]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fullycnnspeechenhancement_b200.engine import Enhancer, PinnedArray, host_tables      # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance                          # noqa: E402
from oracle import network, rebuild                                                       # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
variant = sys.argv[2] if len(sys.argv) > 2 else "tc"
rng = np.random.default_rng(123)
w = network.random_weights("FullyCNNV2", seed=5, randomize_bn=True)
eng = Enhancer("FullyCNNV2", w, device=0, variant=variant)
pool = [noisy_utterance(3000 + i, int(n)) for i, n in enumerate(rng.integers(1, 20000, 48))] + [noisy_utterance(9, 1), noisy_utterance(10, 255)]
alone = [eng.enhance([x])[0] for x in pool]
t_end = time.time() + budget
calls = utts = 0
worst = 1e9
while time.time() < t_end:
    eng.host_config(chunk_rows=int(rng.integers(20, 4000)), chunk_rows_async=int(rng.integers(50, 20000)))
    inflight = []
    for _ in range(int(rng.integers(1, 5))):
        idx = rng.integers(0, len(pool), int(rng.integers(1, 120)))
        t = host_tables(np.array([len(pool[i]) for i in idx]))
        hin, hout = PinnedArray(t["total"]), PinnedArray(t["total"])
        hout.array[:] = -9.0
        for i, o in zip(idx, t["wav_off"]):
            hin.array[o:o + len(pool[i])] = pool[i]
        sync = rng.random() < 0.3
        eng.enhance_host(hin.array, hout.array, t, sync=bool(sync))
        inflight.append((idx, t, hin, hout))
        calls += 1
    eng.host_sync()
    for idx, t, hin, hout in inflight:
        for i, o in zip(idx, t["out_off"]):
            got, ref = hout.array[o:o + len(pool[i])], alone[i]
            if variant == "ffma":
                assert np.array_equal(got, ref), (i, len(ref))
            else:
                snr = rebuild.sdr_db(ref.astype(np.float64), got.astype(np.float64)) if np.any(ref) else 200.0
                worst = min(worst, snr)
                assert snr >= 100.0 or np.abs(got - ref).max() < 1e-6, (i, len(ref), snr)
            utts += 1
        hin.close()
        hout.close()
print("host stress ok: %s, %d calls, %d utterances checked, worst SNR vs alone %.1f dB, guard %s" % (variant, calls, utts, worst, eng.tc_status() if variant == "tc" else ""))
