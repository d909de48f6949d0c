"""Small K1 -> K2-TC -> K3 run for compute-sanitizer (memcheck / synccheck):
   compute-sanitizer --tool memcheck python tools/tc_sanitize_run.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from fullycnnspeechenhancement_b200.engine import Enhancer            # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import fold           # noqa: E402

for name in ("FullyCNN", "FullyCNNV2", "FullyCNNV3"):
    eng = Enhancer(name, fold.glorot_weights(name, seed=0), device=0)
    eng.set_variant("tc")
    rng = np.random.default_rng(1)
    waves = [rng.normal(0, 0.1, n).astype(np.float32) for n in (4000, 900, 12345, 256, 7000)]
    out = eng.enhance(waves)
    torch.cuda.synchronize()
    print(name, [len(o) for o in out], eng.tc_status())
