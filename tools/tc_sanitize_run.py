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
    out = eng.enhance(waves)                      # small launch: 1 frame per CTA batch
    torch.cuda.synchronize()
    print(name, [len(o) for o in out], eng.tc_status())
    many = [rng.normal(0, 0.1, 4000).astype(np.float32) for _ in range(80)]      # 2,640 rows: full 7-frame batches, several per CTA
    eng.host_config(chunk_rows=700)               # several chunks: copy-in / compute / copy-out streams, buffer-set ring
    out = eng.enhance(many)
    print(name, len(out), eng.tc_status())
    eng.set_variant("ffma")
    eng.set_skip_in_tmem(False)                   # FFMA kernel with its skips in the claimed global scratch regions
    out = eng.enhance(waves)
    print(name, "ffma/global skips", [len(o) for o in out])
    eng.close()
