import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from fullycnnspeechenhancement_b200.engine import Enhancer
from fullycnnspeechenhancement_b200.model_utils import fold
eng = Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2", seed=0), device=0)
eng.set_variant("tc")
rows = 148 * 7 * 3
mag = torch.rand((rows, 129), device="cuda")
ro = torch.tensor([0, rows], dtype=torch.int64, device="cuda")
eng.forward_device(mag, ro); torch.cuda.synchronize()
os.environ["RCED_TC_TRACE"] = "gpurun_out/tc_trace.txt"
eng.forward_device(mag, ro); torch.cuda.synchronize()
print(eng.tc_status())
