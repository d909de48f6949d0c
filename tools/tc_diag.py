import sys, numpy as np, torch
sys.path.insert(0, '.')
from fullycnnspeechenhancement_b200.engine import Enhancer
from oracle import network
name="FullyCNNV2"
w = network.random_weights(name, seed=4321, randomize_bn=True)
eng = Enhancer(name, w, device=0)
rng = np.random.default_rng(3)
for lens in ([40], [5], [1], [8, 13], [1100], [148*7*2+3]):
    ro = np.concatenate([[0], np.cumsum(lens)])
    mag = np.abs(rng.normal(0, 3, (ro[-1], 129))).astype(np.float32)
    d_mag = torch.from_numpy(mag).cuda(); d_ro = torch.from_numpy(ro.astype(np.int64)).cuda()
    eng.set_variant("tc"); tc = eng.forward_device(d_mag, d_ro).cpu().numpy(); st = eng.tc_status()
    eng.set_variant("ffma"); ff = eng.forward_device(d_mag, d_ro).cpu().numpy()
    err = np.abs(tc - ff).max(axis=1) / np.abs(ff).max()
    bad = np.nonzero(err > 1e-4)[0]
    print(lens, "status", st, "max row err %.3g" % err.max(), "bad rows", bad[:20], "of", len(bad))
