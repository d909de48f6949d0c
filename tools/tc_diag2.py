import sys, numpy as np, torch
sys.path.insert(0, '.')
from fullycnnspeechenhancement_b200.engine import Enhancer
from fullycnnspeechenhancement_b200.synth import noisy_utterance
from oracle import network, rebuild
name="FullyCNNV2"
w = network.random_weights(name, seed=13, randomize_bn=True)
eng = Enhancer(name, w, device=0)
minute = noisy_utterance(9000, 8000 * 60)
gain = np.linspace(0.5, 1.0, 5, dtype=np.float32)
wv = np.concatenate([minute * g for g in gain])
res = {}
for v in ("ffma", "tc"):
    eng.set_variant(v)
    res[v, "whole"] = eng.enhance([wv])[0]
    res[v, "chunk"] = eng.enhance_stream(wv, chunk_seconds=4.0)
    if v == "tc": print("tc status", eng.tc_status())
def sdr(a, b): return rebuild.sdr_db(a, b)
print("ffma whole vs chunk", sdr(res["ffma","whole"], res["ffma","chunk"]))
print("tc   whole vs chunk", sdr(res["tc","whole"], res["tc","chunk"]))
print("whole ffma vs tc   ", sdr(res["ffma","whole"], res["tc","whole"]))
print("chunk ffma vs tc   ", sdr(res["ffma","chunk"], res["tc","chunk"]))
d = np.abs(res["tc","whole"].astype(np.float64) - res["tc","chunk"])
idx = np.argsort(-d)[:15]
print("largest |diff| at samples", sorted(idx.tolist()), "values", d[idx][:5], "signal there", np.abs(res["tc","whole"][idx][:5]))
seg = d.reshape(-1, 128).max(axis=1)
bad = np.nonzero(seg > 1e-3 * np.abs(res["tc","whole"]).max())[0]
print("segments with large diff:", len(bad), bad[:40])
d2 = np.abs(res["ffma","whole"].astype(np.float64) - res["tc","whole"])
seg2 = d2.reshape(-1, 128).max(axis=1)
bad2 = np.nonzero(seg2 > 1e-3 * np.abs(res["tc","whole"]).max())[0]
print("ffma-vs-tc whole: segments with large diff:", len(bad2), bad2[:40])
