"""CPU: layout and arithmetic of the tensor-core network kernel, without a GPU.

tools/tc_emulator.py executes the kernel's data flow (FP16 hi/lo weight image and descriptor table
from librced_b200.so, flattened row space, two instructions per K step, "taps in N" output layer)
in numpy.  Comparing it with the float64 oracle pins (a) the weight image / descriptor arithmetic
the CUDA kernel shares through csrc/rced_tc.cuh and (b) the accuracy of the FP16 x3 split, which
the north star requires to be stated separately from the FP32 path: max |err| / max |ref| <= 1e-4."""
import numpy as np
import pytest

from fullycnnspeechenhancement_b200 import _lib
from fullycnnspeechenhancement_b200.model_utils import fold
from oracle import network

TC_TOL = 1e-4   # same bar as the FP32 kernel (BASELINE.json north_star); measured ~1e-6


def oracle_ragged(name, w, mag, row_off):
    out = np.zeros_like(mag, dtype=np.float64)
    for u in range(len(row_off) - 1):
        a, b = row_off[u], row_off[u + 1]
        out[a:b] = network.forward(name, w, mag[a:b][None, :, :, None], np.float64)[0, :, :, 0]
    return out


@pytest.mark.parametrize("name", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_tc_emulator_matches_oracle(name):
    import tc_emulator
    lib = _lib.lib()
    w = network.random_weights(name, 11, True)
    folded = fold.fold_batch_norm(w, name)
    rng = np.random.default_rng(5)
    # ragged batch: utterance boundaries inside a 7-frame batch, a 1-frame utterance, a partial last batch
    lens = [5, 1, 9, 3]
    row_off = np.concatenate([[0], np.cumsum(lens)])
    mag = np.abs(rng.normal(0, 3, (row_off[-1], 129))).astype(np.float32)
    ref = oracle_ragged(name, w, mag, row_off)
    got, amax = tc_emulator.run(lib, fold.arch_id(name), folded, mag, row_off, network.layer_table(name))
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < TC_TOL, err
    assert err < 2e-5, "FP16 x3 split should be within a few 1e-6 of float64: %g" % err
    assert 0 < amax < 65504


def _per_utterance_err(got, ref, row_off):
    return max(float(np.abs(got[a:b] - ref[a:b]).max() / np.abs(ref[a:b]).max()) for a, b in zip(row_off[:-1], row_off[1:]))


@pytest.mark.parametrize("scale", [1.0, 1e-2, 1e-4, 1e-6, 1e3])
@pytest.mark.parametrize("randomize_bn", [False, True])
def test_tc_emulator_is_scale_invariant(scale, randomize_bn):
    """Round-1 finding: the unscaled FP16 residual fell into the subnormals for small inputs (3.9e-4 at max |mag| = 1e-3
    with the bench's bias-free weights).  With the residual stored times 2^11, per-step weight scales and per-frame
    activation scales the error is the same at every input scale -- normalised PER UTTERANCE."""
    import tc_emulator
    lib = _lib.lib()
    name = "FullyCNNV2"
    w = network.random_weights(name, 0 if not randomize_bn else 11, randomize_bn)   # seed 0, identity BN: bench.py's weights
    folded = fold.fold_batch_norm(w, name)
    rng = np.random.default_rng(5)
    row_off = np.concatenate([[0], np.cumsum([5, 1, 9, 3])])
    mag = (np.abs(rng.normal(0, 3, (row_off[-1], 129))) * scale).astype(np.float32)
    ref = oracle_ragged(name, w, mag, row_off)
    got, amax = tc_emulator.run(lib, fold.arch_id(name), folded, mag, row_off, network.layer_table(name))
    assert _per_utterance_err(got, ref, row_off) < 2e-6
    assert 1.0 < amax < 65504     # the scaled domain is the same whatever the input scale


@pytest.mark.parametrize("name", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_tc_emulator_loud_and_quiet_utterances_share_a_batch(name):
    """A -80 dB utterance next to a loud one inside the same 7-frame CTA batch: every frame has its own scale."""
    import tc_emulator
    lib = _lib.lib()
    w = network.random_weights(name, 11, True)
    folded = fold.fold_batch_norm(w, name)
    rng = np.random.default_rng(6)
    row_off = np.concatenate([[0], np.cumsum([5, 1, 9, 3])])
    mag = np.abs(rng.normal(0, 3, (row_off[-1], 129))).astype(np.float32)
    mag[5:6] *= 1e-4
    mag[15:] *= 1e-4
    ref = oracle_ragged(name, w, mag, row_off)
    got, _ = tc_emulator.run(lib, fold.arch_id(name), folded, mag, row_off, network.layer_table(name))
    assert _per_utterance_err(got, ref, row_off) < 2e-6


def test_tc_emulator_weights_of_any_magnitude():
    """Per-step power-of-two weight scales: layers whose folded weights are tiny or huge (BN with extreme gamma / variance)
    lose nothing; the old packing refused weights beyond 65504 and lost the residual of weights below 1e-3."""
    import tc_emulator
    lib = _lib.lib()
    name = "FullyCNNV2"
    w = network.random_weights(name, 3, True)
    table = network.layer_table(name)
    # layer 2 shrinks by 2e-5, layer 3 grows by 5e4 (their product keeps the activations where they were)
    w[table[2]["scope"] + "/kernel"] = w[table[2]["scope"] + "/kernel"] * np.float32(2e-5)
    w[table[2]["scope"] + "/bias"] = w[table[2]["scope"] + "/bias"] * np.float32(2e-5)
    w[table[2]["scope"] + "/batch_norm/beta"] = w[table[2]["scope"] + "/batch_norm/beta"] * np.float32(2e-5)
    w[table[2]["scope"] + "/batch_norm/moving_mean"] = w[table[2]["scope"] + "/batch_norm/moving_mean"] * np.float32(2e-5)
    w[table[3]["scope"] + "/kernel"] = w[table[3]["scope"] + "/kernel"] * np.float32(5e4)
    folded = fold.fold_batch_norm(w, name)
    assert np.abs(folded).max() > 65504 / 16      # far outside what an unscaled FP16 image could hold precisely
    rng = np.random.default_rng(7)
    row_off = np.array([0, 9])
    mag = np.abs(rng.normal(0, 3, (9, 129))).astype(np.float32)
    ref = oracle_ragged(name, w, mag, row_off)
    got, _ = tc_emulator.run(lib, fold.arch_id(name), folded, mag, row_off, network.layer_table(name))
    assert _per_utterance_err(got, ref, row_off) < 1e-5


def test_tc_layout_fits_the_sm():
    import tc_emulator
    lib = _lib.lib()
    for arch in (1, 2, 3):
        lay = tc_emulator.layout(lib, arch)
        assert lay["smem"] <= 227 * 1024
        assert lay["tiles"] * 64 <= 512                      # tensor-memory columns
        assert lay["fb"] * lay["fs"] <= lay["tiles"] * 128   # frames of a batch fit the row tiles
        # every A descriptor stays inside the allocated planes, LBO is positive
        for st in lay["steps"]:
            for u in range(st["units"]):
                off, lbo = lay["units"][st["unit_base"] + u]
                assert lbo > 0
                assert lay["front_rows"] + lay["lead"] + off >= 0            # the -64-row shift reads the zero front rows
                if not st["final"]:
                    assert lay["lead"] + 128 * (lay["tiles"] - 1) + off + lbo + 128 <= 4 * lay["plane16"]
                else:   # even-frame copy in planes 0-1; the +64 shift of the last row tile ends in the next plane's lead rows
                    assert lbo == lay["plane16"] and abs(off) <= 64
                    # the last E row any output needs is bin 128 of the last frame + 31 columns: its shifted read
                    # must end inside the zero lead rows of the next plane (rows behind it feed no output)
                    last_needed = (lay["fb"] - 1) * lay["fs"] + 128 + lay["final_n"] - 1
                    assert lay["lead"] + last_needed + off < lay["plane16"] + lay["lead"]
            assert st["tile_bytes"] % 16 == 0 and st["w_off"] % 16 == 0
