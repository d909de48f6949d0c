"""CPU: the oracle against the committed golden vectors (which were produced by the unmodified
reference code, tests/golden/make_golden.py) and against itself."""
import os

import numpy as np
import pytest

from oracle import network, rebuild, stft


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "stft_rebuild_ref.npz"))


def test_frame_counts_and_indices_bit_exact(golden_dir):
    fc = np.load(os.path.join(golden_dir, "frame_counts_ref.npz"))
    for L, T, first, last in zip(fc["lengths"], fc["counts"], fc["first_start"], fc["last_start"]):
        assert stft.frame_count(int(L)) == T
        idx = stft.frame_indices(int(L))
        assert idx.shape == (T, 256)
        assert idx[0, 0] == first == 0
        if last >= 0:                      # last frame starts inside the signal
            assert idx[-1, 0] == last
        assert idx[-1, 0] == (T - 1) * 128
    # values called out in SURVEY.md section 8
    for L, T in [(32000, 249), (16000, 124), (24001, 187), (64000, 499), (28800000, 224999), (256, 1), (100, 3)]:
        assert stft.frame_count(L) == T


def test_stft_matches_reference_bit_exact(g):
    for seed, L in zip(g["case_seeds"], g["case_lengths"]):
        x = g["wav_%d" % seed]
        assert x.dtype == np.float32 and len(x) == L
        X = stft.compute_spectrogram(x, 8000, 0.032, 0.016, 256, True)
        assert np.array_equal(X, g["spec_%d" % seed])
        assert np.array_equal(stft.power_spectrum(X), g["mag_%d" % seed])
        assert np.array_equal(stft.divide_phase(X), g["phase_%d" % seed])


def test_padding_batch_matches_reference(g):
    seeds = g["case_seeds"][3:]
    pb = stft.padding_batch([g["spec_%d" % s] for s in seeds])
    assert np.array_equal(pb, g["padded_batch"])
    # padded frames: magnitude 0, phase exactly 1+0j
    T_short = g["spec_%d" % seeds[0]].shape[1]
    assert np.all(stft.power_spectrum(pb)[0, T_short:] == 0)
    assert np.all(stft.divide_phase(pb)[0, T_short:] == 1 + 0j)


@pytest.mark.parametrize("nfft", [512, 256])
def test_rebuild_matches_reference_bit_exact(g, nfft):
    pb = g["padded_batch"]
    phase = stft.divide_phase(pb).squeeze(-1)
    lens = [int(x) for x in g["case_lengths"][3:]]
    for loop in (False, True):
        out = rebuild.rebuild_audio(lens, g["pred"], phase, 8000, 32.0, 16.0, nfft=nfft, faithful_loop=loop)
        for i, L in enumerate(lens):
            assert len(out[i]) == L
            assert np.array_equal(out[i], g["rebuild%d_%d" % (nfft, i)])


def test_de_emphasis_filter_equals_reference_loop():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 5000))
    assert np.array_equal(rebuild.de_emphasis(x), rebuild.de_emphasis_loop(x))


def test_pre_emphasis_is_float32_without_fma():
    x = np.random.default_rng(1).normal(size=1000).astype(np.float32)
    e = stft.pre_emphasis(x)
    assert e.dtype == np.float32
    manual = np.float32(x[1:]) - np.float32(np.float32(0.97) * x[:-1])
    assert np.array_equal(e[1:], manual) and e[0] == x[0]


def test_identity_network_known_answer():
    """irfft 256 inverts the analysis (> 100 dB), the shipped
    default irfft 512 does not (about -15 dB): both are reference behaviour."""
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    x = noisy_utterance(3, 32000)
    X = stft.compute_spectrogram(x, 8000, 0.032, 0.016, 256, True).T[None]
    mag, ph = stft.power_spectrum(X), stft.divide_phase(X)
    y256 = rebuild.rebuild_audio([32000], mag, ph, nfft=256)[0]
    y512 = rebuild.rebuild_audio([32000], mag, ph, nfft=512)[0]
    assert rebuild.sdr_db(x.astype(np.float64), y256) > 100
    assert -20 < rebuild.sdr_db(x.astype(np.float64), y512) < -10


def test_parameter_counts_match_readme():
    # readme.md:63-67 of the reference
    assert network.trainable_param_count("FullyCNN") == 32765
    assert network.trainable_param_count("FullyCNNV2") == 32192
    assert network.trainable_param_count("FullyCNNV3") == 32653
    # SURVEY.md section 8a MAC counts
    assert network.mac_per_frame("FullyCNNV2", False) == 4054728
    assert network.mac_per_frame("FullyCNNV2", True) == 3959092


@pytest.mark.parametrize("arch", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_network_two_evaluators_and_golden(arch, golden_dir):
    n = np.load(os.path.join(golden_dir, "network_oracle.npz"))
    w = network.random_weights(arch, seed=1234, randomize_bn=True)
    s = [float(np.sum([np.sum(v.astype(np.float64)) for v in w.values()])),
         float(np.sum([np.sum(np.abs(v.astype(np.float64))) for v in w.values()]))]
    assert np.allclose(s, n["wsum_" + arch], rtol=0, atol=1e-9)      # seeded weights are reproducible
    for T in (1, 7, 8, 9, 12):
        x = n["x_%s_%d" % (arch, T)]
        y = network.forward(arch, w, x, np.float64)
        assert np.allclose(y, n["y_%s_%d" % (arch, T)], rtol=0, atol=1e-12)
        y2 = network.forward_torch(arch, w, x, "float64")
        assert np.abs(y - y2).max() <= 1e-12 * max(1.0, np.abs(y).max())
        y32 = network.forward_torch(arch, w, x, "float32")
        assert np.abs(y - y32).max() / np.abs(y).max() < 1e-5


@pytest.mark.parametrize("arch", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_network_oracle_matches_the_reference_model_code(arch, golden_dir):
    """network_ref_model.npz holds what the reference's OWN model classes (model_utils/model.py, unmodified) compute with
    oracle/tf_standin.py in place of TensorFlow (tests/golden/make_golden.py): the wiring is the reference's, executed.
    The oracle must reproduce it, and the variables the reference's code creates must be exactly the ones the weight
    dicts (oracle.network.random_weights, model_utils/fold.py) carry, with the same shapes."""
    g = np.load(os.path.join(golden_dir, "network_ref_model.npz"))
    n = np.load(os.path.join(golden_dir, "network_oracle.npz"))
    w = network.random_weights(arch, seed=1234, randomize_bn=True)
    for T in (1, 8, 12):
        x = n["x_%s_%d" % (arch, T)]
        assert abs(float(x.astype(np.float64).sum()) - float(g["xsum_%s_%d" % (arch, T)][0])) < 1e-9
        ref = g["y_%s_%d" % (arch, T)]
        y = network.forward(arch, w, x, np.float64)
        assert y.shape == ref.shape
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    names, shapes = [str(v) for v in g["names_" + arch]], [str(v) for v in g["shapes_" + arch]]
    assert sorted(names) == sorted(w.keys())
    for nm, sh in zip(names, shapes):
        assert "x".join(str(d) for d in w[nm].shape) == sh, nm
    # creation order of the conv scopes = the oracle's layer table (the order checkpoints and frozen graphs list them in)
    scopes = [nm[:-len("/kernel")] for nm in names if nm.endswith("/kernel")]
    assert scopes == [L["scope"] for L in network.layer_table(arch)]
    from fullycnnspeechenhancement_b200.model_utils import fold
    assert sorted(fold.glorot_weights(arch, seed=0).keys()) == sorted(names)


@pytest.mark.parametrize("arch", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_reference_model_code_live(arch):
    """Where /root/reference exists (the authoring container): run the reference's model classes through the stand-in on
    fresh weights and inputs, ragged T included, against the oracle."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    from oracle import tf_standin
    cls = ref_import.load_models()[arch]
    rng = np.random.default_rng(31)
    for seed, T, bn in ((3, 1, True), (4, 5, True), (5, 16, False)):
        w = network.random_weights(arch, seed=seed, randomize_bn=bn)
        x = np.abs(rng.normal(0, 2.0, (3, T, 129, 1)))
        tf_standin.set_variables(w)
        ref = cls(is_training=False)(x)
        y = network.forward(arch, w, x, np.float64)
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        assert len(tf_standin.requested()) == len(w)


def test_network_padding_invariance():
    arch = "FullyCNNV3"
    w = network.random_weights(arch, seed=2)
    rng = np.random.default_rng(0)
    a = np.abs(rng.normal(size=(1, 6, 129, 1)))
    batch = np.zeros((2, 11, 129, 1))
    batch[0, :6] = a[0]
    batch[1] = np.abs(rng.normal(size=(11, 129, 1)))
    alone = network.forward(arch, w, a)
    both = network.forward(arch, w, batch)
    # frames whose 8-tap time window stays inside the 6 valid frames (+ zero padding == SAME padding)
    assert np.allclose(alone[0, :2], both[0, :2], atol=1e-12)


@pytest.mark.parametrize("arch", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_network_third_evaluator_scipy(arch):
    """A third, independently written evaluation of the graph: per (cin, cout) pair
    ``scipy.signal.correlate2d`` in 'full' mode, cropped to TensorFlow's SAME window (output pixel i reads
    inputs i - (k-1)//2 ... i + k-1 - (k-1)//2), layer wiring re-read from model_utils/model.py rather
    than taken from the oracle's executor.  TensorFlow itself is absent (parity unpinned);
    this pins the oracle's conv / padding / BN / skip arithmetic against another formulation."""
    from scipy.signal import correlate2d
    w = network.random_weights(arch, seed=21, randomize_bn=True)
    rng = np.random.default_rng(4)
    T = 6
    x = np.abs(rng.normal(0, 2, (T, 129))).astype(np.float64)

    def conv_same(inp, kernel, bias):                       # inp [cin][T][F], kernel HWIO
        kh, kw, cin, cout = kernel.shape
        out = np.zeros((cout, T, 129))
        for o in range(cout):
            for c in range(cin):
                full = correlate2d(inp[c], kernel[:, :, c, o].astype(np.float64), mode="full")
                # full[i + kh - 1 - pt] is the response centred so that SAME's "pt before" holds
                r0, c0 = kh - 1 - (kh - 1) // 2, kw - 1 - (kw - 1) // 2
                out[o] += full[r0:r0 + T, c0:c0 + 129]
            out[o] += np.float64(bias[o])
        return out

    def bn(y, s):
        g_, b_, m_, v_ = (w[s + "/batch_norm/" + n].astype(np.float64) for n in ("gamma", "beta", "moving_mean", "moving_variance"))
        return (y - m_[:, None, None]) / np.sqrt(v_[:, None, None] + 1e-3) * g_[:, None, None] + b_[:, None, None]

    def cbr(inp, s, norm=True, act=True, skip=None):
        y = conv_same(inp, w[s + "/kernel"], w[s + "/bias"])
        if norm:
            y = bn(y, s)
        if skip is not None:
            y = y + skip
        return np.maximum(y, 0) if act else y

    a = x[None]
    if arch == "FullyCNNV3":                                 # model.py:64-96
        def block(inp, name, skip=None):
            e = cbr(cbr(cbr(inp, name + "_encode_1"), name + "_encode_2"), name + "_decode")
            return e if skip is None else e + skip
        c1 = block(a, "CE1")
        c2 = block(c1, "CE2")
        c3 = block(c2, "CE3")
        d = block(block(c3, "CD1", c2), "CD2", c1)
        out = cbr(d, "decode_final", norm=False, act=False)
    elif arch == "FullyCNNV2":                               # model.py:32-61
        e = [a]
        for i in range(1, 9):
            e.append(cbr(e[-1], "encode_%d" % i))
        d = e[8]
        for i in range(1, 8):
            d = cbr(d, "decode_%d" % i, skip=e[8 - i])
        out = cbr(d, "decode_8", norm=False, act=False)
    else:                                                    # model.py:6-29 (the fifth encoder scope is "encode_8")
        e1 = cbr(a, "encode_1")
        e2 = cbr(e1, "encode_2")
        e3 = cbr(e2, "encode_3")
        e4 = cbr(e3, "encode_4")
        e5 = cbr(e4, "encode_8")
        d = cbr(e5, "decode_1", skip=e4)
        d = cbr(d, "decode_2", skip=e3)
        d = cbr(d, "decode_3", skip=e2)
        d = cbr(d, "decode_4", skip=e1)
        out = cbr(d, "decode_5", norm=False, act=False)
    ref = network.forward(arch, w, x[None, :, :, None], np.float64)[0, :, :, 0]
    assert np.abs(out[0] - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.parametrize("arch", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_network_fourth_evaluator_library_same_padding_and_batch_norm(arch):
    """A fourth evaluation that does not hand-write what the three others share: the SAME padding is
    ``torch.nn.functional.conv2d(padding='same')`` -- the library's own rule for even kernels (kh = 8: the extra row goes
    behind, like TensorFlow's (k-1)//2 before) -- and the inference batch norm is ``F.batch_norm(training=False,
    eps=1e-3)`` with the stored moving statistics, i.e. library code instead of the oracle's formula.  The graph is
    written as the reference writes it (module.py:11-34: conv -> BN -> + skip -> ReLU; model.py for the wiring).
    TensorFlow itself stays absent (parity unpinned): this removes the builder's reading of padding and BN
    from the list of things all evaluators could share."""
    import torch
    import torch.nn.functional as F
    w = network.random_weights(arch, seed=33, randomize_bn=True)
    rng = np.random.default_rng(8)
    T = 11
    x = np.abs(rng.normal(0, 2, (2, T, 129, 1)))
    tw = {k: torch.from_numpy(v).double() for k, v in w.items()}

    def conv_bn_relu(t, scope, use_norm=True, use_act=True, skip_input=None):
        k = tw[scope + "/kernel"].permute(3, 2, 0, 1).contiguous()          # HWIO -> OIHW
        t = F.conv2d(t, k, tw[scope + "/bias"], stride=1, padding="same")
        if use_norm:
            t = F.batch_norm(t, tw[scope + "/batch_norm/moving_mean"], tw[scope + "/batch_norm/moving_variance"],
                             tw[scope + "/batch_norm/gamma"], tw[scope + "/batch_norm/beta"], training=False, eps=1e-3)
        if skip_input is not None:
            t = t + skip_input
        return torch.relu(t) if use_act else t

    t = torch.from_numpy(x).double().permute(0, 3, 1, 2)                   # NHWC -> NCHW (H = time, W = frequency)
    with torch.no_grad():
        if arch == "FullyCNNV2":
            enc = [t]
            for i in range(1, 9):
                enc.append(conv_bn_relu(enc[-1], "encode_%d" % i))
            d = enc[8]
            for i in range(1, 8):
                d = conv_bn_relu(d, "decode_%d" % i, skip_input=enc[8 - i])
            out = conv_bn_relu(d, "decode_8", use_norm=False, use_act=False)
        elif arch == "FullyCNNV3":
            def simple_rced(inp, name, skip=None):
                e = conv_bn_relu(conv_bn_relu(conv_bn_relu(inp, name + "_encode_1"), name + "_encode_2"), name + "_decode")
                return e if skip is None else e + skip
            c1 = simple_rced(t, "CE1")
            c2 = simple_rced(c1, "CE2")
            c3 = simple_rced(c2, "CE3")
            out = conv_bn_relu(simple_rced(simple_rced(c3, "CD1", c2), "CD2", c1), "decode_final", use_norm=False, use_act=False)
        else:
            e1 = conv_bn_relu(t, "encode_1")
            e2 = conv_bn_relu(e1, "encode_2")
            e3 = conv_bn_relu(e2, "encode_3")
            e4 = conv_bn_relu(e3, "encode_4")
            e5 = conv_bn_relu(e4, "encode_8")
            d = conv_bn_relu(e5, "decode_1", skip_input=e4)
            d = conv_bn_relu(d, "decode_2", skip_input=e3)
            d = conv_bn_relu(d, "decode_3", skip_input=e2)
            d = conv_bn_relu(d, "decode_4", skip_input=e1)
            out = conv_bn_relu(d, "decode_5", use_norm=False, use_act=False)
    got = out.permute(0, 2, 3, 1).numpy()
    ref = network.forward(arch, w, x, np.float64)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
