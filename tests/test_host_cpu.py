"""CPU: host-side logic and the C ABI surface (no compute call needs a GPU here)."""
import configparser
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from fullycnnspeechenhancement_b200 import _lib                       # noqa: E402
from fullycnnspeechenhancement_b200.model_utils import ckpt, fold       # noqa: E402
from oracle import network, stft                                        # noqa: E402


def header_functions():
    text = open(os.path.join(ROOT, "include", "rced.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rced_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "librced_b200.so does not export " + n
    # and the python binding covers the whole header
    assert sorted(_lib.SIGNATURES) == names


def test_sass_is_blackwell_native():
    """The shipped cubin is sm_100a and uses tensor memory and bulk async copies."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    # tensor-memory store / load, bulk async copy, FP32 FMA, tcgen05.mma and its commit, uniform constant loads
    for mnemonic in ("STTM", "LDTM", "UBLKCP", "FFMA", "UTCHMMA", "UTCBAR", "LDCU"):
        assert mnemonic in out, mnemonic


def test_num_frames_bit_exact_vs_reference(golden_dir):
    fc = np.load(os.path.join(golden_dir, "frame_counts_ref.npz"))
    lib = _lib.lib()
    for L, T in zip(fc["lengths"], fc["counts"]):
        assert lib.rced_num_frames(int(L)) == T
    from fullycnnspeechenhancement_b200.engine import num_frames
    assert np.array_equal(num_frames(fc["lengths"]), fc["counts"])
    assert lib.rced_num_frames(28800000) == 224999


def test_layer_tables_agree_between_python_cuda_and_oracle():
    lib = _lib.lib()
    for name, arch in fold.ARCH_IDS.items():
        table = network.layer_table(name)
        scopes = fold.layer_scopes(name)
        assert lib.rced_num_layers(arch) == len(table) == len(scopes)
        for i, (L, (scope, bn)) in enumerate(zip(table, scopes)):
            v = [ctypes.c_int() for _ in range(4)]
            _lib.check(lib.rced_layer_shape(arch, i, *v))
            assert tuple(x.value for x in v) == (L["kh"], L["kw"], L["cin"], L["cout"])
            assert scope == L["scope"] and bn == L["norm"]
        assert lib.rced_mac_per_frame(arch, 0) == network.mac_per_frame(name, False)
        assert lib.rced_mac_per_frame(arch, 1) == network.mac_per_frame(name, True)
        w = network.random_weights(name, 0)
        assert fold.trainable_parameter_count(w, name) == network.trainable_param_count(name)
        assert fold.fold_batch_norm(w, name).size == lib.rced_folded_weight_count(arch)
    assert fold.arch_id("anything else") == 1          # tester.py:80-82 default


@pytest.mark.parametrize("name", ["FullyCNN", "FullyCNNV2", "FullyCNNV3"])
def test_packed_weights_and_kernel_addressing(name):
    """The numpy emulator executes the network kernel's shared-memory addressing on the image
    produced by rced_pack_weights and must reproduce the oracle (BN folding included)."""
    import net_emulator
    lib = _lib.lib()
    w = network.random_weights(name, 3, True)
    table = network.layer_table(name)
    folded = fold.fold_batch_norm(w, name)
    x = np.abs(np.random.default_rng(1).normal(0, 2, (1, 5, 129, 1))).astype(np.float32)
    ref = network.forward(name, w, x, np.float64)[0, :, :, 0]
    got = net_emulator.run(lib, fold.arch_id(name), folded, x[0, :, :, 0].astype(np.float64),
                           [t["act"] for t in table], [t["skip_after_act"] for t in table])
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-6


def test_fold_batch_norm_is_exact_algebra():
    name = "FullyCNNV2"
    w = network.random_weights(name, 5, True)
    folded = fold.fold_batch_norm(w, name)
    # rebuild a BN-free weight dict from the folded vector and run the oracle on it
    w2, pos = {}, 0
    for L in network.layer_table(name):
        n = L["kh"] * L["kw"] * L["cin"] * L["cout"]
        w2[L["scope"] + "/kernel"] = folded[pos:pos + n].reshape(L["kh"], L["kw"], L["cin"], L["cout"]); pos += n
        w2[L["scope"] + "/bias"] = folded[pos:pos + L["cout"]]; pos += L["cout"]
        if L["norm"]:
            c = L["cout"]
            w2[L["scope"] + "/batch_norm/gamma"] = np.full(c, np.sqrt(1 + 1e-3))
            w2[L["scope"] + "/batch_norm/beta"] = np.zeros(c)
            w2[L["scope"] + "/batch_norm/moving_mean"] = np.zeros(c)
            w2[L["scope"] + "/batch_norm/moving_variance"] = np.ones(c)
    assert pos == folded.size
    x = np.abs(np.random.default_rng(2).normal(size=(1, 4, 129, 1)))
    a = network.forward(name, w, x)
    b = network.forward(name, {k: np.asarray(v, np.float64) for k, v in w2.items()}, x)
    assert np.abs(a - b).max() / np.abs(a).max() < 1e-6


def test_checkpoint_and_frozen_graph_round_trip(tmp_path):
    assert ckpt.crc32c(b"123456789") == 0xE3069283
    for name in ("FullyCNN", "FullyCNNV2", "FullyCNNV3"):
        w = network.random_weights(name, 7, True)
        extra = dict(w)
        extra["global_step"] = np.array(123, np.int64)           # trainer.py:28 extras are ignored
        extra["beta1_power"] = np.array(0.9, np.float32)
        for k in list(w):
            if k.endswith("kernel"):
                extra[k + "/Adam"] = np.zeros_like(w[k])
        prefix = str(tmp_path / name / ("RCED_%s_0_9.ckpt" % name))
        ckpt.write_checkpoint(prefix, extra)
        assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
        r = ckpt.load_weights(prefix, name)
        assert set(r) == set(w) and all(np.array_equal(r[k], w[k]) for k in w)
        everything = ckpt.read_checkpoint(prefix)
        assert everything["global_step"] == 123
        pb = str(tmp_path / (name + ".pb"))
        ckpt.write_frozen_graph(pb, name, w)
        r2 = ckpt.load_weights(pb, name)
        assert all(np.array_equal(r2[k], w[k]) for k in w)
    with pytest.raises(KeyError):
        ckpt.read_checkpoint(prefix, ["not/a/variable"])
    with pytest.raises(FileNotFoundError):
        ckpt.read_checkpoint(str(tmp_path / "missing.ckpt"))
    # corrupt one byte of the data file: checksum must catch it
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[100] ^= 0xFF
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        ckpt.read_checkpoint(prefix)


def test_frozen_graph_parses_with_protobuf_library(tmp_path):
    graph_pb2 = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    w = network.random_weights("FullyCNNV2", 7, True)
    pb = str(tmp_path / "v2.pb")
    ckpt.write_frozen_graph(pb, "FullyCNNV2", w)
    g = graph_pb2.GraphDef()
    g.ParseFromString(open(pb, "rb").read())
    consts = {n.name: n for n in g.node if n.op == "Const"}
    assert set(consts) == set(w)
    arr = np.frombuffer(consts["encode_1/kernel"].attr["value"].tensor.tensor_content, np.float32)
    assert np.array_equal(arr, w["encode_1/kernel"].ravel())
    assert any(n.name == fold.output_node_name("FullyCNNV2") for n in g.node)
    assert any(n.name == "input" and n.op == "Placeholder" for n in g.node)


def test_config_loader_contract(tmp_path):
    from fullycnnspeechenhancement_b200.config import load_conf_info, section_with
    assert load_conf_info(str(tmp_path / "missing.cfg")).sections() == []      # silently empty, like the reference
    p = tmp_path / "infer.cfg"
    p.write_text("[inference]\ncheckpoint_filepath=a/b.ckpt\n[model]\nnet_arch=RCED\nnet_work=FullyCNNV2\n"
                 "[data]\nsample_rate=8000\nfeature_dim=129\nwindow_ms=32\nstride_ms=16\naudio_save_path=x/\n")
    cfg = load_conf_info(str(p))
    assert section_with(cfg, "checkpoint_filepath") == "inference"             # reference bug worked around
    with pytest.raises(configparser.Error):
        section_with(load_conf_info(str(tmp_path / "missing.cfg")), "checkpoint_filepath")


def test_partition_and_plan_metadata():
    from fullycnnspeechenhancement_b200.engine import num_frames, partition_utterances
    lens = np.full(1024, 32000)
    parts = partition_utterances(lens, 8)
    assert [len(p) for p in parts] == [128] * 8 and np.array_equal(np.concatenate(parts), np.arange(1024))
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 64001, 4096)
    parts = partition_utterances(lens, 8)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(4096))
    loads = np.array([num_frames(lens[p]).sum() for p in parts])
    assert loads.max() / loads.mean() < 1.002           # LPT balance on frames
    assert partition_utterances(lens, 1)[0].size == 4096


def test_product_path_fails_loudly_without_gpu_and_never_imports_oracle():
    import torch
    pkg = os.path.join(ROOT, "fullycnnspeechenhancement_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    if not torch.cuda.is_available():
        from fullycnnspeechenhancement_b200.engine import Enhancer
        with pytest.raises(_lib.RcedError):
            Enhancer("FullyCNNV2", fold.glorot_weights("FullyCNNV2"))
        lib = _lib.lib()
        h = ctypes.c_void_p()
        n = lib.rced_folded_weight_count(2)
        z = np.zeros(n, np.float32)
        rc = lib.rced_create(2, z.ctypes.data_as(ctypes.c_void_p), n, 0, ctypes.byref(h))
        assert rc != 0 and b"no CPU fallback" in lib.rced_last_error()
        from fullycnnspeechenhancement_b200.model_utils.model import FullyCNNSEModelV2
        with pytest.raises(NotImplementedError):
            FullyCNNSEModelV2(is_training=True)


def test_fft_emulator_matches_numpy():
    import fft_emulator as fe
    rng = np.random.default_rng(0)
    x = rng.normal(size=256)
    assert np.abs(fe.rfft256_via128(x) - np.fft.rfft(x, 256)).max() < 1e-12
    Y = rng.normal(size=129) + 1j * rng.normal(size=129)
    assert np.abs(fe.irfft512_first256(Y) - np.fft.irfft(Y, 512)[:256]).max() < 1e-14
    assert np.abs(fe.irfft256_full(Y) - np.fft.irfft(Y, 256)).max() < 1e-14


def test_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def _write_manifest(path, items):
    import json
    with open(path, "w") as f:
        for it in items:
            f.write(json.dumps(it) + "\n")


def test_dataset_sampler_loader_follow_the_reference(tmp_path):
    """Manifest filtering, clean/noise pairing, Sampler bins and SDR / AverageMeter arithmetic
    (data_utils/data_loader.py:64-160, model_utils/utils.py:14-90 of the reference).  Where the
    reference tree is present its own classes are run beside ours on the same seeded draws."""
    from fullycnnspeechenhancement_b200.data_utils.data_loader import AudioParser, DataLoader, DataSet, Sampler
    from fullycnnspeechenhancement_b200.model_utils.utils import SDR, AverageMeter
    from oracle import ref_import
    speech = str(tmp_path / "speech.manifest")
    noise = str(tmp_path / "noise.manifest")
    _write_manifest(speech, [{"audio_filepath": "s%d.wav" % i, "duration": d}
                             for i, d in enumerate([0.3, 0.4, 1.0, 2.0, 5.0, 9.0, 0.39])])
    _write_manifest(noise, [{"audio_filepath": "n%d.wav" % i, "duration": 1.0} for i in range(2)])
    ds = DataSet(speech, noise, sample_rate=8000, min_duration=0.4, max_duration=5.0)
    assert [it["audio_filepath"] for it in ds.item_list] == ["s1.wav", "s2.wav", "s3.wav", "s4.wav"]
    assert [it["audio_filepath"] for it in ds.noise_list] == ["n0.wav", "n1.wav", "n0.wav", "n1.wav"]
    assert ds() is ds and len(ds) == 4 and ds.item_name(2) == "s3.wav"
    bad = str(tmp_path / "bad.manifest")
    open(bad, "w").write("{not json}\n")
    with pytest.raises(IOError):
        DataSet(bad, None)
    nodur = str(tmp_path / "nodur.manifest")
    _write_manifest(nodur, [{"audio_filepath": "a.wav"}])
    with pytest.raises(KeyError):
        DataSet(nodur, None)
    # waveform pairing without touching the STFT (no GPU here): patch the decoder
    waves = {"s%d.wav" % i: np.random.default_rng(i).normal(size=900 + 100 * i).astype(np.float32) for i in range(7)}
    waves.update({"n%d.wav" % i: np.random.default_rng(50 + i).normal(size=400 + 2000 * i).astype(np.float32) for i in range(2)})
    ds.load_audio = lambda path: (waves[path], 8000)
    ds.snr = 10.0
    np.random.seed(7)
    mix, clean = ds.load_pair(1)                       # s2 (1100 samples) + n1 (2400 samples: random crop)
    assert clean is waves["s2.wav"] and len(mix) == 1100
    got = 10 * np.log10(np.sum(clean.astype(np.float64) ** 2) / np.sum((mix - clean).astype(np.float64) ** 2))
    assert abs(got - 10.0) < 1e-4
    np.random.seed(7)
    mix0, _ = ds.load_pair(0)                          # s1 (1000) + n0 (400: doubled twice, then cut)
    assert len(mix0) == 1000
    if ref_import.available():
        RefParser = ref_import.load()[3]
        rp = RefParser(sample_rate=8000, snr=10.0)
        np.random.seed(7)
        assert np.array_equal(rp.add_noise(waves["s2.wav"], waves["n1.wav"]), mix)
        np.random.seed(7)
        assert np.array_equal(rp.add_noise(waves["s1.wav"], waves["n0.wav"]), mix0)
    # paired manifest
    paired = str(tmp_path / "paired.manifest")
    _write_manifest(paired, [{"clean_audio_filepath": "s0.wav", "mix_audio_filepath": "s1.wav", "duration": 1.0}])
    dp = DataSet(paired, None, sample_rate=8000)
    dp.load_audio = lambda path: (waves[path], 8000)
    m, c = dp.load_pair(0)
    assert m is waves["s1.wav"] and c is waves["s0.wav"] and dp.item_name(0) == "s0.wav"
    # default loader bins and the Sampler's list surgery
    ld = DataLoader(ds, 3)
    assert ld.bins == [[0, 1, 2], [3]] and len(ld) == 2 and ld.num_works == 2
    np.random.seed(0)
    sm = Sampler(ds, 3)                                 # extends 4 items to 6 with the last two
    assert len(ds) == 6 and [it["audio_filepath"] for it in ds.item_list[4:]] == ["s3.wav", "s4.wav"]
    assert sorted(sorted(b) for b in sm) == [[0, 1, 2], [3, 4, 5]] and len(sm) == 2 and sm.iter_num() == 2
    ds2 = DataSet(speech, None, sample_rate=8000, min_duration=0.0)
    Sampler(ds2, 3, drop_last=True)
    assert len(ds2) == 6
    ds3 = DataSet(speech, None, sample_rate=8000, min_duration=0.0)
    ds3.item_list = ds3.item_list[:6]
    Sampler(ds3, 3)                                      # divides evenly: the reference still appends a whole batch
    assert len(ds3) == 9
    # scoring helpers
    y = np.random.default_rng(1).normal(size=1000)
    e = y + 0.01 * np.random.default_rng(2).normal(size=1000)
    want = 10 * np.log10(np.power(y, 2).sum() / (np.power(e - y, 2).sum() + np.finfo(np.float32).eps))
    assert SDR()(y, e) == want
    assert np.isfinite(SDR()(y, y))                     # the epsilon keeps a perfect estimate finite
    am = AverageMeter()
    am.update(2.0)
    am.update(4.0, n=3)
    assert am.sum == 6.0 and am.count == 4 and am.avg == 1.5
    assert AudioParser(8000, 32, 16).window_s == 0.032


def test_streaming_bookkeeping_with_a_stand_in_engine():
    """StreamingEnhancer cuts halo'd pieces on the hop grid, returns every sample exactly once and in
    order, and keeps no more history than the look-back (checked with an engine that marks each
    sample with its absolute position, no GPU needed)."""
    from fullycnnspeechenhancement_b200.streaming import LOOK_AHEAD, LOOK_BACK, StreamingEnhancer

    class Echo(object):
        def __init__(self):
            self.calls = []

        def enhance(self, waves):
            self.calls.append(len(waves[0]))
            return [np.asarray(w, np.float32) * 2.0 for w in waves]

    rng = np.random.default_rng(0)
    x = rng.normal(size=50000).astype(np.float32)
    eng = Echo()
    st = StreamingEnhancer(eng, block=1000)            # rounded down to 896 = 7 hops
    assert st.block == 896 and st.latency_samples == 896 + LOOK_AHEAD
    got, pos, pushed = [], 0, 0
    while pos < len(x):
        n = int(rng.integers(1, 3000))
        out = st.push(x[pos:pos + n])
        pos += n
        pushed = min(pos, len(x))
        got.append(out)
        done = sum(len(g) for g in got)
        assert done % 896 == 0 and done <= max(0, pushed - LOOK_AHEAD)        # only final samples are returned ...
        assert pushed - done < 896 + LOOK_AHEAD + 3000                        # ... and without undue delay
        assert len(st._buf) <= LOOK_BACK + 896 + LOOK_AHEAD + 3000            # bounded history
    got.append(st.flush())
    y = np.concatenate(got)
    assert len(y) == len(x) and np.array_equal(y, 2.0 * x)
    assert max(eng.calls) <= LOOK_BACK + 16 * 128 + LOOK_AHEAD
    with pytest.raises(ValueError):
        StreamingEnhancer(eng, block=64)


def test_bench_has_no_collective_after_the_non_zero_ranks_leave():
    """bench.py lets ranks != 0 return once the timed sections are over; a barrier or all-reduce behind
    that point hangs every multi-GPU run (it happened once).  Checked on the source."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    main = src[src.index("def main():"):]
    leave = main.index("    if rank != 0:\n        if world > 1:\n            dist.destroy_process_group()\n        return")
    tail = main[leave + 10:]
    tail = tail[tail.index("return") + 6:]
    for call in ("barrier()", "max_over_ranks(", "dist.all_reduce", "dist.barrier", "dist.broadcast"):
        assert call not in tail, call


def test_entry_point_helpers_on_the_cpu(tmp_path):
    """test.py's loader construction from a cfg and infer.py's reference feed layout (the reshape quirk),
    without a GPU."""
    import types
    from fullycnnspeechenhancement_b200.config import load_conf_info
    from fullycnnspeechenhancement_b200.infer import _as_reference_feeds
    from fullycnnspeechenhancement_b200.test import build_loader
    manifest = str(tmp_path / "m.testset")
    _write_manifest(manifest, [{"clean_audio_filepath": "c%d.wav" % i, "mix_audio_filepath": "m%d.wav" % i, "duration": 1.0 + i}
                               for i in range(5)])
    cfg = tmp_path / "t.cfg"
    cfg.write_text("[testing]\nbatch_size=2\ncheckpoint_filepath=x\n[model]\nnet_arch=RCED\nnet_work=FullyCNNV2\n"
                   "[data]\ntest_manifest_path=%s\nsnr=5\nsample_rate=8000\nfeature_dim=129\nwindow_ms=32\nstride_ms=16\n"
                   "audio_save_path=%s\n" % (manifest, tmp_path / "out"))
    loader = build_loader(load_conf_info(str(cfg)), num_works=3)
    ds = loader.dataset
    assert len(ds) == 5 and ds.noise_manifest is None and ds.snr == 5.0 and ds.sample_rate == 8000 and ds.complex
    assert (ds.window_s, ds.stride_s) == (0.032, 0.016) and loader.bins == [[0, 1], [2, 3], [4]] and loader.num_works == 3
    # infer.py:57-61: [F,T] arrays reshaped (not transposed) to (1,T,F,1) / (1,T,F)
    rng = np.random.default_rng(0)
    spec = rng.normal(size=(129, 7)) + 1j * rng.normal(size=(129, 7))
    ext = types.SimpleNamespace(power_spectrum=stft.power_spectrum, divide_phase=stft.divide_phase)
    mag, ph = _as_reference_feeds(spec, ext)
    assert mag.shape == (1, 7, 129, 1) and ph.shape == (1, 7, 129)
    assert np.array_equal(mag.reshape(-1), np.abs(spec).reshape(-1))            # memory order kept: a reshape, no transpose
    assert np.allclose(ph.reshape(129, 7) * np.abs(spec), spec)


def test_host_tables_follow_the_reference_truncation():
    """Offset tables of rced_enhance_host: aligned starts, outputs truncated like rebuild_audio does -- the reference rebuilds
    (T+1)*128 samples and slices them to len(clean_sig[i]) (model_utils/tester.py:107-113, utils.py:181-182), so an
    output can be LONGER than the noisy input but never longer than what was rebuilt."""
    from fullycnnspeechenhancement_b200.engine import host_tables, num_frames
    lens = np.array([1, 255, 256, 257, 1000, 32000, 4321])
    t = host_tables(lens)
    assert np.all(t["wav_off"] % 4 == 0) and np.array_equal(t["wav_len"], lens) and np.array_equal(t["out_len"], lens)
    assert np.all(t["wav_off"][1:] >= t["wav_off"][:-1] + lens[:-1]) and t["total"] >= t["wav_off"][-1] + lens[-1]
    rebuilt = (num_frames(lens) + 1) * 128
    want = np.array([5000, 100, 256, 300, 1010, 31000, 4400])          # len(clean) per utterance
    t2 = host_tables(lens, out_lens=want)
    assert np.array_equal(t2["out_len"], np.minimum(want, rebuilt))
    assert t2["out_len"][4] == 1010 > lens[4]                            # longer than the input, still inside (T+1)*128 = 1024
    assert np.all(t2["wav_off"][1:] >= t2["out_off"][:-1] + t2["out_len"][:-1])   # outputs never overlap the next utterance
    with pytest.raises(ValueError):
        host_tables(np.array([10, 0]))
    with pytest.raises(ValueError):
        host_tables(np.array([10, 20]), out_lens=[5])


def test_relay_plan():
    """Which ranks copy through which peer GPU.  The box of this project: with all eight ranks copying, GPUs 0-3 get
    8 GB/s per direction (their path to host memory is shared), GPUs 4-7 get 11 GB/s; with only GPUs 4-7 copying those reach
    > 20 GB/s: slow_i -> fast_i.  Uniform boxes, and boxes whose fast links cannot carry two streams, get no relay."""
    from fullycnnspeechenhancement_b200.engine import plan_relays, relay_candidates
    need = 9.4
    all_busy = [8.0, 8.1, 7.9, 8.0, 11.2, 11.1, 11.3, 11.2]
    slow, fast = relay_candidates(all_busy, need)
    assert slow == [2, 0, 3, 1] and fast == [6, 4, 7, 5]
    assert plan_relays(all_busy, need) == [-1] * 8                          # judged by the saturated figures alone: no
    got = plan_relays(all_busy, need, relay_gbs=[0, 0, 0, 0, 24.0, 23.0, 25.0, 24.5])
    assert sorted(got[:4]) == [4, 5, 6, 7] and got[4:] == [-1] * 4 and got[2] == 6   # the slowest rank gets the fastest relay
    assert plan_relays(all_busy, need, relay_gbs=[0, 0, 0, 0, 15.0, 15.0, 15.0, 15.0]) == [-1] * 8   # cannot carry two streams
    assert plan_relays([24.0] * 8, need) == [-1] * 8            # uniform and fast: nothing to do
    assert plan_relays([7.8] * 8, need) == [-1] * 8             # uniformly slow: nobody can help
    assert plan_relays([8.0, 30.0], need) == [1, -1]
