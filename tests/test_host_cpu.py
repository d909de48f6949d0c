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
    for mnemonic in ("STTM", "LDTM", "UBLKCP", "FFMA"):
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
