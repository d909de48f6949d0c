"""CPU: model_utils/ckpt.py reading files it did not write.

No TensorFlow-written checkpoint exists (the reference ships none and TF 1.14 cannot be installed), so the
readers used to be verified by round trip against the writers of the same module only.
These fixtures come from producers that share no code with ckpt.py:

* a tensor-bundle ``.index`` built here, byte by byte, following the LevelDB table format TensorFlow uses
  (tensorflow/core/lib/io/table_builder.cc, format.cc; tensor_bundle.proto): data blocks with a restart
  interval of 2, so that most keys are prefix-compressed against their predecessor; several data blocks;
  index-block keys that are SHORTENED separators (not keys of the table); a metaindex block; entries of other
  dtypes (int64 global_step, the Adam slots of a training checkpoint) between the float tensors;
* a frozen GraphDef serialised by the protobuf library (tensorboard's TensorFlow protos) whose constants
  use ``float_val`` -- packed, one broadcast value, and mixed with ``tensor_content`` -- as
  ``convert_variables_to_constants`` + ``make_tensor_proto`` emit for small tensors (freeze.py:42-47).

Reference call sites: Saver.restore at model_utils/tester.py:36-39, freeze.py:31-48."""
import struct

import numpy as np
import pytest

from fullycnnspeechenhancement_b200.model_utils import ckpt, fold
from oracle import network

MASK_DELTA = 0xA282EAD8


def _crc32c(data):
    """Castagnoli CRC, bit by bit (no table, no code shared with ckpt.crc32c)."""
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xFFFFFFFF


def _masked(data):
    c = _crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


def _vi(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _field(no, wire, payload):
    return _vi((no << 3) | wire) + payload


def _bundle_entry(dtype, shape, offset, size, crc):
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto: repeated dim=2 {size=1}), shard_id=3, offset=4, size=5,
    crc32c=6 (fixed32)."""
    dims = b"".join(_field(2, 2, _vi(len(_field(1, 0, _vi(d)))) + _field(1, 0, _vi(d))) for d in shape)
    e = _field(1, 0, _vi(dtype)) + _field(2, 2, _vi(len(dims)) + dims)
    if offset:
        e += _field(4, 0, _vi(offset))
    e += _field(5, 0, _vi(size)) + _field(6, 5, struct.pack("<I", crc))
    return e


def _block(items, restart_interval):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    for r in restarts or [0]:
        out += struct.pack("<I", r)
    out += struct.pack("<I", max(1, len(restarts)))
    return bytes(out)


def _separator(a, b):
    """leveldb BytewiseComparator::FindShortestSeparator: a short key in [a, b)."""
    n = 0
    while n < min(len(a), len(b)) and a[n] == b[n]:
        n += 1
    if n < min(len(a), len(b)) and a[n] < 0xFF and a[n] + 1 < b[n]:
        return a[:n] + bytes([a[n] + 1])
    return a


def _table(items, per_block, restart_interval):
    out = bytearray()

    def put(blk):
        off = len(out)
        out.extend(blk + b"\x00" + struct.pack("<I", _masked(blk + b"\x00")))
        return _vi(off) + _vi(len(blk))

    index = []
    blocks = [items[i:i + per_block] for i in range(0, len(items), per_block)]
    for bi, blk_items in enumerate(blocks):
        handle = put(_block(blk_items, restart_interval))
        last = blk_items[-1][0]
        sep = _separator(last, blocks[bi + 1][0][0]) if bi + 1 < len(blocks) else last + b"\x00"   # (FindShortSuccessor-like)
        index.append((sep, handle))
    meta = put(_block([], 16))
    idx = put(_block(index, 1))
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57))
    return bytes(out)


DT_FLOAT, DT_INT64 = 1, 9


@pytest.mark.parametrize("name", ["FullyCNNV2", "FullyCNNV3"])
def test_reads_a_hand_built_training_checkpoint(tmp_path, name):
    w = network.random_weights(name, 5, True)
    rng = np.random.default_rng(0)
    tensors = dict(w)
    # what a training checkpoint holds besides (trainer.py:28,51,177): ignored by the loader
    tensors["global_step"] = np.array(1234, np.int64)
    tensors["beta1_power"] = np.array(0.9 ** 7, np.float32)
    for k in list(w):
        if k.endswith("/kernel"):
            tensors[k + "/Adam"] = rng.normal(size=w[k].shape).astype(np.float32)
            tensors[k + "/Adam_1"] = rng.normal(size=w[k].shape).astype(np.float32)
    data, entries = bytearray(), []
    for key in sorted(tensors):
        arr = tensors[key]
        raw = arr.tobytes()
        entries.append((key.encode(), _bundle_entry(DT_INT64 if arr.dtype == np.int64 else DT_FLOAT, arr.shape, len(data), len(raw), _masked(raw))))
        data += raw
    header = _field(1, 0, _vi(1)) + _field(3, 2, _vi(2) + _field(1, 0, _vi(1)))      # num_shards = 1, version.producer = 1
    items = [(b"", header)] + entries
    prefix = str(tmp_path / ("RCED_%s_3_9.ckpt" % name))
    open(prefix + ".index", "wb").write(_table(items, per_block=7, restart_interval=2))
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))

    got = ckpt.load_weights(prefix, name)
    assert set(got) == set(fold.variable_names(name))
    for k in got:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k]), k
    everything = ckpt.read_checkpoint(prefix)
    assert np.array_equal(everything["encode_1/kernel/Adam" if name == "FullyCNNV2" else "CE1_encode_1/kernel/Adam"],
                          tensors["encode_1/kernel/Adam" if name == "FullyCNNV2" else "CE1_encode_1/kernel/Adam"])
    assert np.array_equal(fold.fold_batch_norm(got, name), fold.fold_batch_norm(w, name))
    # a flipped byte in the data file is caught by the entry's crc32c
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        ckpt.read_checkpoint(prefix)


def test_reads_a_protobuf_written_frozen_graph_with_float_val(tmp_path):
    graph_pb2 = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    from tensorboard.compat.proto import tensor_pb2, tensor_shape_pb2
    name = "FullyCNNV2"
    w = network.random_weights(name, 9, True)
    w["encode_3/batch_norm/gamma"] = np.full_like(w["encode_3/batch_norm/gamma"], 0.75)      # -> one broadcast float_val
    g = graph_pb2.GraphDef()
    inp = g.node.add()
    inp.name, inp.op = "input", "Placeholder"
    for i, (k, arr) in enumerate(sorted(w.items())):
        n = g.node.add()
        n.name, n.op = k, "Const"
        n.attr["dtype"].type = 1
        t = tensor_pb2.TensorProto(dtype=1, tensor_shape=tensor_shape_pb2.TensorShapeProto(
            dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(d)) for d in arr.shape]))
        if k == "encode_3/batch_norm/gamma":
            t.float_val.append(0.75)                       # all elements equal: a single value
        elif arr.size <= 32:
            t.float_val.extend(arr.ravel().tolist())       # small tensors: repeated float_val (packed on the wire)
        else:
            t.tensor_content = arr.tobytes()
        n.attr["value"].tensor.CopyFrom(t)
        ident = g.node.add()                               # convert_variables_to_constants keeps the read ops
        ident.name, ident.op = k + "/read", "Identity"
        ident.input.append(k)
    out = g.node.add()
    out.name, out.op = fold.output_node_name(name), "BiasAdd"
    pb = str(tmp_path / "frozen.pb")
    open(pb, "wb").write(g.SerializeToString())
    got = ckpt.load_weights(pb, name)
    assert set(got) == set(fold.variable_names(name))
    for k in got:
        assert got[k].shape == w[k].shape and np.array_equal(got[k], w[k]), k
