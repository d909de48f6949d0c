"""TensorFlow-1.x weight files without TensorFlow.

The reference restores weights with ``tf.train.Saver(tf.global_variables()).restore(sess,
checkpoint_filepath)`` (model_utils/tester.py:36-39) and exports them with
``graph_util.convert_variables_to_constants`` (freeze.py:42-47).  The on-disk formats are
TensorFlow's; they are restated here from the format definitions (tensor_bundle.proto,
table format of tensorflow/core/lib/io, graph.proto):

* checkpoint "V2" tensor bundle: ``<prefix>.index`` is an immutable sorted string table
  (LevelDB table format, uncompressed blocks) mapping ``""`` to a BundleHeaderProto and every
  variable name to a BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c};
  ``<prefix>.data-00000-of-00001`` holds the raw little-endian tensors.
* frozen graph: a serialized GraphDef whose ``Const`` nodes carry the variables under their
  variable names.

No checkpoint ships with the reference, so these readers are verified by round trip against
the writers below (tests/test_host_cpu.py) -- and the writers follow the same specification,
so a real TF-written file remains unverified here.
"""
import os
import struct

import numpy as np

# ----------------------------------------------------------------------------- protobuf wire
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8"), 2: np.dtype("<f8")}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def pb_fields(buf):
    """Iterate (field_number, wire_type, value) over one protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _pb_len(field, payload):
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def _pb_int(field, v):
    return _enc_varint(field << 3) + _enc_varint(v)


def _shape_proto(shape):
    return b"".join(_pb_len(2, _pb_int(1, int(d))) for d in shape)


def _parse_shape(buf):
    dims = []
    for f, _, v in pb_fields(buf):
        if f == 2:
            size = 0
            for f2, _, v2 in pb_fields(v):
                if f2 == 1:
                    size = v2
            dims.append(size)
    return tuple(dims)


# ----------------------------------------------------------------------------- crc32c
_CRC_TABLE = None


def crc32c(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = tbl
    crc ^= 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- table (.index)
_TABLE_MAGIC = 0xDB4775248B80FB57


def _read_block(buf, offset, size, verify=True):
    data = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed table block (type %d) is not supported; TensorFlow writes "
                         "checkpoint indices uncompressed" % ctype)
    if verify:
        stored = struct.unpack("<I", buf[offset + size + 1:offset + size + 5])[0]
        if stored != masked_crc32c(buf[offset:offset + size + 1]):
            raise ValueError("checkpoint index block checksum mismatch")
    n_restarts = struct.unpack("<I", data[-4:])[0]
    end = len(data) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(data, pos)
        non_shared, pos = _varint(data, pos)
        vlen, pos = _varint(data, pos)
        key = key[:shared] + data[pos:pos + non_shared]
        pos += non_shared
        out.append((key, data[pos:pos + vlen]))
        pos += vlen
    return out


def _read_table(buf, verify=True):
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _TABLE_MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = buf[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = []
    for _, handle in _read_block(buf, idx_off, idx_size, verify):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        entries.extend(_read_block(buf, off, size, verify))
    return entries


def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _write_table(items, block_size=4096):
    """items: sorted list of (key bytes, value bytes)."""
    out = bytearray()
    index = []

    def emit(block_items):
        blk = _build_block(block_items)
        off = len(out)
        out.extend(blk)
        out.append(0)
        out.extend(struct.pack("<I", masked_crc32c(blk + b"\x00")))
        return off, len(blk)

    cur, cur_bytes = [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            off, size = emit(cur)
            index.append((cur[-1][0], _enc_varint(off) + _enc_varint(size)))
            cur, cur_bytes = [], 0
    if cur:
        off, size = emit(cur)
        index.append((cur[-1][0], _enc_varint(off) + _enc_varint(size)))
    meta_off, meta_size = emit([])
    # index block: restart at every entry, as the table builder does for index blocks
    blk = _build_block(index, restart_interval=1)
    idx_off = len(out)
    out.extend(blk)
    out.append(0)
    out.extend(struct.pack("<I", masked_crc32c(blk + b"\x00")))
    footer = _enc_varint(meta_off) + _enc_varint(meta_size) + _enc_varint(idx_off) + _enc_varint(len(blk))
    footer += b"\x00" * (40 - len(footer))
    out.extend(footer + struct.pack("<Q", _TABLE_MAGIC))
    return bytes(out)


# ----------------------------------------------------------------------------- tensor bundle
def read_checkpoint(prefix, names=None, verify=True):
    """Read variables from a V2 checkpoint ``prefix`` (as passed to Saver.restore).
    ``names``: iterable of variable names to return (default: all float tensors).
    Missing files raise FileNotFoundError; a requested variable that is absent raises KeyError
    (TensorFlow's restore raises NotFoundError)."""
    index_path = prefix + ".index"
    if not os.path.exists(index_path):
        raise FileNotFoundError("checkpoint index %s not found (checkpoint_filepath is a prefix, "
                                "e.g. .../RCED_FullyCNNV2_0_9.ckpt)" % index_path)
    with open(index_path, "rb") as f:
        buf = f.read()
    entries = dict(_read_table(buf, verify))
    num_shards = 1
    for fno, _, v in pb_fields(entries.get(b"", b"")):
        if fno == 1:
            num_shards = v
        if fno == 2 and v != 0:
            raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    wanted = None if names is None else set(names)
    for key, val in entries.items():
        if key == b"":
            continue
        name = key.decode("utf-8")
        if wanted is not None and name not in wanted:
            continue
        dtype, shape, shard, offset, size, crc = 0, (), 0, 0, 0, None
        sliced = False
        for fno, wt, v in pb_fields(val):
            if fno == 1:
                dtype = v
            elif fno == 2:
                shape = _parse_shape(v)
            elif fno == 3:
                shard = v
            elif fno == 4:
                offset = v
            elif fno == 5:
                size = v
            elif fno == 6:
                crc = struct.unpack("<I", v)[0]
            elif fno == 7:
                sliced = True
        if sliced:
            raise ValueError("partitioned variable %s is not supported" % name)
        if dtype not in _DTYPES:
            if wanted is None:
                continue
            raise ValueError("variable %s has unsupported dtype %d" % (name, dtype))
        if shard not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)
            with open(path, "rb") as f:
                shards[shard] = f.read()
        raw = shards[shard][offset:offset + size]
        if len(raw) != size:
            raise ValueError("checkpoint data file truncated at variable %s" % name)
        if verify and crc is not None and masked_crc32c(raw) != crc:
            raise ValueError("checksum mismatch for variable %s" % name)
        out[name] = np.frombuffer(raw, dtype=_DTYPES[dtype]).reshape(shape).copy()
    if wanted is not None:
        missing = sorted(wanted - set(out))
        if missing:
            raise KeyError("variables not found in checkpoint %s: %s" % (prefix, ", ".join(missing)))
    return out


def write_checkpoint(prefix, variables):
    """Write a single-shard V2 bundle (``.index`` + ``.data-00000-of-00001``) plus the
    ``checkpoint`` state file, in the layout Saver.save produces."""
    d = os.path.dirname(prefix)
    if d and not os.path.isdir(d):
        os.makedirs(d)
    data = bytearray()
    items = []
    for name in sorted(variables):
        arr = np.asarray(variables[name])
        if arr.dtype == np.float32:
            dt = DT_FLOAT
        elif arr.dtype == np.int64:
            dt = DT_INT64
        elif arr.dtype == np.int32:
            dt = DT_INT32
        else:
            raise ValueError("unsupported dtype %s for %s" % (arr.dtype, name))
        raw = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes()
        entry = _pb_int(1, dt) + _pb_len(2, _shape_proto(arr.shape))
        if len(data):
            entry += _pb_int(4, len(data))
        entry += _pb_int(5, len(raw)) + _enc_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw))
        items.append((name.encode("utf-8"), entry))
        data.extend(raw)
    header = _pb_int(1, 1) + _pb_len(3, _pb_int(1, 1))      # num_shards=1, version{producer=1}
    items = [(b"", header)] + items
    with open(prefix + ".index", "wb") as f:
        f.write(_write_table(items))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    base = os.path.basename(prefix)
    with open(os.path.join(d or ".", "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


# ----------------------------------------------------------------------------- GraphDef (.pb)
def _parse_tensor(buf):
    dtype, shape, content, floats = 0, (), None, []
    for fno, wt, v in pb_fields(buf):
        if fno == 1:
            dtype = v
        elif fno == 2:
            shape = _parse_shape(v)
        elif fno == 4:
            content = bytes(v)
        elif fno == 5:
            if wt == 2:
                floats.extend(np.frombuffer(bytes(v), "<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
    if dtype != DT_FLOAT:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None and len(content):
        arr = np.frombuffer(content, "<f4").copy()
    elif len(floats) == 1 and n > 1:
        arr = np.full(n, floats[0], np.float32)
    else:
        arr = np.array(floats, np.float32)
    return arr.reshape(shape)


def read_frozen_graph(pb_path, names=None):
    """Float ``Const`` nodes of a frozen GraphDef, keyed by node name (== variable name after
    convert_variables_to_constants, freeze.py:42-47)."""
    with open(pb_path, "rb") as f:
        buf = f.read()
    out = {}
    wanted = None if names is None else set(names)
    for fno, wt, node in pb_fields(buf):
        if fno != 1 or wt != 2:
            continue
        name, op, tensor = None, None, None
        for f2, _, v in pb_fields(node):
            if f2 == 1:
                name = bytes(v).decode("utf-8")
            elif f2 == 2:
                op = bytes(v).decode("utf-8")
            elif f2 == 5:
                key, val = None, None
                for f3, _, v3 in pb_fields(v):
                    if f3 == 1:
                        key = bytes(v3)
                    elif f3 == 2:
                        val = v3
                if key == b"value" and val is not None:
                    for f4, _, v4 in pb_fields(val):
                        if f4 == 8:
                            tensor = v4
        if op == "Const" and tensor is not None and (wanted is None or name in wanted):
            arr = _parse_tensor(tensor)
            if arr is not None:
                out[name] = arr
    if wanted is not None:
        missing = sorted(wanted - set(out))
        if missing:
            raise KeyError("constants not found in %s: %s" % (pb_path, ", ".join(missing)))
    return out


def _attr(key, value_bytes):
    return _pb_len(5, _pb_len(1, key.encode()) + _pb_len(2, value_bytes))


def _node(name, op, inputs=(), attrs=()):
    msg = _pb_len(1, name.encode()) + _pb_len(2, op.encode())
    for i in inputs:
        msg += _pb_len(3, i.encode())
    for a in attrs:
        msg += a
    return _pb_len(1, msg)


def _const_node(name, arr):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    tensor = _pb_int(1, DT_FLOAT) + _pb_len(2, _shape_proto(arr.shape)) + _pb_len(4, arr.tobytes())
    return _node(name, "Const", (), [_attr("dtype", _pb_int(6, DT_FLOAT)), _attr("value", _pb_len(8, tensor))])


def write_frozen_graph(pb_path, net_work, weights):
    """Write an inference GraphDef in the shape freeze.py produces: placeholder ``input``, one
    ``Const`` per variable (named like the variable) with its ``/read`` Identity, and the op chain
    Conv2D -> BiasAdd -> FusedBatchNorm -> add -> Relu per layer.  Only the Const nodes are read
    back by this package; the op nodes document the graph for other consumers."""
    from . import fold
    from .. import _lib
    import ctypes
    a = fold.arch_id(net_work)
    f32 = _attr("T", _pb_int(6, DT_FLOAT))
    nodes = [_node("input", "Placeholder", (), [_attr("dtype", _pb_int(6, DT_FLOAT))])]
    prev = "input"
    outputs = {}
    lib = _lib.lib()
    scopes = fold.layer_scopes(net_work)
    skip_of = _skip_table(net_work)
    for i, (scope, bn) in enumerate(scopes):
        for v in ("kernel", "bias"):
            nodes.append(_const_node("%s/%s" % (scope, v), weights["%s/%s" % (scope, v)]))
            nodes.append(_node("%s/%s/read" % (scope, v), "Identity", ["%s/%s" % (scope, v)], [f32]))
        nodes.append(_node(scope + "/Conv2D", "Conv2D", [prev, scope + "/kernel/read"],
                           [f32, _attr("padding", _pb_len(2, b"SAME")), _attr("data_format", _pb_len(2, b"NHWC"))]))
        nodes.append(_node(scope + "/BiasAdd", "BiasAdd", [scope + "/Conv2D", scope + "/bias/read"], [f32]))
        cur = scope + "/BiasAdd"
        if bn:
            ins = [cur]
            for v in ("gamma", "beta", "moving_mean", "moving_variance"):
                n = "%s/batch_norm/%s" % (scope, v)
                nodes.append(_const_node(n, weights[n]))
                nodes.append(_node(n + "/read", "Identity", [n], [f32]))
                ins.append(n + "/read")
            cur = scope + "/batch_norm/FusedBatchNorm"
            nodes.append(_node(cur, "FusedBatchNorm", ins,
                               [f32, _attr("epsilon", _enc_varint((4 << 3) | 5) + struct.pack("<f", fold.BN_EPSILON)),
                                _attr("is_training", _pb_int(5, 0)), _attr("data_format", _pb_len(2, b"NHWC"))]))
        skip, after = skip_of.get(scope, (None, False))
        is_last = i == len(scopes) - 1
        if skip and not after:
            nodes.append(_node("add_%d" % i, "Add", [cur, outputs[skip]], [f32]))
            cur = "add_%d" % i
        if not is_last:
            nodes.append(_node("Relu_%d" % i, "Relu", [cur], [f32]))
            cur = "Relu_%d" % i
        if skip and after:
            nodes.append(_node("add_%d" % i, "Add", [cur, outputs[skip]], [f32]))
            cur = "add_%d" % i
        outputs[scope] = cur
        prev = cur
    with open(pb_path, "wb") as f:
        f.write(b"".join(nodes))
    return len(nodes)


def _skip_table(net_work):
    """scope -> (scope whose output is added, added_after_relu)   (model.py:19-22,48-54,75-76,86-87)"""
    from . import fold
    a = fold.arch_id(net_work)
    if a == 2:
        return {"decode_%d" % i: ("encode_%d" % (8 - i), False) for i in range(1, 8)}
    if a == 3:
        return {"CD1_decode": ("CE2_decode", True), "CD2_decode": ("CE1_decode", True)}
    return {"decode_%d" % i: ("encode_%d" % (5 - i), False) for i in range(1, 5)}


def load_weights(path, net_work):
    """Load the variables of `net_work` from a checkpoint prefix or a frozen ``.pb``
    (what BaseTester._load_checkpoint restores, model_utils/tester.py:36-39)."""
    from . import fold
    names = fold.variable_names(net_work)
    if path.endswith(".pb"):
        return read_frozen_graph(path, names)
    return read_checkpoint(path, names)
