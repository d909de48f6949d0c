"""Layer tables of the three reference models and batch-norm folding.

The tables restate model_utils/model.py:6-29 (FullyCNNSEModel), :32-61 (FullyCNNSEModelV2)
and :64-96 (FullyCNNSEModelV3) of the reference as data: scope names are the TensorFlow
variable scopes the reference creates (model_utils/module.py:27,29), which is also what a
checkpoint or a freeze.py export is keyed by.  The numeric shapes are cross-checked against
the tables compiled into librced_b200.so (csrc/rced_arch.cuh) at import of the engine.
"""
import numpy as np

BN_EPSILON = 1e-3          # tf.layers.batch_normalization default (module.py:29)
ARCH_IDS = {"FullyCNN": 1, "FullyCNNV2": 2, "FullyCNNV3": 3}


def arch_id(net_work):
    """tester.py:76-82: anything that is not V2/V3 falls back to FullyCNN (V1)."""
    return ARCH_IDS.get(net_work, 1)


def layer_scopes(net_work):
    """[(scope, has_batch_norm)] in execution order."""
    a = arch_id(net_work)
    if a == 2:
        names = ["encode_%d" % i for i in range(1, 9)] + ["decode_%d" % i for i in range(1, 9)]
    elif a == 3:
        names = []
        for blk in ("CE1", "CE2", "CE3", "CD1", "CD2"):
            names += [blk + "_encode_1", blk + "_encode_2", blk + "_decode"]
        names.append("decode_final")
    else:
        names = ["encode_1", "encode_2", "encode_3", "encode_4", "encode_8",
                 "decode_1", "decode_2", "decode_3", "decode_4", "decode_5"]
    return [(n, i != len(names) - 1) for i, n in enumerate(names)]


def output_node_name(net_work):
    """freeze.py:31-37 (with the V3 name corrected: its last scope is decode_final)."""
    a = arch_id(net_work)
    return {1: "decode_5/BiasAdd", 2: "decode_8/BiasAdd", 3: "decode_final/BiasAdd"}[a]


def variable_names(net_work):
    """All variables an inference checkpoint of `net_work` must contain."""
    out = []
    for scope, bn in layer_scopes(net_work):
        out += [scope + "/kernel", scope + "/bias"]
        if bn:
            out += [scope + "/batch_norm/" + n for n in ("gamma", "beta", "moving_mean", "moving_variance")]
    return out


def trainable_parameter_count(weights, net_work):
    """What BaseTester.param_count prints (tester.py:41-47): kernel, bias, gamma, beta."""
    n = 0
    for scope, bn in layer_scopes(net_work):
        n += weights[scope + "/kernel"].size + weights[scope + "/bias"].size
        if bn:
            n += weights[scope + "/batch_norm/gamma"].size + weights[scope + "/batch_norm/beta"].size
    return int(n)


def fold_batch_norm(weights, net_work):
    """Fold inference-mode batch norm into each conv:  W' = W * s,  b' = (b - mean) * s + beta,
    s = gamma / sqrt(var + 1e-3)  (computed in float64, stored as float32).

    Returns the flat float32 vector rced_create expects: per layer the HWIO kernel then the bias.
    Missing variables raise KeyError (TensorFlow's restore raises NotFoundError)."""
    parts = []
    for scope, bn in layer_scopes(net_work):
        k = np.asarray(weights[scope + "/kernel"], dtype=np.float64)
        b = np.asarray(weights[scope + "/bias"], dtype=np.float64)
        if bn:
            g = np.asarray(weights[scope + "/batch_norm/gamma"], dtype=np.float64)
            beta = np.asarray(weights[scope + "/batch_norm/beta"], dtype=np.float64)
            mean = np.asarray(weights[scope + "/batch_norm/moving_mean"], dtype=np.float64)
            var = np.asarray(weights[scope + "/batch_norm/moving_variance"], dtype=np.float64)
            s = g / np.sqrt(var + BN_EPSILON)
            k = k * s                      # broadcast over the cout (last) axis
            b = (b - mean) * s + beta
        parts.append(k.astype(np.float32).ravel())
        parts.append(b.astype(np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts))


def glorot_weights(net_work, seed=0):
    """Random-init weights as the reference's graph would create them: Glorot-uniform kernels
    (tf.layers.conv2d default), zero bias, identity batch norm (gamma 1, beta 0, mean 0, var 1)."""
    from .. import _lib
    rng = np.random.default_rng(seed)
    a = arch_id(net_work)
    import ctypes
    w = {}
    for i, (scope, bn) in enumerate(layer_scopes(net_work)):
        kh, kw, cin, cout = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
        _lib.check(_lib.lib().rced_layer_shape(a, i, kh, kw, cin, cout))
        kh, kw, cin, cout = kh.value, kw.value, cin.value, cout.value
        limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
        w[scope + "/kernel"] = rng.uniform(-limit, limit, (kh, kw, cin, cout)).astype(np.float32)
        w[scope + "/bias"] = np.zeros(cout, np.float32)
        if bn:
            w[scope + "/batch_norm/gamma"] = np.ones(cout, np.float32)
            w[scope + "/batch_norm/beta"] = np.zeros(cout, np.float32)
            w[scope + "/batch_norm/moving_mean"] = np.zeros(cout, np.float32)
            w[scope + "/batch_norm/moving_variance"] = np.ones(cout, np.float32)
    return w
