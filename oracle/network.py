"""CPU restatement of the reference network graph (test oracle; wiring pinned against the reference's executed model
code, the arithmetic of TensorFlow's three operations PARITY UNPINNED).

Restates /root/reference/model_utils/module.py:11-34 (conv_bn_relu) and
/root/reference/model_utils/model.py:6-96 (R-CED V1/V2, CR-CED V3) with the
TensorFlow-1.14 semantics those call sites imply:

* ``tf.layers.conv2d(x, cout, (kh,kw), (1,1), 'SAME')``: cross-correlation, NHWC,
  HWIO kernel, bias add; SAME pads (k-1)//2 before and k-1-(k-1)//2 after
  (kh=8 -> 3 before / 4 after).
* ``tf.layers.batch_normalization(training=False)``: (x-mean)*gamma/sqrt(var+1e-3)+beta.
* then ``+ skip_input``, then ``relu`` (module.py:30-33).

TensorFlow is a third-party dependency that is absent here (pinned
``tensorflow-gpu==1.14.0`` in requriements.txt:4), so this file cannot be checked
against the reference's own execution under TensorFlow; see oracle/__init__.py.  The wiring IS checked against the
reference's model classes, executed with oracle/tf_standin.py in place of TensorFlow
(tests/test_oracle_golden.py::test_network_oracle_matches_the_reference_model_code).
"""
import numpy as np

BN_EPS = 1e-3  # tf.layers.batch_normalization default epsilon


def _L(scope, kh, kw, cin, cout, norm=True, act=True, skip=None, skip_after_act=False):
    return dict(scope=scope, kh=kh, kw=kw, cin=cin, cout=cout, norm=norm, act=act,
                skip=skip, skip_after_act=skip_after_act)


def layer_table(net_work):
    """Layer list in execution order.  ``skip`` names the scope whose *output* is
    added (module.py:30-31); ``skip_after_act`` marks V3 blocks where the add comes
    after the ReLU with no ReLU afterwards (model.py:75-76)."""
    if net_work == "FullyCNNV2":        # model.py:32-61
        return [
            _L("encode_1", 8, 11, 1, 10), _L("encode_2", 1, 7, 10, 12),
            _L("encode_3", 1, 5, 12, 14), _L("encode_4", 1, 5, 14, 15),
            _L("encode_5", 1, 5, 15, 19), _L("encode_6", 1, 5, 19, 21),
            _L("encode_7", 1, 7, 21, 23), _L("encode_8", 1, 11, 23, 25),
            _L("decode_1", 1, 7, 25, 23, skip="encode_7"), _L("decode_2", 1, 5, 23, 21, skip="encode_6"),
            _L("decode_3", 1, 5, 21, 19, skip="encode_5"), _L("decode_4", 1, 5, 19, 15, skip="encode_4"),
            _L("decode_5", 1, 5, 15, 14, skip="encode_3"), _L("decode_6", 1, 7, 14, 12, skip="encode_2"),
            _L("decode_7", 1, 11, 12, 10, skip="encode_1"),
            _L("decode_8", 1, 129, 10, 1, norm=False, act=False),
        ]
    if net_work == "FullyCNNV3":        # model.py:64-96
        table = []

        def block(name, first_kernel, cin, skip=None):   # simple_RCED, model.py:68-78
            table.append(_L(name + "_encode_1", first_kernel[0], first_kernel[1], cin, 18))
            table.append(_L(name + "_encode_2", 1, 5, 18, 30))
            table.append(_L(name + "_decode", 1, 9, 30, 8, skip=skip, skip_after_act=True))
        block("CE1", (8, 9), 1)
        block("CE2", (1, 9), 8)
        block("CE3", (1, 9), 8)
        block("CD1", (1, 9), 8, skip="CE2_decode")
        block("CD2", (1, 9), 8, skip="CE1_decode")
        table.append(_L("decode_final", 1, 129, 8, 1, norm=False, act=False))
        return table
    # default / "FullyCNN": model.py:6-29 (5th encoder scope really is "encode_8", :15)
    return [
        _L("encode_1", 8, 13, 1, 12), _L("encode_2", 1, 11, 12, 16),
        _L("encode_3", 1, 9, 16, 20), _L("encode_4", 1, 7, 20, 24),
        _L("encode_8", 1, 7, 24, 32),
        _L("decode_1", 1, 7, 32, 24, skip="encode_4"), _L("decode_2", 1, 9, 24, 20, skip="encode_3"),
        _L("decode_3", 1, 11, 20, 16, skip="encode_2"), _L("decode_4", 1, 13, 16, 12, skip="encode_1"),
        _L("decode_5", 1, 129, 12, 1, norm=False, act=False),
    ]


def trainable_param_count(net_work):
    """kernel + bias + BN gamma/beta (moving stats are not trainable)."""
    n = 0
    for L in layer_table(net_work):
        n += L["kh"] * L["kw"] * L["cin"] * L["cout"] + L["cout"]
        if L["norm"]:
            n += 2 * L["cout"]
    return n


def random_weights(net_work, seed, randomize_bn=True):
    """Glorot-uniform kernels (tf.layers.conv2d default), zero bias as TF initialises
    it -- plus, for test strength, small random biases and randomised BN statistics
    when ``randomize_bn`` (so that BN folding is actually exercised)."""
    rng = np.random.default_rng(seed)
    w = {}
    for L in layer_table(net_work):
        s = L["scope"]
        fan_in = L["kh"] * L["kw"] * L["cin"]
        fan_out = L["kh"] * L["kw"] * L["cout"]
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        w[s + "/kernel"] = rng.uniform(-limit, limit, (L["kh"], L["kw"], L["cin"], L["cout"])).astype(np.float32)
        if randomize_bn:
            w[s + "/bias"] = rng.normal(0, 0.05, L["cout"]).astype(np.float32)
        else:
            w[s + "/bias"] = np.zeros(L["cout"], np.float32)
        if L["norm"]:
            if randomize_bn:
                w[s + "/batch_norm/gamma"] = rng.uniform(0.5, 1.5, L["cout"]).astype(np.float32)
                w[s + "/batch_norm/beta"] = rng.normal(0, 0.1, L["cout"]).astype(np.float32)
                w[s + "/batch_norm/moving_mean"] = rng.normal(0, 0.1, L["cout"]).astype(np.float32)
                w[s + "/batch_norm/moving_variance"] = rng.uniform(0.5, 1.5, L["cout"]).astype(np.float32)
            else:
                w[s + "/batch_norm/gamma"] = np.ones(L["cout"], np.float32)
                w[s + "/batch_norm/beta"] = np.zeros(L["cout"], np.float32)
                w[s + "/batch_norm/moving_mean"] = np.zeros(L["cout"], np.float32)
                w[s + "/batch_norm/moving_variance"] = np.ones(L["cout"], np.float32)
    return w


def conv2d_same_nhwc(x, kernel, bias):
    """Tap-loop SAME cross-correlation, NHWC x HWIO, in the dtype of ``x``."""
    kh, kw, cin, cout = kernel.shape
    n, h, wd, _ = x.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((n, h + kh - 1, wd + kw - 1, cin), x.dtype)
    xp[:, pt:pt + h, pl:pl + wd, :] = x
    out = np.zeros((n, h, wd, cout), x.dtype)
    k = kernel.astype(x.dtype)
    for dh in range(kh):
        for dw in range(kw):
            out += xp[:, dh:dh + h, dw:dw + wd, :] @ k[dh, dw]
    return out + bias.astype(x.dtype)


def conv_bn_relu(x, L, w, outputs):
    """module.py:11-34 for one layer-table row."""
    s = L["scope"]
    y = conv2d_same_nhwc(x, w[s + "/kernel"], w[s + "/bias"])
    dt = x.dtype
    if L["norm"]:
        g = w[s + "/batch_norm/gamma"].astype(dt)
        b = w[s + "/batch_norm/beta"].astype(dt)
        m = w[s + "/batch_norm/moving_mean"].astype(dt)
        v = w[s + "/batch_norm/moving_variance"].astype(dt)
        y = (y - m) * (g / np.sqrt(v + dt.type(BN_EPS))) + b
    if L["skip"] is not None and not L["skip_after_act"]:
        y = y + outputs[L["skip"]]
    if L["act"]:
        y = np.maximum(y, 0)
    if L["skip"] is not None and L["skip_after_act"]:
        y = y + outputs[L["skip"]]
    return y


def forward(net_work, w, x, dtype=np.float64, return_all=False):
    """model.py ``__call__``: x [N,T,129,1] -> [N,T,129,1]."""
    x = np.asarray(x).astype(dtype)
    outputs = {}
    for L in layer_table(net_work):
        x = conv_bn_relu(x, L, w, outputs)
        outputs[L["scope"]] = x
    return (x, outputs) if return_all else x


def forward_torch(net_work, w, x, dtype="float32", num_threads=None):
    """Independent evaluator: torch CPU conv2d (NCHW/OIHW) with explicit SAME padding.
    Also the float32 multi-threaded network leg of the CPU baseline (stand-in for the
    TF-CPU Conv2D/Eigen path the reference runs under ``CUDA_VISIBLE_DEVICES=''``)."""
    import torch
    import torch.nn.functional as F
    if num_threads:
        torch.set_num_threads(num_threads)
    td = getattr(torch, dtype)
    with torch.no_grad():
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).to(td).permute(0, 3, 1, 2)  # N,C,T,F
        outs = {}
        for L in layer_table(net_work):
            s = L["scope"]
            k = torch.from_numpy(w[s + "/kernel"]).to(td).permute(3, 2, 0, 1).contiguous()
            b = torch.from_numpy(w[s + "/bias"]).to(td)
            pt, pl = (L["kh"] - 1) // 2, (L["kw"] - 1) // 2
            t = F.conv2d(F.pad(t, (pl, L["kw"] - 1 - pl, pt, L["kh"] - 1 - pt)), k, b)
            if L["norm"]:
                g = torch.from_numpy(w[s + "/batch_norm/gamma"]).to(td)
                be = torch.from_numpy(w[s + "/batch_norm/beta"]).to(td)
                m = torch.from_numpy(w[s + "/batch_norm/moving_mean"]).to(td)
                v = torch.from_numpy(w[s + "/batch_norm/moving_variance"]).to(td)
                sc = (g / torch.sqrt(v + BN_EPS)).view(1, -1, 1, 1)
                t = (t - m.view(1, -1, 1, 1)) * sc + be.view(1, -1, 1, 1)
            if L["skip"] is not None and not L["skip_after_act"]:
                t = t + outs[L["skip"]]
            if L["act"]:
                t = torch.relu(t)
            if L["skip"] is not None and L["skip_after_act"]:
                t = t + outs[L["skip"]]
            outs[s] = t
        return t.permute(0, 2, 3, 1).contiguous().numpy()


def mac_per_frame(net_work, valid_only=True, F=129):
    """Multiply-accumulates per interior frame.  ``valid_only``
    drops frequency taps that fall on SAME zero padding; time taps are all counted
    (interior frame)."""
    total = 0
    for L in layer_table(net_work):
        pl = (L["kw"] - 1) // 2
        if valid_only:
            taps = sum(1 for f in range(F) for k in range(L["kw"]) if 0 <= f + k - pl < F)
        else:
            taps = F * L["kw"]
        total += taps * L["kh"] * L["cin"] * L["cout"]
    return total
