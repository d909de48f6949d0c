"""TEST INFRASTRUCTURE -- a stand-in for the three TensorFlow-1.x calls the reference's model code makes, so that the
UNMODIFIED model definitions of /root/reference/model_utils/model.py can be executed in the authoring container
(TensorFlow 1.14 is not installable here).  Only tests/ and tests/golden/make_golden.py use it.

What this pins and what it does not: running the reference's own `FullyCNNSEModel{,V2,V3}.__call__` through this module
takes the WIRING from the reference itself -- the layer sequence, widths, kernel sizes, which tensors are skip inputs, where
the addition sits relative to BN and ReLU (model_utils/module.py:27-34), and the variable scopes (hence the checkpoint
names, e.g. V1's fifth encoder layer living in scope "encode_8", model_utils/model.py:15).  The arithmetic of the three
operations is restated here from TensorFlow's documented behaviour, independently of oracle/network.py:

* tf.layers.conv2d(inputs, filters, kernel_size, strides, padding, name): NHWC cross-correlation with an HWIO kernel
  `<name>/kernel` and `<name>/bias`; "SAME" at stride 1 pads k - 1 in total, (k - 1) // 2 in front (the extra row / column
  of an even kernel goes behind);
* tf.layers.batch_normalization(x, training=False, name): (x - moving_mean) / sqrt(moving_variance + 1e-3) * gamma + beta
  with the variables `<name>/{gamma,beta,moving_mean,moving_variance}` (epsilon 1e-3 is the layer's default);
* tf.nn.relu.

Tensors are float64 numpy arrays.  `requested()` lists the variables the model asked for, in creation order.
"""
import sys
import types

import numpy as np

_STORE = {"vars": None, "requested": []}


def set_variables(weights):
    """weights: dict TensorFlow variable name -> array (the dict layout of oracle.network.random_weights)."""
    _STORE["vars"] = {k: np.asarray(v, np.float64) for k, v in weights.items()}
    _STORE["requested"] = []


def requested():
    return list(_STORE["requested"])


def _var(name, shape):
    if _STORE["vars"] is None:
        raise RuntimeError("tf_standin.set_variables() first")
    if name not in _STORE["vars"]:
        raise KeyError("the reference's model asks for variable %r, which the weight dict does not hold" % name)
    v = _STORE["vars"][name]
    if tuple(v.shape) != tuple(shape):
        raise ValueError("variable %r: the reference's model needs shape %r, the weight dict holds %r" % (name, tuple(shape), v.shape))
    _STORE["requested"].append((name, tuple(shape)))
    return v


def _conv2d(inputs, filters, kernel_size, strides=(1, 1), padding="valid", name=None, **kwargs):
    import torch
    import torch.nn.functional as F
    if kwargs:
        raise NotImplementedError("tf.layers.conv2d stand-in: unexpected arguments %r" % sorted(kwargs))
    if tuple(strides) != (1, 1):
        raise NotImplementedError("stride %r" % (strides,))
    x = np.asarray(inputs, np.float64)
    kh, kw = kernel_size
    cin = x.shape[-1]
    k = _var(name + "/kernel", (kh, kw, cin, filters))
    b = _var(name + "/bias", (filters,))
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))          # NCHW
    if padding.upper() == "SAME":
        ph, pw = kh - 1, kw - 1
        xt = F.pad(xt, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))            # (left, right, top, bottom)
    elif padding.upper() != "VALID":
        raise ValueError(padding)
    wt = torch.from_numpy(np.ascontiguousarray(k.transpose(3, 2, 0, 1)))          # OIHW; conv2d is a cross-correlation
    y = F.conv2d(xt, wt, torch.from_numpy(b.copy()))
    return y.numpy().transpose(0, 2, 3, 1)


def _batch_normalization(inputs, training=False, name=None, **kwargs):
    if kwargs:
        raise NotImplementedError("tf.layers.batch_normalization stand-in: unexpected arguments %r" % sorted(kwargs))
    if training:
        raise NotImplementedError("the stand-in runs the inference graph only (training=False)")
    x = np.asarray(inputs, np.float64)
    c = x.shape[-1]
    gamma, beta = _var(name + "/gamma", (c,)), _var(name + "/beta", (c,))
    mean, var = _var(name + "/moving_mean", (c,)), _var(name + "/moving_variance", (c,))
    return (x - mean) / np.sqrt(var + 1e-3) * gamma + beta


def _relu(x):
    return np.maximum(np.asarray(x, np.float64), 0.0)


def install():
    """Registers the stand-in as `tensorflow` (and the `tensorflow.contrib.slim` the reference imports for its unused
    separable_conv) unless a real TensorFlow is importable.  Returns the module."""
    if "tensorflow" in sys.modules and not getattr(sys.modules["tensorflow"], "_rced_standin", False):
        raise RuntimeError("a real tensorflow is loaded: use it instead of the stand-in")
    tf = types.ModuleType("tensorflow")
    tf._rced_standin = True
    tf.layers = types.SimpleNamespace(conv2d=_conv2d, batch_normalization=_batch_normalization)
    tf.nn = types.SimpleNamespace(relu=_relu)
    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")
    contrib.slim = slim
    tf.contrib = contrib
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.slim"] = slim
    return tf
