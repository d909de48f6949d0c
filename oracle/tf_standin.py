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

Graph mode.  The reference's tester (model_utils/tester.py:24-90) builds the model on a `tf.placeholder`, restores a
checkpoint with `tf.train.Saver(...).restore(sess, path)` and evaluates with `sess.run(pred, feed_dict)`.  When an operation
receives a placeholder (or a node derived from one) it returns a lazy node; `Session.run` evaluates the node with the fed
value cast to the placeholder's dtype (float32, like TensorFlow's feed) and returns float32.  `Saver.restore` reads
`<checkpoint path>.standin.npz` (the variables by TensorFlow name: the checkpoint FORMAT is the business of
model_utils/ckpt.py and its own tests, not of this stand-in).  `tf.trainable_variables()` / `tf.shape(v.value()).eval()`
serve the reference's param_count().
"""
import sys
import types

import numpy as np

_STORE = {"vars": None, "requested": [], "created": []}


def set_variables(weights, keep_created=False):
    """weights: dict TensorFlow variable name -> array (the dict layout of oracle.network.random_weights)."""
    _STORE["vars"] = {k: np.asarray(v, np.float64) for k, v in weights.items()}
    _STORE["requested"] = []
    if not keep_created:
        _STORE["created"] = []


def reset_graph():
    _STORE["created"] = []


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


class Node(object):
    """A lazily evaluated tensor of the stand-in's graph mode."""

    def __init__(self, fn, parents=()):
        self.fn, self.parents = fn, tuple(parents)

    def evaluate(self, feeds, cache):
        if id(self) not in cache:
            cache[id(self)] = self.fn(*[p.evaluate(feeds, cache) if isinstance(p, Node) else p for p in self.parents], feeds=feeds)
        return cache[id(self)]

    def __add__(self, other):
        out = Node(lambda a, b, feeds=None: a + b, (self, other))
        out.channels = getattr(self, "channels", None)
        return out

    __radd__ = __add__


class Placeholder(Node):
    def __init__(self, shape, dtype, name):
        Node.__init__(self, None)
        self.shape, self.dtype, self.name = shape, dtype, name

    def evaluate(self, feeds, cache):
        for k, v in feeds.items():
            if k is self:
                x = np.asarray(v)
                if len(self.shape) != x.ndim or any(d is not None and d != n for d, n in zip(self.shape, x.shape)):
                    raise ValueError("Cannot feed value of shape %r for Tensor %r, which has shape %r" % (x.shape, self.name, self.shape))
                return x.astype(self.dtype).astype(np.float64)
        raise ValueError("placeholder %r was not fed" % self.name)


def _lazy(op):
    """op(x, ...) on arrays -> the same call returning a Node when x is one."""
    def call(x, *args, **kwargs):
        if isinstance(x, Node):
            return Node(lambda v, feeds=None: op(v, *args, **kwargs), (x,))
        return op(x, *args, **kwargs)
    return call


class Variable(object):
    def __init__(self, name, shape):
        self.name, self.shape = name + ":0", tuple(shape)

    def value(self):
        return self


class _Shape(object):
    def __init__(self, v):
        self.v = v

    def eval(self, session=None):
        return np.array(self.v.shape, np.int64)


class Saver(object):
    def __init__(self, var_list=None):
        self.var_list = var_list

    def restore(self, sess, path):
        with np.load(path + ".standin.npz") as z:
            set_variables({k: z[k] for k in z.files}, keep_created=True)


class Session(object):
    def __init__(self, config=None):
        self.config = config

    def as_default(self):
        return self

    def run(self, fetch, feed_dict=None):
        out = fetch.evaluate(feed_dict or {}, {})
        return np.asarray(out, np.float64).astype(np.float32)


def _conv2d_array(inputs, filters, kernel_size, strides=(1, 1), padding="valid", name=None, **kwargs):
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


def _batch_normalization_array(inputs, training=False, name=None, **kwargs):
    if kwargs:
        raise NotImplementedError("tf.layers.batch_normalization stand-in: unexpected arguments %r" % sorted(kwargs))
    if training:
        raise NotImplementedError("the stand-in runs the inference graph only (training=False)")
    x = np.asarray(inputs, np.float64)
    c = x.shape[-1]
    gamma, beta = _var(name + "/gamma", (c,)), _var(name + "/beta", (c,))
    mean, var = _var(name + "/moving_mean", (c,)), _var(name + "/moving_variance", (c,))
    return (x - mean) / np.sqrt(var + 1e-3) * gamma + beta


def _relu_array(x):
    return np.maximum(np.asarray(x, np.float64), 0.0)


def _declare(name, shape, trainable=True):
    _STORE["created"].append((Variable(name, shape), trainable))


def _conv2d(inputs, filters, kernel_size, strides=(1, 1), padding="valid", name=None, **kwargs):
    if isinstance(inputs, Node):      # graph mode: the variables are created now, read at run time
        cin = inputs.channels
        _declare(name + "/kernel", (kernel_size[0], kernel_size[1], cin, filters))
        _declare(name + "/bias", (filters,))
        out = _lazy(_conv2d_array)(inputs, filters, kernel_size, strides, padding, name=name, **kwargs)
        out.channels = filters
        return out
    return _conv2d_array(inputs, filters, kernel_size, strides, padding, name=name, **kwargs)


def _batch_normalization(inputs, training=False, name=None, **kwargs):
    if isinstance(inputs, Node):
        c = inputs.channels
        for v, tr in (("gamma", True), ("beta", True), ("moving_mean", False), ("moving_variance", False)):
            _declare(name + "/" + v, (c,), tr)
        out = _lazy(_batch_normalization_array)(inputs, training=training, name=name, **kwargs)
        out.channels = c
        return out
    return _batch_normalization_array(inputs, training=training, name=name, **kwargs)


def _relu(x):
    out = _lazy(_relu_array)(x)
    if isinstance(x, Node):
        out.channels = x.channels
    return out


def _placeholder(shape=None, dtype=np.float32, name=None):
    p = Placeholder(list(shape), dtype, name)
    p.channels = shape[-1]
    return p


def install():
    """Registers the stand-in as `tensorflow` (and the `tensorflow.contrib.slim` the reference imports for its unused
    separable_conv) unless a real TensorFlow is importable.  Returns the module."""
    if "tensorflow" in sys.modules and not getattr(sys.modules["tensorflow"], "_rced_standin", False):
        raise RuntimeError("a real tensorflow is loaded: use it instead of the stand-in")
    tf = types.ModuleType("tensorflow")
    tf._rced_standin = True
    tf.layers = types.SimpleNamespace(conv2d=_conv2d, batch_normalization=_batch_normalization)
    tf.nn = types.SimpleNamespace(relu=_relu)
    tf.float32 = np.float32
    tf.placeholder = _placeholder
    tf.Session = Session
    tf.GPUOptions = lambda **kw: kw
    tf.ConfigProto = lambda **kw: kw
    tf.train = types.SimpleNamespace(Saver=Saver)
    tf.global_variables = lambda: [v for v, _ in _STORE["created"]]
    tf.trainable_variables = lambda: [v for v, tr in _STORE["created"] if tr]
    tf.shape = _Shape
    tf.reset_default_graph = reset_graph
    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")
    contrib.slim = slim
    tf.contrib = contrib
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.slim"] = slim
    return tf
