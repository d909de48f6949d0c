"""R-CED V1 / V2 and CR-CED V3 with the reference's constructor and call signature
(model_utils/model.py:6-96 of the reference): ``Model(is_training)(x[N,T,129,1]) -> [N,T,129,1]``.

There is no TensorFlow graph: a model object holds its (TensorFlow-named) variables and runs the
fused sm_100a network kernel through the C ABI.  The layer tables live in fold.py (host) and
csrc/rced_arch.cuh (device).  Training mode is not part of the enhancement path."""
import numpy as np
import torch

from .. import runtime
from . import ckpt, fold


class _FusedModel(object):
    net_work = "FullyCNN"

    def __init__(self, is_training, device=None):
        if is_training:
            raise NotImplementedError("only the inference graph (is_training=False) is implemented; training is "
                                      "outside the enhancement forward path")
        self.is_training = False
        self.device_index = runtime.default_device() if device is None else int(device)
        self.weights = None
        self._eng = None

    # -- weights -------------------------------------------------------------------------------
    def set_weights(self, weights):
        missing = [n for n in fold.variable_names(self.net_work) if n not in weights]
        if missing:
            raise KeyError("missing variables: " + ", ".join(missing))
        self.weights = {n: np.asarray(weights[n], dtype=np.float32) for n in fold.variable_names(self.net_work)}
        if self._eng is not None:
            self._eng.close()
            self._eng = None
        return self

    def restore(self, checkpoint_path):
        """Checkpoint prefix (``Saver.restore``) or frozen ``.pb`` (``freeze.py``)."""
        return self.set_weights(ckpt.load_weights(checkpoint_path, self.net_work))

    def initialize(self, seed=0):
        """Random initialisation as TensorFlow would create the graph's variables."""
        return self.set_weights(fold.glorot_weights(self.net_work, seed))

    def engine(self):
        if self._eng is None:
            if self.weights is None:
                raise RuntimeError("model has no weights: call restore(checkpoint), set_weights(dict) or initialize()")
            from ..engine import Enhancer
            self._eng = Enhancer(self.net_work, self.weights, device=self.device_index)
        return self._eng

    def param_count(self):
        return fold.trainable_parameter_count(self.weights, self.net_work)

    # -- forward -------------------------------------------------------------------------------
    def __call__(self, x):
        """x: [N, T, 129, 1] magnitudes -- a numpy array (any float dtype; cast to float32 like a
        TensorFlow feed) or a CUDA float32 tensor.  Returns the same kind, float32."""
        eng = self.engine()
        is_np = not torch.is_tensor(x)
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))) if is_np else x
        if t.dim() != 4 or t.shape[2] != 129 or t.shape[3] != 1:
            raise ValueError("expected input of shape [N, T, 129, 1], got %s" % (tuple(t.shape),))
        n, T = int(t.shape[0]), int(t.shape[1])
        d = t.to(device=eng.device, dtype=torch.float32).contiguous().view(n * T, 129)
        row_off = torch.arange(n + 1, dtype=torch.int64, device=eng.device) * T
        pred = eng.forward_device(d, row_off).view(n, T, 129, 1)
        if is_np:
            torch.cuda.synchronize(eng.device)
            return pred.cpu().numpy()
        return pred


class FullyCNNSEModel(_FusedModel):
    """R-CED, 10 layers, 32,765 parameters (model.py:6-29)."""
    net_work = "FullyCNN"


class FullyCNNSEModelV2(_FusedModel):
    """R-CED, 16 layers, 32,192 parameters (model.py:32-61)."""
    net_work = "FullyCNNV2"


class FullyCNNSEModelV3(_FusedModel):
    """CR-CED with cascaded skips, 16 layers, 32,653 parameters (model.py:64-96)."""
    net_work = "FullyCNNV3"


def build_model(net_work, is_training=False, device=None):
    """Model selection of tester.py:76-82 / infer.py:45-51 / freeze.py:22-27."""
    if net_work == "FullyCNNV2":
        return FullyCNNSEModelV2(is_training, device)
    if net_work == "FullyCNNV3":
        return FullyCNNSEModelV3(is_training, device)
    print("net_work set default or not wright. Use FullyCNN")
    return FullyCNNSEModel(is_training, device)
