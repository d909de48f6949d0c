"""Checkpoint -> frozen GraphDef export with the reference's FreezeEngine interface
(freeze.py of the reference), written without TensorFlow by model_utils/ckpt.py."""
from .model_utils import ckpt, fold


class FreezeEngine(object):
    def __init__(self, net_work, feature_dim=129):
        self.net_work = net_work
        self.feature_dim = feature_dim

    def freeze_graph(self, checkpoint_file, pb_file):
        weights = ckpt.load_weights(checkpoint_file, self.net_work)
        n = ckpt.write_frozen_graph(pb_file, self.net_work, weights)
        print("%d ops in the final graph." % n)
        return fold.output_node_name(self.net_work)
