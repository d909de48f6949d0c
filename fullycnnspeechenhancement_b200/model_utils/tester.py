"""BaseTester / FullyCNNTester with the reference's interface (model_utils/tester.py of the
reference).  ``creat_graph`` builds the fused-kernel model instead of a TensorFlow graph,
``_load_checkpoint`` reads the same checkpoint prefix (or a freeze.py ``.pb``) without
TensorFlow, ``test_step`` runs the network kernel, ``test`` the whole batch loop."""
import os
import time

import numpy as np

from .. import audio_io
from ..config import section_with
from . import fold
from .model import build_model
from .utils import PESQ, SDR, STOI, AudioReBuild, AverageMeter, sdr_batch


class BaseTester(object):
    def __init__(self, test_config):
        sec = section_with(test_config, "checkpoint_filepath")
        self.checkpoint_file = test_config.get(sec, "checkpoint_filepath")
        self.net_arch = test_config.get("model", "net_arch")
        self.net_work = test_config.get("model", "net_work")
        self.model = None

    def creat_graph(self):
        """tester.py:67-83 / infer.py:36-52: pick the model by ``net_work``.  The placeholder
        shape [None, None, feature_dim, 1] becomes a runtime shape check in Model.__call__."""
        self.model = build_model(self.net_work, is_training=False)
        self.pred = self.model

    def _init_session(self):
        """Nothing to configure up front: the CUDA handle is created when weights arrive."""
        self.sess = None

    def _load_checkpoint(self):
        self.model.restore(self.checkpoint_file)
        self.model.engine()
        print("recover from checkpoint_file: {}".format(self.checkpoint_file))

    def param_count(self):
        w = self.model.weights
        total = 0
        for scope, bn in fold.layer_scopes(self.model.net_work):
            names = [scope + "/kernel", scope + "/bias"]
            if bn:
                names += [scope + "/batch_norm/gamma", scope + "/batch_norm/beta"]
            for n in names:
                print("{}:0 layer parameter numbers | {}".format(n, w[n].size))
                total += w[n].size
        print("\nTotal number of Parameters: {}\n".format(total))
        return total

    def test_step(self, input_x):
        """[N,T,129,1] magnitudes -> [N,T,129,1] float32 (tester.py:85-90)."""
        return self.model(input_x)


class FullyCNNTester(BaseTester):
    def __init__(self, test_config):
        super(FullyCNNTester, self).__init__(test_config)
        self.sample_rate = int(test_config.get("data", "sample_rate"))
        self.feature_dim = int(test_config.get("data", "feature_dim"))
        self.audio_save_path = test_config.get("data", "audio_save_path")
        self.batch_size = int(test_config.get("testing", "batch_size"))
        if self.feature_dim != 129:
            raise NotImplementedError("feature_dim must be 129 (nfft 256), the only value the reference's STFT produces")
        self.creat_graph()
        self._init_session()
        self._load_checkpoint()
        self.param_count()
        self.pesq_score = AverageMeter()
        self.stoi_score = AverageMeter()
        self.sdr_score = AverageMeter()
        if not os.path.exists(self.audio_save_path):
            os.makedirs(self.audio_save_path)

    def test(self, valid_loader):
        """The batch loop of tester.py:92-167.  Enhancement runs waveform -> waveform on the GPU
        (K1 -> K2 -> K3 from ``mix_sig``; the loader's complex batch is the same STFT and is not
        needed).  SDR (GPU energy sums) and STOI (numpy restatement of pystoi) are scored like the
        reference does (tester.py:130-146); PESQ needs the ITU-T P.862 C code, which is not available:
        it is printed as n/a, never as a number."""
        from concurrent.futures import ThreadPoolExecutor
        Pesq = PESQ(sr=self.sample_rate)
        Stoi = STOI(sr=self.sample_rate)
        eng = self.model.engine()
        # the reference scores and writes the utterances of a batch in parallel (joblib, tester.py:128-160); here a thread
        # pool does (numpy's FFTs and the wav writer release the interpreter lock)
        pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
        for index, (batch_mix, batch_clean, mix_sig, clean_sig) in enumerate(valid_loader):
            start = time.time()
            audio_bins = valid_loader.bins[index]
            # the reference rebuilds (T+1)*128 samples per utterance and truncates them to len(clean_sig[i])
            # (tester.py:107-113, utils.py:181-182)
            denoise = eng.enhance(mix_sig, out_lens=[len(c) for c in clean_sig])
            denoise = [np.asarray(d, dtype=np.float64) for d in denoise]
            lens = [min(len(c), len(d)) for c, d in zip(clean_sig, denoise)]
            scores = sdr_batch([np.asarray(c[:n]) for c, n in zip(clean_sig, lens)], [d[:n] for d, n in zip(denoise, lens)],
                               device=eng.device.index)
            def finish(i):
                n = lens[i]
                st = float(Stoi(np.asarray(clean_sig[i][:n], dtype=np.float64), denoise[i][:n]))
                pq = float(Pesq(np.asarray(clean_sig[i][:n]), denoise[i][:n])) if Pesq.available else None
                name = os.path.basename(valid_loader.dataset.item_name(audio_bins[i]))
                audio_io.write_wav(os.path.join(self.audio_save_path, name), clean_sig[i], self.sample_rate)
                audio_io.write_wav(os.path.join(self.audio_save_path, name.replace(".wav", "_mix.wav")), mix_sig[i], self.sample_rate)
                audio_io.write_wav(os.path.join(self.audio_save_path, name.replace(".wav", "_de.wav")), denoise[i], self.sample_rate)
                return st, pq

            for i, (st, pq) in enumerate(pool.map(finish, range(len(audio_bins)))):
                self.sdr_score.update(float(scores[i]))
                self.stoi_score.update(st)
                if pq is not None:
                    self.pesq_score.update(pq)
            print("Testing %d  STOI=%.4f  SDR=%.4f  BatchTime=%.3f" % (index, self.stoi_score.avg, self.sdr_score.avg, time.time() - start))
        pool.shutdown()
        p_score = "{:.4f}".format(self.pesq_score.avg) if Pesq.available else "n/a (pypesq / ITU-T P.862 not available)"
        print("Average p_score: {}; Average st_score: {:.4f}; Average sd_score: {:.4f}.\n".format(
            p_score, self.stoi_score.avg, self.sdr_score.avg))
        return self.sdr_score.avg
