"""Generate the committed golden fixtures (run in the AUTHORING container only).

    python tests/golden/make_golden.py

* ``stft_rebuild_ref.npz`` -- produced by the UNMODIFIED reference code imported
  from /root/reference (oracle/ref_import.py shims only np.mat and missing audio
  libraries): AudioFeature.compute_spectrogram / power_spectrum / divide_phase,
  DataLoader.padding_batch and AudioReBuild.rebuild_audio.  These pin oracle/stft.py
  and oracle/rebuild.py.
* ``frame_counts_ref.npz`` -- T(L) and frame start indices from the reference's
  en_frame for L in 1..1100 plus the sizes named in SURVEY.md section 8.
* ``network_oracle.npz`` -- outputs of oracle/network.py (float64) for seeded
  weights.  NOT produced by the reference (TensorFlow 1.14 is not installable):
  parity unpinned, see oracle/__init__.py.  Kept so the GPU box can check the CUDA
  path and the oracle against a value computed elsewhere.
* ``network_ref_model.npz`` -- outputs of the reference's OWN model classes (model_utils/model.py, imported unmodified)
  executed with oracle/tf_standin.py in place of TensorFlow, on the inputs and weights of ``network_oracle.npz``, plus the
  variables each model asks for in creation order.  Pins the wiring (layers, widths, kernel sizes, skip inputs, position of
  the addition, scopes / checkpoint names) to the reference's source; the arithmetic of conv2d / batch_normalization / relu
  is the stand-in's restatement of TensorFlow's documented behaviour.  `python tests/golden/make_golden.py network_ref_model`
  writes only this file.
* ``reference_test_entry.npz`` -- the reference's own evaluation entry point (test.py main(): DataSet, DataLoader,
  FullyCNNTester with graph creation, checkpoint restore, param_count, the batch loop with rebuild_audio and the scores),
  UNMODIFIED, run on the small job of tests/golden/ref_entry_case.py with the stand-ins of oracle/ref_import.load_test_entry
  (TensorFlow: oracle/tf_standin.py; librosa.load / soundfile.write / pystoi / joblib.Parallel: see there): the enhanced
  waveforms it hands to soundfile.write, its average SDR and STOI and the parameter total it prints, for V1 / V2 / V3;
  and what infer.py's InferenceEngine(cfg).denoise(file) writes for one of the files.
  `python tests/golden/make_golden.py reference_test_entry` writes only this file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import, network  # noqa: E402
from fullycnnspeechenhancement_b200.synth import noisy_utterance  # noqa: E402


def make_network_ref_model():
    from oracle import tf_standin
    models = ref_import.load_models()
    net = {}
    for arch in ("FullyCNN", "FullyCNNV2", "FullyCNNV3"):
        w = network.random_weights(arch, seed=1234, randomize_bn=True)      # the weights of network_oracle.npz
        rng = np.random.default_rng(77)
        for T in (1, 7, 8, 9, 12):
            x = np.abs(rng.normal(0, 3.0, (2, T, 129, 1))).astype(np.float32)   # the inputs of network_oracle.npz
            if T not in (1, 8, 12):
                continue
            tf_standin.set_variables(w)
            net["y_%s_%d" % (arch, T)] = models[arch](is_training=False)(x.astype(np.float64))
            net["xsum_%s_%d" % (arch, T)] = np.array([float(x.astype(np.float64).sum())])
        req = tf_standin.requested()
        net["names_" + arch] = np.array([n for n, _ in req])
        net["shapes_" + arch] = np.array(["x".join(str(d) for d in sh) for _, sh in req])
    np.savez_compressed(os.path.join(HERE, "network_ref_model.npz"), **net)


def make_reference_test_entry():
    import contextlib
    import io
    import tempfile
    sys.path.insert(0, HERE)
    import ref_entry_case as case
    from oracle import tf_standin
    entry, load_conf = ref_import.load_test_entry()

    RefTester = entry.FullyCNNTester

    class Recording(RefTester):
        last = None

        def __init__(self, cfg):
            RefTester.__init__(self, cfg)
            Recording.last = self
    entry.FullyCNNTester = Recording
    out = {}
    for arch in case.ARCHS:
        with tempfile.TemporaryDirectory() as d:
            prefix = os.path.join(d, "ckpt", "RCED_%s.ckpt" % arch)
            os.makedirs(os.path.dirname(prefix))
            w = network.random_weights(arch, seed=case.WEIGHT_SEED, randomize_bn=True)
            np.savez(prefix + ".standin.npz", **w)
            cfg, items = case.build(d, arch, prefix)
            ref_import.WRITTEN.clear()
            tf_standin.reset_graph()               # a fresh default graph per model, as a fresh process would have
            printed = io.StringIO()
            with contextlib.redirect_stdout(printed):
                entry.main(load_conf(cfg), 1)
            t = Recording.last
            for i, (pc, pm, L) in enumerate(items):
                de, rate = ref_import.WRITTEN[os.path.join(d, "out", "utt%d_de.wav" % i)]
                assert rate == 8000 and len(de) == L
                out["de_%s_%d" % (arch, i)] = np.asarray(de, np.float32)
                mix, _ = ref_import.WRITTEN[os.path.join(d, "out", "utt%d_mix.wav" % i)]
                out["mixsum_%s_%d" % (arch, i)] = np.array([float(np.abs(mix.astype(np.float64)).sum())])
            # infer.py: InferenceEngine(cfg).denoise(file) on the second noisy file (its reshape of the [F, T] arrays included)
            ref_import.WRITTEN.clear()
            tf_standin.reset_graph()
            with contextlib.redirect_stdout(printed):
                entry.infer.InferenceEngine(load_conf(cfg)).denoise(items[1][1])
            (path, (de, rate)), = ref_import.WRITTEN.items()
            assert path == os.path.join(d, "out", "utt1_noisy_de.wav") and rate == 8000 and len(de) == items[1][2]
            out["infer_de_" + arch] = np.asarray(de, np.float32)
            out["sdr_avg_" + arch] = np.array([t.sdr_score.avg])
            out["stoi_avg_" + arch] = np.array([t.stoi_score.avg])
            total = [ln for ln in printed.getvalue().splitlines() if ln.startswith("Total number of Parameters")][0]
            out["param_total_" + arch] = np.array([int(total.split(":")[1])])
            print(arch, "SDR %.4f STOI %.4f" % (t.sdr_score.avg, t.stoi_score.avg), total)
    np.savez_compressed(os.path.join(HERE, "reference_test_entry.npz"), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "reference_test_entry":
        make_reference_test_entry()
        print("reference_test_entry.npz", os.path.getsize(os.path.join(HERE, "reference_test_entry.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "network_ref_model":
        make_network_ref_model()
        print("network_ref_model.npz", os.path.getsize(os.path.join(HERE, "network_ref_model.npz")))
        return
    AudioFeature, AudioReBuild, DataLoader, AudioParser = ref_import.load()
    af = AudioFeature()

    # ---- frame counts / indices from the reference's en_frame -------------------------
    lengths = list(range(1, 1101)) + [16000, 24001, 31999, 32000, 32001, 64000]
    counts, first_start, last_start = [], [], []
    for L in lengths:
        _, frames = af.en_frame(0.032, 0.016, 8000, np.arange(1, L + 1, dtype=np.float64))
        frames = np.asarray(frames)
        counts.append(frames.shape[0])
        # frames hold sample_index+1 (0 where zero padded): recover the gather table exactly
        first_start.append(int(frames[0, 0]) - 1)
        last_start.append(int(frames[-1, 0]) - 1 if frames[-1, 0] > 0 else -1)
    np.savez_compressed(os.path.join(HERE, "frame_counts_ref.npz"),
                        lengths=np.array(lengths, np.int64), counts=np.array(counts, np.int64),
                        first_start=np.array(first_start, np.int64),
                        last_start=np.array(last_start, np.int64))

    # ---- STFT + rebuild through the reference ----------------------------------------
    out = {}
    cases = [(11, 100), (12, 256), (13, 257), (14, 1000), (15, 2049), (16, 4000)]
    out["case_seeds"] = np.array([c[0] for c in cases], np.int64)
    out["case_lengths"] = np.array([c[1] for c in cases], np.int64)
    specs = []
    for seed, L in cases:
        x = noisy_utterance(seed, L)
        out["wav_%d" % seed] = x
        X = np.asarray(af.compute_spectrogram(x, 8000, 0.032, 0.016, 256, True))   # [F,T] c128
        out["spec_%d" % seed] = X
        out["mag_%d" % seed] = af.power_spectrum(X)
        out["phase_%d" % seed] = af.divide_phase(X)
        specs.append(X)
    # batch layout (data_loader.py:198-209) on the three longest cases
    pb = DataLoader.padding_batch(None, specs[3:])
    out["padded_batch"] = pb                                       # [3,Tmax,129,1] c128
    mag = af.power_spectrum(pb)
    ph = af.divide_phase(pb)
    rng = np.random.default_rng(5)
    # a non-trivial "prediction": magnitude scaled per bin and offset, can go negative
    pred = (mag.squeeze(-1) * rng.uniform(0.2, 1.2, (1, 1, 129)) - 0.05).astype(np.float32)
    out["pred"] = pred
    lens = [c[1] for c in cases[3:]]
    for nfft in (512, 256):
        rb = AudioReBuild(nfft=nfft).rebuild_audio(lens, pred, ph.squeeze(-1), 8000, 32.0, 16.0)
        for i, r in enumerate(rb):
            out["rebuild%d_%d" % (nfft, i)] = np.asarray(r)
    np.savez_compressed(os.path.join(HERE, "stft_rebuild_ref.npz"), **out)

    # ---- network oracle outputs -------------------------------------------------------
    net = {}
    for arch in ("FullyCNN", "FullyCNNV2", "FullyCNNV3"):
        w = network.random_weights(arch, seed=1234, randomize_bn=True)
        net["wsum_" + arch] = np.array([float(np.sum([np.sum(v.astype(np.float64)) for v in w.values()])),
                                        float(np.sum([np.sum(np.abs(v.astype(np.float64))) for v in w.values()]))])
        rng = np.random.default_rng(77)
        for T in (1, 7, 8, 9, 12):
            x = np.abs(rng.normal(0, 3.0, (2, T, 129, 1))).astype(np.float32)
            net["x_%s_%d" % (arch, T)] = x
            net["y_%s_%d" % (arch, T)] = network.forward(arch, w, x, np.float64)
    np.savez_compressed(os.path.join(HERE, "network_oracle.npz"), **net)
    make_network_ref_model()
    make_reference_test_entry()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
