"""CPU oracle for the enhancement forward path of phecda-xu/FullyCNNSpeechEnhancement.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``fullycnnspeechenhancement_b200/`` imports it,
and the product path fails loudly when the CUDA library is missing.

Parity pinning status:

* ``oracle.stft`` and ``oracle.rebuild`` restate numpy code that *does* run in the
  authoring container.  They are pinned bit-for-bit against the unmodified
  reference (``/root/reference/data_utils/audio_feature.py`` and
  ``/root/reference/model_utils/utils.py:AudioReBuild``) by
  ``tests/golden/make_golden.py``; the resulting vectors are committed under
  ``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.
* ``oracle.network`` restates TensorFlow-1.14 graph semantics
  (``model_utils/module.py:11-34``, ``model_utils/model.py:6-96``).  TensorFlow
  1.14 (``requriements.txt:4``) is not installable here and the reference ships no
  tests, checkpoints or golden vectors, so the arithmetic of TensorFlow's operations
  is **parity unpinned**: it is anchored on (a) the published parameter counts 32,765 / 32,192 / 32,653
  (``readme.md:63-67``), (b) agreement between two independently written
  evaluators (a tap-loop NHWC SAME convolution in numpy float64 and
  ``torch.nn.functional.conv2d`` with explicit asymmetric padding), and (c) the
  reference's OWN model classes (``model_utils/model.py``, imported unmodified by
  ``oracle/ref_import.load_models``) executed with ``oracle/tf_standin.py`` in place of
  TensorFlow: the WIRING -- layers, widths, kernel sizes, skip inputs, the position of
  the addition, the variable scopes -- is then the reference's source, run, not a
  restatement (``tests/golden/network_ref_model.npz``, agreement 2e-15).  What stays
  unpinned is the arithmetic of the three TensorFlow operations themselves (SAME
  padding of the even time kernel, batch-norm epsilon), stated from TensorFlow's
  documented behaviour in four independently written forms.  The same stand-in in graph
  mode runs the reference's ENTRY POINTS unmodified (``test.py main()``,
  ``infer.py InferenceEngine.denoise``; ``oracle/ref_import.load_test_entry``):
  ``tests/golden/reference_test_entry.npz``, ``tests/test_reference_entry.py``.
"""
