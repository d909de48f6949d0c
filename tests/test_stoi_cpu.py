"""CPU: the host-side STOI score (model_utils/stoi.py) that replaces pystoi in FullyCNNTester.test
(reference: model_utils/utils.py:48-62, model_utils/tester.py:136-140), against an independent loop
restatement (oracle/stoi_ref.py) and the measure's known properties."""
import numpy as np
import pytest

from fullycnnspeechenhancement_b200.model_utils import stoi as S
from fullycnnspeechenhancement_b200.model_utils.utils import PESQ, STOI
from fullycnnspeechenhancement_b200.synth import noisy_utterance
from oracle import stoi_ref


def _pair(seed, n, snr_db, sr=8000):
    mix, clean = noisy_utterance(seed, n, sr, snr_db=snr_db, return_clean=True)
    return np.asarray(clean, np.float64), np.asarray(mix, np.float64)


@pytest.mark.parametrize("sr", [8000, 10000, 16000])
def test_stoi_matches_the_loop_restatement(sr):
    clean, mix = _pair(3, 3 * sr, 5.0, sr)
    a = S.stoi(clean, mix, sr)
    b = stoi_ref.stoi(clean, mix, sr)
    assert abs(a - b) < 1e-9, (a, b)
    assert 0.0 < a < 1.0


def test_stoi_known_answers():
    clean, _ = _pair(5, 32000, 0.0)
    # identical signals, and any positive gain (the measure normalises each segment's energy): 1
    assert abs(S.stoi(clean, clean, 8000) - 1.0) < 1e-9
    assert abs(S.stoi(clean, 0.25 * clean, 8000) - 1.0) < 1e-9
    # a polarity flip leaves every band envelope unchanged
    assert abs(S.stoi(clean, -clean, 8000) - 1.0) < 1e-9
    # fewer than 30 frames after the silent-frame removal: pystoi's sentinel value
    assert S.stoi(clean[:2000], clean[:2000], 8000) == 1e-5
    with pytest.raises(Exception):
        S.stoi(clean, clean[:-1], 8000)


def test_stoi_decreases_with_the_noise_level():
    scores = []
    for snr in (20.0, 10.0, 0.0, -10.0):
        clean, mix = _pair(7, 32000, snr)
        scores.append(S.stoi(clean, mix, 8000))
    assert all(a > b for a, b in zip(scores, scores[1:])), scores
    # (the synthetic voices have slow, shallow envelopes: absolute scores are lower than for real speech)
    assert scores[0] - scores[-1] > 0.2, scores


def test_silent_frames_do_not_count():
    """Frames 40 dB below the loudest clean frame are removed from both signals before the comparison: garbage in the
    processed signal during the clean signal's silence does not change the score."""
    rng = np.random.default_rng(0)
    clean, mix = _pair(9, 24000, 5.0)
    gap = np.zeros(8000)
    x = np.concatenate([clean, gap, clean])
    y1 = np.concatenate([mix, gap, mix])
    y2 = np.concatenate([mix, 0.5 * rng.normal(size=8000), mix])
    assert abs(S.stoi(x, y1, 8000) - S.stoi(x, y2, 8000)) < 0.02


def test_metric_classes_follow_the_reference_interface():
    clean, mix = _pair(11, 16000, 5.0)
    st = STOI(sr=8000)
    assert st.available and abs(st(clean, mix) - S.stoi(clean, mix, 8000)) == 0.0
    pq = PESQ(sr=8000)
    assert not pq.available and np.isnan(pq(clean, mix))     # reported as unavailable, never as a number
    with pytest.raises(AssertionError):
        st(clean, mix[:-1])
