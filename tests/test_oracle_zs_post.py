"""CPU checks of oracle/ref_zs_post.py (zero-shot post-processing): its F1 equals scikit-learn's, the threshold
search equals a direct numpy + sklearn transcription of zero_shot_predict.py:112-159, and the two-view merge follows the
documented cases of :183-213."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import ref_zs_post as Z  # noqa: E402

f1_score = pytest.importorskip("sklearn.metrics").f1_score


def test_f1_equals_sklearn():
    rng = np.random.default_rng(0)
    for n in (1, 7, 200):
        for p in (0.0, 0.1, 0.5, 1.0):
            y = (rng.random(n) < p).astype(int)
            z = (rng.random(n) < 0.4).astype(int)
            assert abs(Z.f1_binary(y, z) - f1_score(y, z, zero_division=0)) < 1e-12


def _thresholds_with_sklearn(scores, labels):
    out = {}
    for j in range(scores.shape[1]):
        s, y = scores[:, j], labels[:, j]
        pos, neg = s[y == 1], s[y == 0]
        if len(pos) == 0:
            out[j] = 0.8
            continue
        if len(neg) == 0:
            out[j] = 0.2
            continue
        best_f1, best = 0, 0.5
        lo, hi = max(0.1, np.mean(neg) - np.std(neg)), min(0.9, np.mean(pos) + np.std(pos))
        for thr in np.linspace(lo, hi, 20):
            f1 = f1_score(y, (s >= thr).astype(int), zero_division=0)
            if f1 > best_f1:
                best_f1, best = f1, thr
        out[j] = best
    return np.array([out[j] for j in range(scores.shape[1])])


def test_dynamic_thresholds_match_sklearn_transcription():
    rng = np.random.default_rng(1)
    N, L = 300, 14
    labels = (rng.random((N, L)) < 0.08).astype(int)
    labels[:, 3] = 0                                   # a label without positives -> 0.8
    labels[:, 5] = 1                                   # a label without negatives -> 0.2
    scores = 1 / (1 + np.exp(-(rng.normal(size=(N, L)) + 1.5 * labels)))
    got = Z.dynamic_thresholds(scores, labels)
    np.testing.assert_allclose(got, _thresholds_with_sklearn(scores, labels), rtol=0, atol=1e-12)
    assert got[3] == 0.8 and got[5] == 0.2
    assert np.all(Z.dynamic_thresholds(np.zeros((0, L)), np.zeros((0, L))) == 0.3)


def test_view_predictions_and_merge_cases():
    prob = np.array([0.875, 0.25, 0.5625, 0.125])                                # exactly representable in float32
    assert Z.view_predictions(prob, 0.5) == ([0, 2], [0.875, 0.5625])
    assert Z.view_predictions(prob, 0.95) == ([0], [0.875])                     # nothing passes -> top-1
    assert Z.view_predictions(prob, {0: 0.95, 2: 0.5}) == ([2], [0.5625])       # labels missing from the dict never pass
    assert Z.view_predictions(prob, 0.5, top_k=3) == ([0, 2, 1], [0.875, 0.5625, 0.25])   # padded from the top-k list
    assert Z.view_predictions(prob, 0.05, top_k=2) == ([0, 2], [0.875, 0.5625])  # cut to the best k
    # the per-view test happens in float32 (torch keeps the tensor dtype): a threshold a hair above the float32 value still passes
    p32 = float(np.float32(0.3))
    assert Z.view_predictions(np.array([p32]), p32 + 1e-10) == ([0], [p32])
    thr = np.array([0.6, 0.6, 0.6, 0.6])
    # lateral view counts 0.8: 0.7 * 0.8 = 0.56 < 0.6 is dropped, the frontal 0.65 stays
    assert Z.merge_two_views([[0], [1]], [[0.65], [0.7]], thr) == ([0], [0.65])
    # same label in both views: weighted maximum
    p, s = Z.merge_two_views([[2], [2]], [[0.5], [0.9]], thr)
    assert p == [2] and abs(s[0] - 0.72) < 1e-12
    # nothing passes: the single best weighted score, first in insertion order on ties
    assert Z.merge_two_views([[0, 1], [3]], [[0.3, 0.3], [0.2]], thr) == ([0], [0.3])


def test_merged_prediction_matrix_shape_and_fallback():
    rng = np.random.default_rng(2)
    pv = rng.random((50, 2, 14))
    thr = np.full(14, 0.97)
    m = Z.merged_prediction_matrix(pv, thr)
    assert m.shape == (50, 14) and np.all(m.sum(1) >= 1)                        # every sample gets at least one label


def _main_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "zs_main_golden.npz"))


def test_oracle_equals_reference_main_run():
    """The reference's own main() (multimodal_attention/zero_shot_predict.py:14-261) was executed unmodified by
    oracle/make_golden_zs_main.py with the data side stubbed; its thresholds dict and final prediction matrix are the golden."""
    g = _main_golden()
    pv, labels, n_thr = g["prob_views"], g["labels"], int(g["n_threshold_studies"])
    thr = Z.dynamic_thresholds(pv[:n_thr].astype(np.float64).max(axis=1), labels[:n_thr])
    assert np.array_equal(thr, g["thresholds"])                                  # bit-exact float64
    assert thr[3] == 0.8 and thr[9] == 0.2
    assert np.array_equal(Z.merged_prediction_matrix(pv, g["thresholds"]), g["pred_matrix"])
    assert g["pred_matrix"].sum() > 0 and (g["pred_matrix"].sum(axis=1) >= 1).all()   # the fallback keeps one label per study
