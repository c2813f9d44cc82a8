"""Zero-shot post-processing on the device vs oracle/ref_zs_post.py: dynamic thresholds identical to the
numpy/sklearn search, merged prediction matrix bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

from gpu_util import dev, gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import ref_zs_post as Z  # noqa: E402

pytestmark = gpu


def _scores(seed, N, L, density):
    rng = np.random.default_rng(seed)
    labels = (rng.random((N, L)) < density).astype(np.float32)
    s = (1 / (1 + np.exp(-(rng.normal(size=(N, L)) + 1.5 * labels)))).astype(np.float32)
    return s, labels


@pytest.mark.parametrize("N,L,density", [(300, 14, 0.08), (5000, 16, 0.05), (40, 5, 0.3), (20000, 32, 0.02)])
def test_dynamic_thresholds(N, L, density):
    import b200clip
    d = dev()
    s, y = _scores(1, N, L, density)
    if L > 4:
        y[:, 3] = 0                                    # no positives -> 0.8
        y[:, 4] = 1                                    # no negatives -> 0.2
    ref = Z.dynamic_thresholds(s, y)
    thr, f1 = b200clip.dynamic_thresholds(torch.from_numpy(s).to(d), torch.from_numpy(y).to(d), return_f1=True)
    np.testing.assert_allclose(thr.cpu().numpy(), ref, rtol=0, atol=1e-9)
    assert bool((f1 >= 0).all()) and bool((f1 <= 1).all())
    if L > 4:
        assert float(thr[3]) == 0.8 and float(thr[4]) == 0.2


@pytest.mark.parametrize("N,L", [(1, 14), (257, 14), (5000, 16), (1000, 32)])
def test_merge_two_views_bit_identical(N, L):
    import b200clip
    d = dev()
    rng = np.random.default_rng(7)
    pv = rng.random((N, 2, L)).astype(np.float32)
    pv[: N // 4] *= 0.3                                # a block of samples where nothing passes: fallback path
    if N > 8:
        pv[5, 1] = pv[5, 0]                            # identical views: weighted maximum keeps the frontal score
        pv[6, :, :] = 0.25                             # all ties: first inserted label wins the fallback
    thr = (0.35 + 0.4 * rng.random(L))
    thr[0] = float(np.float32(pv[0, 0, 0]))           # a threshold exactly on a score (>= must pass)
    ref = Z.merged_prediction_matrix(pv, thr)
    got = b200clip.merge_two_views(torch.from_numpy(pv).to(d), torch.from_numpy(thr).to(d))
    assert np.array_equal(got.cpu().numpy().astype(np.float64), ref)


def test_thresholds_then_merge_pipeline_matches_oracle():
    import b200clip
    d = dev()
    N, L = 2000, 14
    rng = np.random.default_rng(3)
    labels = (rng.random((N, L)) < 0.08).astype(np.float32)
    pv = (1 / (1 + np.exp(-(rng.normal(size=(N, 2, L)) + 1.2 * labels[:, None, :])))).astype(np.float32)
    mx = pv.max(1)
    thr_ref = Z.dynamic_thresholds(mx, labels)
    thr = b200clip.dynamic_thresholds(torch.from_numpy(mx).to(d), torch.from_numpy(labels).to(d))
    np.testing.assert_allclose(thr.cpu().numpy(), thr_ref, rtol=0, atol=1e-9)
    got = b200clip.merge_two_views(torch.from_numpy(pv).to(d), thr)
    ref = Z.merged_prediction_matrix(pv, thr.cpu().numpy())
    assert np.array_equal(got.cpu().numpy().astype(np.float64), ref)


def test_device_path_equals_reference_main_run():
    """Golden minted by the reference's unmodified main() (oracle/make_golden_zs_main.py): thresholds from the first 25 % of
    the validation studies, merged prediction matrix over all of them."""
    import b200clip
    d = dev()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "zs_main_golden.npz"))
    pv, labels, n_thr = g["prob_views"], g["labels"], int(g["n_threshold_studies"])
    mx = torch.from_numpy(pv[:n_thr]).to(d).amax(dim=1)
    thr = b200clip.dynamic_thresholds(mx, torch.from_numpy(labels[:n_thr]).to(d))
    np.testing.assert_allclose(thr.cpu().numpy(), g["thresholds"], rtol=0, atol=1e-12)
    got = b200clip.merge_two_views(torch.from_numpy(pv).to(d), thr)
    assert np.array_equal(got.cpu().numpy().astype(np.float64), g["pred_matrix"])
    got = b200clip.merge_two_views(torch.from_numpy(pv).to(d), torch.from_numpy(g["thresholds"]).to(d))
    assert np.array_equal(got.cpu().numpy().astype(np.float64), g["pred_matrix"])
