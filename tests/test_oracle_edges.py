"""Pins the round-2 oracle functions to outputs of the UNMODIFIED reference (tests/golden/edges_golden.npz, minted by
oracle/make_golden_edges.py): predict_zero_shot of both variants called through their own `models`-dict convention,
calculate_multilabel_metrics and the prompt-mean pooling.  CPU only."""
import os

import numpy as np
import pytest
import torch

import ref_head as R
import stubs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def edges():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "edges_golden.npz")))


def _as_idx(names):
    return [[stubs.DISEASES.index(n) for n in row] for row in names]


def test_zero_shot_topk_lists_match_reference(edges):
    feats, text = torch.from_numpy(edges["feats"]), torch.from_numpy(edges["text"])
    names, scores = R.zero_shot_lists_topk(feats, text, stubs.DISEASES, 3)
    assert np.array_equal(np.array(_as_idx(names)), edges["z1_idx"])
    np.testing.assert_allclose(np.array(scores, dtype=np.float32), edges["z1_val"], rtol=2e-6, atol=1e-8)
    i1, v1 = R.zero_shot_softmax_topk(feats[5:6], text, 3)
    assert np.array_equal(i1[0].numpy(), edges["z1_single_idx"])
    np.testing.assert_allclose(v1[0].numpy(), edges["z1_single_val"], rtol=2e-6)


@pytest.mark.parametrize("tag", sorted(stubs.z2_cases()))
def test_zero_shot_multimodal_lists_match_reference(edges, tag):
    feats, text = torch.from_numpy(edges["feats"]), torch.from_numpy(edges["text"])
    names, scores = R.zero_shot_lists_multimodal(feats, text, stubs.DISEASES, **stubs.z2_cases()[tag])
    ref_idx, ref_val = stubs.unpad_lists(edges[f"z2_{tag}_idx"], edges[f"z2_{tag}_val"])
    assert _as_idx(names) == ref_idx
    for a, b in zip(scores, ref_val):
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-8)


def test_zero_shot_multimodal_with_attention_and_single(edges):
    text = torch.from_numpy(edges["text"])
    names, scores = R.zero_shot_lists_multimodal(torch.from_numpy(edges["feats_attn"]), text, stubs.DISEASES, threshold=0.5, top_k=2)
    ref_idx, ref_val = stubs.unpad_lists(edges["z2_attn_idx"], edges["z2_attn_val"])
    assert _as_idx(names) == ref_idx
    for a, b in zip(scores, ref_val):
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-8)
    n1, s1 = R.zero_shot_lists_multimodal(torch.from_numpy(edges["feats"][7:8]), text, stubs.DISEASES, 0.5, 2)
    assert _as_idx(n1)[0] == list(edges["z2_single_idx"])
    np.testing.assert_allclose(s1[0], edges["z2_single_val"], rtol=2e-6)


@pytest.mark.parametrize("tag,pseed,lseed,n,dens", [("a", 411, 412, 200, 0.2), ("b", 413, 414, 37, 0.0524)])
def test_multilabel_metrics_match_reference(edges, tag, pseed, lseed, n, dens):
    import synth
    pred = torch.sigmoid(synth.randn(pseed, n, 16) * 2.0)
    lab = synth.labels(lseed, n, 16, density=dens)
    m = R.calculate_multilabel_metrics(pred, lab)
    got = np.array([m[k] for k in ("sample_acc", "label_acc", "hamming_score", "exact_match", "top1_acc", "top3_acc", "f1_score")])
    np.testing.assert_allclose(got, edges[f"metrics_{tag}"], rtol=0, atol=0)
    _, cls = R.multilabel_batch_metrics((pred > 0.5).float(), lab)
    np.testing.assert_allclose((cls * 100).numpy(), edges[f"metrics_{tag}_class_acc"], rtol=0, atol=0)


def test_prompt_mean_pool_matches_reference(edges):
    out = R.prompt_mean_pool(torch.from_numpy(edges["pool_in"]), edges["pool_counts"].tolist())
    np.testing.assert_allclose(out.numpy(), edges["pool_out"], rtol=0, atol=0)
    assert len(set(edges["pool_counts"].tolist())) >= 3          # ragged prompt counts are covered
