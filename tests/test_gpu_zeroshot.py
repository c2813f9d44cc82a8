"""Zero-shot scoring (zeroshot.cu) vs the oracle evaluated in fp64 on the same bf16-rounded operands: label sets and
argmax must be IDENTICAL (guard band + on-device fp64 re-evaluation)."""
import pytest
import torch

from gpu_util import dev, gpu
import ref_head as R
import synth

pytestmark = gpu


def _data(N, NP, seed=51, D=512):
    X = synth.bf16_round(synth.randn(seed, N, D) * 3.0)
    P = synth.bf16_round(synth.unit_rows(seed + 1, NP, D))
    return X, P


@pytest.mark.parametrize("N", [1, 15, 16, 200, 5000, 100003])
def test_posneg_14x2_exact(N):
    import b200clip
    X, P = _data(N, 28)
    am_ref, mask_ref, _ = R.zero_shot_posneg(X.double(), P.double().reshape(14, 2, 512), 0.07, 0.5)
    am, mask = b200clip.zero_shot_posneg(X.to(dev()).to(torch.bfloat16), P.to(dev()).to(torch.bfloat16).reshape(14, 2, 512))
    assert torch.equal(am.cpu().long(), am_ref)
    assert torch.equal(b200clip.unpack_mask(mask, 14).cpu(), mask_ref)


def test_posneg_threshold_other_than_half():
    import b200clip
    X, P = _data(3000, 28)
    am_ref, mask_ref, _ = R.zero_shot_posneg(X.double(), P.double().reshape(14, 2, 512), 0.07, 0.7)
    am, mask = b200clip.zero_shot_posneg(X.to(dev()).to(torch.bfloat16), P.to(dev()).to(torch.bfloat16).reshape(14, 2, 512),
                                         threshold=0.7)
    assert torch.equal(b200clip.unpack_mask(mask, 14).cpu(), mask_ref)
    assert torch.equal(am.cpu().long(), am_ref)


def test_softmax_topk_16():
    import b200clip
    X, P = _data(4000, 16)
    idx_ref, val_ref = R.zero_shot_softmax_topk(X.double(), P.double(), 3, 0.07)
    idx, val = b200clip.zero_shot_topk(X.to(dev()), P.to(dev()), 3, 0.07)
    assert torch.equal(idx.cpu().long(), idx_ref)
    assert torch.allclose(val.cpu().double(), val_ref, rtol=2e-4, atol=1e-6)


def test_sigmoid_threshold_16_scalar_and_per_label():
    import b200clip
    X, P = _data(4000, 16)
    for thr in (0.5, [0.45 + 0.01 * i for i in range(16)]):
        mask_ref, _, am_ref = R.zero_shot_sigmoid_threshold(X.double(), P.double(), thr, 0.5)
        mask, am = b200clip.zero_shot_threshold(X.to(dev()), P.to(dev()), thr, 0.5)
        assert torch.equal(b200clip.unpack_mask(mask, 16).cpu(), mask_ref)
        assert torch.equal(am.cpu().long(), am_ref)


def test_golden_small_case(golden):
    """the committed fixtures (D=64) embedded in D=512 by zero padding"""
    import numpy as np
    import b200clip
    X = torch.zeros(200, 512); X[:, :64] = synth.randn(51, 200, 64)
    T16 = torch.zeros(16, 512); T16[:, :64] = synth.unit_rows(52, 16, 64)
    Xb, Tb = synth.bf16_round(X), synth.bf16_round(T16)
    idx, _ = b200clip.zero_shot_topk(Xb.to(dev()), Tb.to(dev()), 3, 0.07)
    idx_ref, _ = R.zero_shot_softmax_topk(Xb.double(), Tb.double(), 3, 0.07)
    assert torch.equal(idx.cpu().long(), idx_ref)
    # against the fp32 reference outputs on UNROUNDED inputs, bf16 rounding may legitimately flip near-ties:
    agree = (idx.cpu().numpy()[:, 0] == golden["z1_idx"][:, 0]).mean()
    assert agree > 0.97


def test_guard_band_rows_are_rare_and_counted():
    from b200clip import ops
    X, P = _data(20000, 28)
    out = ops.zeroshot_score(X.to(dev()).to(torch.bfloat16), P.to(dev()).to(torch.bfloat16), pair_mode=True, temperature=0.07,
                             thresholds=[0.5], count_guard=True)
    frac = out["guard_rows"].item() / 20000
    assert 0 < frac < 0.1


@pytest.mark.parametrize("N", [1, 9, 4000, 100003])
def test_width_768_all_modes_exact(N):
    """shared_embedding_size = 768 (0426/config.py:30 is a knob; BASELINE.json configs[4] names 512 and 768): 14-stage ring of
    12 KB blocks, 24 k32 steps per row, 24 elements per lane in the exact re-evaluation."""
    import b200clip
    D = 768
    X, P = _data(N, 28, seed=71, D=D)
    am_ref, mask_ref, _ = R.zero_shot_posneg(X.double(), P.double().reshape(14, 2, D), 0.07, 0.5)
    am, mask = b200clip.zero_shot_posneg(X.to(dev()).to(torch.bfloat16), P.to(dev()).to(torch.bfloat16).reshape(14, 2, D))
    assert torch.equal(am.cpu().long(), am_ref)
    assert torch.equal(b200clip.unpack_mask(mask, 14).cpu(), mask_ref)
    X, P = _data(N, 16, seed=73, D=D)
    k = 3
    idx_ref, val_ref = R.zero_shot_softmax_topk(X.double(), P.double(), k, 0.07)
    idx, val = b200clip.zero_shot_topk(X.to(dev()), P.to(dev()), k, 0.07)
    assert torch.equal(idx.cpu().long(), idx_ref)
    assert torch.allclose(val.cpu().double(), val_ref, rtol=2e-4, atol=1e-6)
    mask_ref, _, am_ref = R.zero_shot_sigmoid_threshold(X.double(), P.double(), [0.45 + 0.01 * i for i in range(16)], 0.5)
    mask, am = b200clip.zero_shot_threshold(X.to(dev()), P.to(dev()), [0.45 + 0.01 * i for i in range(16)], 0.5)
    assert torch.equal(b200clip.unpack_mask(mask, 16).cpu(), mask_ref)
    assert torch.equal(am.cpu().long(), am_ref)


def test_rejects_other_widths():
    import b200clip
    X, P = _data(64, 28, D=640)
    with pytest.raises(RuntimeError):
        b200clip.zero_shot_posneg(X.to(dev()).to(torch.bfloat16), P.to(dev()).to(torch.bfloat16).reshape(14, 2, 640))


def test_posneg_at_benchmark_size_against_chunked_fp64_oracle():
    """BASELINE.json configs[3] at its full size: 1 000 000 embeddings x 14 (positive, negative) prompt pairs, D = 512.  The
    fp64 oracle is evaluated in 8 chunks of 125 000 rows on the host; arg-max and label sets must be identical for every row."""
    import b200clip
    d = dev()
    N = 1_000_000
    g = torch.Generator(device=d).manual_seed(5)
    X = (torch.randn(N, 512, device=d, generator=g) * 3.0).to(torch.bfloat16)
    P = synth.bf16_round(synth.unit_rows(52, 28, 512))
    am, mask = b200clip.zero_shot_posneg(X, P.to(d).to(torch.bfloat16).reshape(14, 2, 512))
    am, sets = am.cpu().long(), b200clip.unpack_mask(mask, 14).cpu()
    for s in range(0, N, 125_000):
        am_ref, mask_ref, _ = R.zero_shot_posneg(X[s:s + 125_000].cpu().double(), P.double().reshape(14, 2, 512), 0.07, 0.5)
        assert torch.equal(am[s:s + 125_000], am_ref)
        assert torch.equal(sets[s:s + 125_000], mask_ref)
