"""MultiModalAttention (multimodal_attention/train.py:1069-1110) vs the oracle: golden output of the
unmodified reference module, and autograd on the oracle for every gradient (inputs, all four Linear layers), including a
gradient flowing into the returned attention weights."""
import numpy as np
import pytest
import torch

from gpu_util import dev, gpu, rel_l2
import ref_head as R
import synth

pytestmark = gpu


def _params():
    return {"wi": synth.uniform(81, -0.04, 0.04, 512, 512), "bi": synth.uniform(82, -0.04, 0.04, 512),
            "wt": synth.uniform(83, -0.04, 0.04, 512, 512), "bt": synth.uniform(84, -0.04, 0.04, 512),
            "wa": synth.uniform(85, -0.04, 0.04, 1, 512), "ba": synth.uniform(86, -0.04, 0.04, 1),
            "wo": synth.uniform(87, -0.04, 0.04, 512, 512), "bo": synth.uniform(88, -0.04, 0.04, 512)}


def _module(ap, d):
    import b200clip
    m = b200clip.MultiModalAttention().to(d)
    m.load_state_dict({"image_proj.weight": ap["wi"], "image_proj.bias": ap["bi"], "text_proj.weight": ap["wt"],
                       "text_proj.bias": ap["bt"], "attention.weight": ap["wa"], "attention.bias": ap["ba"],
                       "output_proj.weight": ap["wo"], "output_proj.bias": ap["bo"]})
    return m


def test_attention_matches_reference_golden(golden):
    d = dev()
    m = _module(_params(), d)
    enh, w = m(synth.randn(89, 6, 512).to(d), synth.unit_rows(90, 16, 512).to(d))
    assert rel_l2(enh, torch.from_numpy(golden["attn_enh"])) < 6e-3            # bf16 GEMM operands
    assert rel_l2(w, torch.from_numpy(golden["attn_w"])) < 2e-3
    np.testing.assert_allclose(w.detach().sum(1).cpu().numpy(), np.ones(6), atol=1e-5)


@pytest.mark.parametrize("B,C", [(64, 16), (1000, 14), (4096, 16), (1001, 3), (1, 16), (20001, 14)])   # odd B: the row-pair tail
def test_attention_forward_backward(B, C):
    d = dev()
    rnd = synth.bf16_round
    ap = _params()
    apr = {k: (rnd(v) if k in ("wi", "wo") else v.clone()).requires_grad_(True) for k, v in ap.items()}
    x = rnd(synth.randn(1, B, 512))
    t = synth.unit_rows(2, C, 512) * 3.0
    xr, tr = x.clone().requires_grad_(True), t.clone().requires_grad_(True)
    enh_ref, w_ref = R.multimodal_attention(xr, tr, apr)
    g = synth.randn(3, B, 512)
    gw = synth.randn(4, B, C)
    ((enh_ref * g).sum() + (w_ref * gw).sum()).backward()
    m = _module({k: v.detach() for k, v in apr.items()}, d)
    xg, tg = x.to(d).requires_grad_(True), t.to(d).requires_grad_(True)
    enh, w = m(xg, tg)
    ((enh * g.to(d)).sum() + (w * gw.to(d)).sum()).backward()
    assert rel_l2(enh, enh_ref) < 6e-3
    assert rel_l2(w, w_ref) < 3e-3
    assert rel_l2(xg.grad, xr.grad) < 2e-2
    assert rel_l2(tg.grad, tr.grad) < 2e-2
    for name, key in (("image_proj.weight", "wi"), ("image_proj.bias", "bi"), ("text_proj.weight", "wt"), ("text_proj.bias", "bt"),
                      ("attention.weight", "wa"), ("output_proj.weight", "wo"), ("output_proj.bias", "bo")):
        got = dict(m.named_parameters())[name].grad
        assert rel_l2(got, apr[key].grad) < 2e-2, name
    # attention.bias: a constant added to every score leaves the softmax unchanged -> zero gradient
    assert float(m.attention.bias.grad.abs().max()) == 0.0 and float(apr["ba"].grad.abs().max()) < 1e-4


def test_attention_rejects_unsupported_shapes():
    import b200clip
    d = dev()
    m = b200clip.MultiModalAttention().to(d)
    with pytest.raises(RuntimeError):
        m(torch.zeros(8, 512, device=d), torch.zeros(17, 512, device=d))          # C > 16
    with pytest.raises(RuntimeError):
        m(torch.zeros(8, 512), torch.zeros(16, 512))                              # CPU tensors
