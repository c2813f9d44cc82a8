"""MultiViewFusion (0426/train.py:988-1000) vs the oracle / golden fixture: forward, backward (autograd on
the oracle), train-mode dropout with the oracle fed the same mask."""
import numpy as np
import pytest
import torch

from gpu_util import dev, gpu, rel_l2
import ref_head as R
import synth

pytestmark = gpu


def _params():
    return {"w0": synth.uniform(61, -0.03, 0.03, 512, 1024), "b0": synth.uniform(62, -0.03, 0.03, 512),
            "w3": synth.uniform(63, -0.04, 0.04, 512, 512), "b3": synth.uniform(64, -0.04, 0.04, 512)}


def _module(fp, d, **kw):
    import b200clip
    m = b200clip.MultiViewFusion(**kw).to(d)
    m.load_state_dict({"fusion.0.weight": fp["w0"], "fusion.0.bias": fp["b0"], "fusion.3.weight": fp["w3"], "fusion.3.bias": fp["b3"]})
    return m


def test_fusion_matches_reference_golden(golden):
    """Same inputs as the fixture generated from the unmodified reference module (oracle/make_golden.py)."""
    d = dev()
    m = _module(_params(), d).eval()
    y = m(synth.randn(65, 6, 512).to(d), synth.randn(66, 6, 512).to(d))
    np.testing.assert_allclose(y.detach().cpu().numpy(), golden["fusion_y"], atol=2e-2, rtol=2e-2)      # bf16 operands
    assert rel_l2(y, torch.from_numpy(golden["fusion_y"])) < 6e-3


@pytest.mark.parametrize("B", [64, 1000, 4096])
def test_fusion_forward_backward(B):
    d = dev()
    fp = _params()
    rnd = synth.bf16_round
    f, l = rnd(synth.randn(1, B, 512)), rnd(synth.randn(2, B, 512))
    fpr = {"w0": rnd(fp["w0"]).requires_grad_(True), "b0": fp["b0"].clone().requires_grad_(True),
           "w3": rnd(fp["w3"]).requires_grad_(True), "b3": fp["b3"].clone().requires_grad_(True)}
    fr, lr = f.clone().requires_grad_(True), l.clone().requires_grad_(True)
    yref = R.multi_view_fusion(fr, lr, fpr)
    g = synth.randn(3, B, 512)
    yref.backward(g)
    m = _module({k: v.detach() for k, v in fpr.items()}, d).eval()
    fg, lg = f.to(d).requires_grad_(True), l.to(d).requires_grad_(True)
    y = m(fg, lg)
    y.backward(g.to(d))
    assert rel_l2(y, yref) < 5e-3
    assert rel_l2(fg.grad, fr.grad) < 2e-2 and rel_l2(lg.grad, lr.grad) < 2e-2
    assert rel_l2(m.fusion[0].weight.grad, fpr["w0"].grad) < 2e-2
    assert rel_l2(m.fusion[3].weight.grad, fpr["w3"].grad) < 2e-2
    assert rel_l2(m.fusion[0].bias.grad, fpr["b0"].grad) < 2e-2
    assert rel_l2(m.fusion[3].bias.grad, fpr["b3"].grad) < 2e-2


def test_fusion_train_mode_dropout_with_the_same_mask():
    """nn.Dropout(0.2) between ReLU and the second Linear (0426/train.py:994): torch's Philox stream cannot be reproduced in
    a fused epilogue, so the oracle is handed the mask the kernel used (ops.dropout_mask of the module's seed)."""
    from b200clip import ops
    d = dev()
    B = 512
    fp = _params()
    rnd = synth.bf16_round
    f, l = rnd(synth.randn(1, B, 512)), rnd(synth.randn(2, B, 512))
    m = _module({"w0": rnd(fp["w0"]), "b0": fp["b0"], "w3": rnd(fp["w3"]), "b3": fp["b3"]}, d).train()
    fg = f.to(d).requires_grad_(True)
    y = m(fg, l.to(d))
    mask = ops.dropout_mask(B, 512, 0.2, m.last_dropout_seed, d).cpu()
    keep = float((mask > 0).float().mean())
    assert 0.75 < keep < 0.85 and abs(float(mask.max()) - 1.25) < 1e-6
    fr = f.clone().requires_grad_(True)
    h = torch.relu(torch.cat([fr, l], dim=1) @ rnd(fp["w0"]).T + fp["b0"]) * mask
    yref = h @ rnd(fp["w3"]).T + fp["b3"]
    g = synth.randn(3, B, 512)
    yref.backward(g)
    y.backward(g.to(d))
    assert rel_l2(y, yref) < 5e-3
    assert rel_l2(fg.grad, fr.grad) < 2e-2
    m.eval()
    y2 = m(f.to(d), l.to(d))
    assert rel_l2(y2, R.multi_view_fusion(f, l, {"w0": rnd(fp["w0"]), "b0": fp["b0"], "w3": rnd(fp["w3"]), "b3": fp["b3"]})) < 5e-3


def test_fusion_rejects_cpu_tensors():
    import b200clip
    m = b200clip.MultiViewFusion()
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 512), torch.zeros(4, 512))
