"""GraphedHeadStep (one CUDA graph per head step) must reproduce the eager step: same loss and gradients, and it must
pick up new inputs on every replay."""
import os
import sys

import torch

from gpu_util import dev, gpu, rel_l2

pytestmark = gpu


def _inputs(seed, B, E, D, C, device):
    import synth
    xi = synth.randn(seed, B, E).to(torch.bfloat16).to(device)
    xt = synth.randn(seed + 1, B, E).to(torch.bfloat16).to(device)
    return xi, xt, synth.unit_rows(3, C, D).to(device), synth.labels(seed + 2, B, C).to(device)


def test_graphed_step_equals_eager_step():
    import b200clip
    d = dev()
    B, E, D, C = 1024, 768, 512, 16
    torch.manual_seed(0)
    head = b200clip.ClipHead(E, E, D, C).to(d)
    xi, xt, ct, lab = _inputs(1, B, E, D, C, d)
    step = b200clip.GraphedHeadStep(head, xi, xt, ct, lab)
    for seed in (1, 7):                                    # second seed: the replay must see the NEW inputs
        xi, xt, ct, lab = _inputs(seed, B, E, D, C, d)
        loss_g = step(xi, xt, ct, lab)
        torch.cuda.synchronize()
        got = {"loss": float(loss_g), "dxi": step.grad_image.float().clone(), "dxt": step.grad_text.float().clone(),
               "gw": head.image_projector.fc.weight.grad.clone(), "gf": head.classifier.weight.grad.clone(),
               "gb": head.text_projector.layer_norm.bias.grad.clone()}
        for p in head.parameters():
            p.grad = None
        xe, te = xi.clone().requires_grad_(True), xt.clone().requires_grad_(True)
        loss_e = head(xe, te, ct, lab)
        loss_e.backward()
        torch.cuda.synchronize()
        assert abs(got["loss"] - float(loss_e)) <= 1e-6 * abs(float(loss_e))
        assert rel_l2(got["dxi"], xe.grad.float()) < 1e-5
        assert rel_l2(got["dxt"], te.grad.float()) < 1e-5
        # split-K weight gradients are accumulated with atomics: order differs run to run
        assert rel_l2(got["gw"], head.image_projector.fc.weight.grad) < 1e-4
        assert rel_l2(got["gf"], head.classifier.weight.grad) < 1e-5
        assert rel_l2(got["gb"], head.text_projector.layer_norm.bias.grad) < 1e-4
        step.bind_grads()                                  # the eager step re-bound .grad: point it at the graph's tensors again


def test_graphed_step_rejects_other_shapes_and_use_after_close():
    import b200clip
    import pytest
    d = dev()
    head = b200clip.ClipHead(768, 768, 512, 16).to(d)
    xi, xt, ct, lab = _inputs(1, 256, 768, 512, 16, d)
    step = b200clip.GraphedHeadStep(head, xi, xt, ct, lab)
    with pytest.raises(RuntimeError):
        step(xi[:128], xt[:128], ct, lab[:128])
    step.close()
    with pytest.raises(RuntimeError):
        step()


def test_graphed_step_refuses_dropout():
    import b200clip
    import pytest
    d = dev()
    head = b200clip.ClipHead(768, 768, 512, 16, dropout_rate=0.1).to(d).train()
    xi, xt, ct, lab = _inputs(1, 256, 768, 512, 16, d)
    with pytest.raises(RuntimeError):
        b200clip.GraphedHeadStep(head, xi, xt, ct, lab)
