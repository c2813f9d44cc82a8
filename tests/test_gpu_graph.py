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
        assert rel_l2(got["gw"], head.image_projector.fc.weight.grad) < 1e-5
        assert rel_l2(got["gf"], head.classifier.weight.grad) < 1e-5
        assert rel_l2(got["gb"], head.text_projector.layer_norm.bias.grad) < 1e-4
        step.bind_grads()                                  # the eager step re-bound .grad: point it at the graph's tensors again


def test_head_step_is_bit_reproducible():
    """Every reduction on the path has a fixed order (split-K weight gradients store per-split partial tiles and add them in
    split order; row/column statistics and dI partials are slot buffers): two runs on the same inputs agree bit for bit."""
    import b200clip
    d = dev()
    B, E, D, C = 1024, 768, 512, 16
    head = b200clip.ClipHead(2048, E, D, C).to(d)
    torch.manual_seed(3)
    xi = torch.randn(B, 2048, device=d).bfloat16()
    xt = torch.randn(B, E, device=d).bfloat16()
    ct = torch.nn.functional.normalize(torch.randn(C, D, device=d), dim=-1)
    lab = (torch.rand(B, C, device=d) < 0.2).float()
    runs = []
    for _ in range(3):
        for p in head.parameters():
            p.grad = None
        a, b = xi.clone().requires_grad_(True), xt.clone().requires_grad_(True)
        loss = head(a, b, ct, lab)
        loss.backward()
        torch.cuda.synchronize()
        runs.append([loss.detach().clone(), a.grad.clone(), b.grad.clone()] + [p.grad.clone() for p in head.parameters()])
    for other in runs[1:]:
        for x, y in zip(runs[0], other):
            assert torch.equal(x, y)


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


def test_graphed_step_trains_with_dropout():
    """nn.Dropout(0.1) inside the captured step (0426/train.py:81,93): the seed is a device word advanced by the graph itself,
    so every replay draws a fresh mask; a replay is reproduced by the eager step given the replay's effective seed, and the
    projection output matches the oracle fed the same mask (ops.dropout_mask)."""
    import b200clip
    import ref_head as R
    import synth
    from b200clip import ops
    d = dev()
    B, E, D, C, pdrop = 512, 768, 512, 16, 0.1
    torch.manual_seed(0)
    head = b200clip.ClipHead(E, E, D, C, dropout_rate=pdrop).to(d).train()
    xi, xt, ct, lab = _inputs(1, B, E, D, C, d)
    step = b200clip.GraphedHeadStep(head, xi, xt, ct, lab)
    losses, seeds = [], []
    for _ in range(3):
        losses.append(float(step(xi, xt, ct, lab)))
        seeds.append(int(step.seed_dev.item()) & 0xFFFFFFFF)
    assert len(set(seeds)) == 3 and len(set(losses)) == 3            # fresh mask per replay
    g_graph = step.grad_image.float().clone()
    # the eager step with the replay's effective seed reproduces the last replay
    eff = (step.drop_seed0 + seeds[-1]) & 0xFFFFFFFF
    xe, te = xi.clone().requires_grad_(True), xt.clone().requires_grad_(True)
    loss_e = b200clip.ClipHeadFn.apply(xe, te, ct, lab, head.tau_nce, head.tau_bce, None, pdrop, eff, *head.params())
    loss_e.backward()
    torch.cuda.synchronize()
    assert abs(float(loss_e) - losses[-1]) <= 1e-6 * abs(losses[-1])
    assert rel_l2(g_graph, xe.grad.float()) < 1e-5
    step.bind_grads()
    # keep-rate of the mask the replay used, and the oracle with that mask on the image side
    mask = ops.dropout_mask(B, D, pdrop, eff, d).cpu()
    assert abs((mask > 0).float().mean().item() - (1 - pdrop)) < 0.01
    ip = head.image_projector
    p = {k: v.detach().cpu().to(torch.bfloat16).float() if k in ("w1", "w2") else v.detach().cpu()
         for k, v in dict(w1=ip.image_projection.weight, b1=ip.image_projection.bias, w2=ip.fc.weight, b2=ip.fc.bias,
                          gamma=ip.layer_norm.weight, beta=ip.layer_norm.bias).items()}
    x = xi.float().cpu()
    proj = x @ p["w1"].T + p["b1"]
    f = (R.gelu_erf(proj) @ p["w2"].T + p["b2"]) * mask
    yref = torch.nn.functional.layer_norm(f + proj, (D,), p["gamma"], p["beta"], 1e-5)
    y, _, _, _ = ops.proj_fwd(xi, ops.cast_bf16(ip.image_projection.weight), ip.image_projection.bias.detach(), ops.cast_bf16(ip.fc.weight),
                              ip.fc.bias.detach(), ip.layer_norm.weight.detach(), ip.layer_norm.bias.detach(), want_yhat=False,
                              drop_p=pdrop, drop_seed=step.drop_seed0, drop_seed_dev=step.seed_dev)
    assert rel_l2(y, yref) < 1e-2
    step.close()
