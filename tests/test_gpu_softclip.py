"""a-S: contrastive_clip_loss_function (soft-target CLIP loss, 0426/train.py:127-152) vs the golden values minted from the
unmodified reference and vs the oracle's autograd at the shapes / temperatures the reference uses it with (LayerNorm-scale
inputs, tau = 0.07 from the config and tau = 2 from the notebooks).  fp32 path: tolerances are fp32 tolerances."""
import numpy as np
import pytest
import torch

from gpu_util import dev, gpu, rel_l2
import ref_head as R
import synth

pytestmark = gpu


def test_soft_target_loss_matches_reference_golden(golden):
    import b200clip
    d = dev()
    Tt = synth.randn(15, 16, 32).to(d).requires_grad_(True)
    Ii = synth.randn(16, 16, 32).to(d).requires_grad_(True)
    loss = b200clip.contrastive_clip_loss_function(Tt, Ii, temperature=2.0, mode="train")
    loss.backward()
    np.testing.assert_allclose(loss.item(), golden["soft_loss"], rtol=1e-5)
    np.testing.assert_allclose(Tt.grad.cpu().numpy(), golden["soft_dT"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(Ii.grad.cpu().numpy(), golden["soft_dI"], rtol=1e-4, atol=1e-6)
    logits = b200clip.contrastive_clip_loss_function(Tt.detach(), Ii.detach(), temperature=2.0, mode="eval")
    np.testing.assert_allclose(logits.cpu().numpy(), golden["soft_logits_eval"], rtol=1e-5, atol=1e-5)
    assert b200clip.contrastive_clip_loss_function(Tt, Ii, 2.0, mode="bogus") is None


# n >= 128 with n % 32 == 0 takes the tensor-core path (K-concatenated 3 x bf16 split operands on the tcgen05 GEMM); the other
# sizes the fp32 SGEMM path
@pytest.mark.parametrize("B,D,tau,scale", [(16, 512, 2.0, 1.0), (32, 512, 0.07, 1.0), (64, 512, 0.07, 0.2), (300, 512, 2.0, 1.0),
                                            (1000, 128, 0.5, 0.5), (128, 512, 0.07, 1.0), (256, 512, 2.0, 1.0), (1024, 512, 0.07, 0.3),
                                            (2048, 768, 0.5, 0.5), (4096, 512, 0.5, 0.05)])
def test_soft_target_loss_and_grads(B, D, tau, scale):
    """LayerNorm-like rows (norm ~ sqrt(D) * scale): at tau = 0.07 the logits reach +-10^3, the regime that needs the true row
    maxima and fp32 Gram products.  The oracle runs in fp64 on the same fp32 inputs."""
    import b200clip
    d = dev()
    T = (synth.randn(1, B, D) * scale)
    I = (0.5 * T + 0.5 * synth.randn(2, B, D) * scale)
    Tr, Ir = T.double().requires_grad_(True), I.double().requires_grad_(True)
    ref = R.soft_target_clip_loss(Tr, Ir, tau, mode="train")
    (3.0 * ref).backward()
    # the reference itself computes in fp32: at tau = 0.07 the logits are ~1e3-1e4 and the gradients are differences of nearly
    # equal softmax terms, so fp32 rounding of the logits alone moves them by percents.  Measure that floor with the fp32 oracle
    # and hold the kernel to the same order.
    T32, I32 = T.clone().requires_grad_(True), I.clone().requires_grad_(True)
    ref32 = R.soft_target_clip_loss(T32, I32, tau, mode="train")
    (3.0 * ref32).backward()
    floor_t, floor_i = rel_l2(T32.grad, Tr.grad), rel_l2(I32.grad, Ir.grad)
    floor_l = abs(ref32.item() - ref.item())
    Tg, Ig = T.to(d).requires_grad_(True), I.to(d).requires_grad_(True)
    loss = b200clip.contrastive_clip_loss_function(Tg, Ig, temperature=tau, mode="train")
    (3.0 * loss).backward()
    assert abs(loss.item() - ref.item()) <= max(1e-4 * abs(ref.item()) + 1e-6, 4.0 * floor_l)
    assert rel_l2(Tg.grad, Tr.grad) < max(1e-3, 4.0 * floor_t), (rel_l2(Tg.grad, Tr.grad), floor_t)
    assert rel_l2(Ig.grad, Ir.grad) < max(1e-3, 4.0 * floor_i), (rel_l2(Ig.grad, Ir.grad), floor_i)
    ev = b200clip.contrastive_clip_loss_function(Tg.detach(), Ig.detach(), temperature=tau)          # default mode: eval
    assert rel_l2(ev, (T.double() @ I.double().T) / tau) < 1e-5


def test_soft_target_loss_rejects_bad_input():
    import b200clip
    d = dev()
    with pytest.raises(RuntimeError):
        b200clip.contrastive_clip_loss_function(torch.zeros(4, 32, device=d), torch.zeros(5, 32, device=d), 0.07, mode="train")
    with pytest.raises(RuntimeError):
        b200clip.contrastive_clip_loss_function(torch.zeros(4, 32), torch.zeros(4, 32), 0.07, mode="train")


def test_tensor_core_path_agrees_with_fp32_path():
    """Same call, the two contraction engines: tcgen05 GEMMs over 3 x bf16 split operands vs the fp32 CUDA-core SGEMM
    (B200CLIP_SOFTCLIP_FP32=1 forces the latter)."""
    import os
    import b200clip
    d = dev()
    T = synth.randn(1, 512, 512).to(d)                          # tau = 2 on LayerNorm-scale rows: the notebooks' regime, loss O(1)
    I = (0.5 * T.cpu() + 0.5 * synth.randn(2, 512, 512)).to(d)
    res = {}
    for mode in ("tc", "fp32"):
        if mode == "fp32":
            os.environ["B200CLIP_SOFTCLIP_FP32"] = "1"
        try:
            Tg, Ig = T.clone().requires_grad_(True), I.clone().requires_grad_(True)
            loss = b200clip.contrastive_clip_loss_function(Tg, Ig, temperature=2.0, mode="train")
            loss.backward()
            torch.cuda.synchronize()
            res[mode] = (loss.item(), Tg.grad.clone(), Ig.grad.clone())
        finally:
            os.environ.pop("B200CLIP_SOFTCLIP_FP32", None)
    assert abs(res["tc"][0] - res["fp32"][0]) <= 1e-4 * abs(res["fp32"][0])
    assert rel_l2(res["tc"][1], res["fp32"][1]) < 2e-3 and rel_l2(res["tc"][2], res["fp32"][2]) < 2e-3
