"""Pins the oracle (oracle/ref_head.py) against outputs of the UNMODIFIED reference
(tests/golden/head_golden.npz, produced by oracle/make_golden.py from /root/reference)."""
import numpy as np
import torch

import ref_head as R
import synth

TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), **kw)


def _proj_case(tag):
    E = 96 if tag == "img" else 80
    p = synth.projection_params(100 if tag == "img" else 200, E, 64)
    x = synth.randn(7 if tag == "img" else 8, 24, E)
    return p, x


def test_projection_forward_backward(golden):
    for tag in ("img", "txt"):
        p, x = _proj_case(tag)
        p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        x = x.clone().requires_grad_(True)
        y = R.projection_forward(x, p)
        close(y.detach(), golden[f"proj_{tag}_y"])
        (y * synth.randn(9, 24, 64)).sum().backward()
        close(x.grad, golden[f"proj_{tag}_dx"], atol=1e-5)
        for k, g in (("w1", "dw1"), ("b1", "db1"), ("w2", "dw2"), ("b2", "db2"), ("gamma", "dgamma"), ("beta", "dbeta")):
            close(p[k].grad, golden[f"proj_{tag}_{g}"], atol=2e-5)


def test_projection_flattens_4d(golden):
    p, x = _proj_case("img")
    close(R.projection_forward(x.reshape(24, 6, 4, 4), p), golden["proj_img_y_4d"])


def test_contrastive_loss_and_grads(golden):
    for tau in (0.07, 1.0):
        I = synth.unit_rows(11, 48, 64).requires_grad_(True)
        T = synth.unit_rows(12, 48, 64).requires_grad_(True)
        loss = R.contrastive_loss(I, T, tau)
        loss.backward()
        close(loss.detach(), golden[f"nce_loss_tau{tau}"])
        close(I.grad, golden[f"nce_dI_tau{tau}"], atol=1e-6)
        close(T.grad, golden[f"nce_dT_tau{tau}"], atol=1e-6)
    close(R.contrastive_loss(synth.randn(13, 20, 32), synth.randn(14, 20, 32), 1.0), golden["nce_loss_unnorm"])


def test_flash_form_equals_reference(golden):
    """The fixed-shift single-exp restatement (what the kernels compute) == reference loss and grads."""
    for tau in (0.07, 1.0):
        I = synth.unit_rows(11, 48, 64).double()
        T = synth.unit_rows(12, 48, 64).double()
        r, c, d, m = R.contrastive_loss_flash(I, T, tau)
        loss = R.flash_loss_from_stats(r, c, d, m, 48)
        close(loss, golden[f"nce_loss_tau{tau}"], rtol=1e-6)
        dI, dT = R.contrastive_grads_flash(I, T, tau, r, c)
        close(dI, golden[f"nce_dI_tau{tau}"], rtol=1e-4, atol=1e-7)
        close(dT, golden[f"nce_dT_tau{tau}"], rtol=1e-4, atol=1e-7)


def test_flash_form_rank_partitioned(golden):
    """Simulated W-rank row partition: row stats local, column sums SUM-combined, dT partials summed."""
    tau, B = 0.07, 48
    I = synth.unit_rows(11, B, 64).double()
    T = synth.unit_rows(12, B, 64).double()
    for W in (2, 4, 8):
        n = B // W
        parts = [R.contrastive_loss_flash(I[k * n:(k + 1) * n], T, tau, row0=k * n) for k in range(W)]
        c = sum(p[1] for p in parts)
        r = torch.cat([p[0] for p in parts])
        d = sum(p[2] for p in parts)
        close(R.flash_loss_from_stats(r, c, d, parts[0][3], B), golden[f"nce_loss_tau{tau}"], rtol=1e-6)
        dT = torch.zeros_like(T)
        dIs = []
        for k in range(W):
            dI_k, dT_k = R.contrastive_grads_flash(I[k * n:(k + 1) * n], T, tau, r[k * n:(k + 1) * n], c, row0=k * n)
            dIs.append(dI_k)
            dT += dT_k
        close(torch.cat(dIs), golden[f"nce_dI_tau{tau}"], rtol=1e-4, atol=1e-7)
        close(dT, golden[f"nce_dT_tau{tau}"], rtol=1e-4, atol=1e-7)


def test_soft_target_loss(golden):
    Tt = synth.randn(15, 16, 32).requires_grad_(True)
    Ii = synth.randn(16, 16, 32).requires_grad_(True)
    loss = R.soft_target_clip_loss(Tt, Ii, 2.0, mode="train")
    loss.backward()
    close(loss.detach(), golden["soft_loss"])
    close(Tt.grad, golden["soft_dT"], atol=1e-6)
    close(Ii.grad, golden["soft_dI"], atol=1e-6)
    close(R.soft_target_clip_loss(Tt.detach(), Ii.detach(), 2.0, mode="eval"), golden["soft_logits_eval"])
    assert R.soft_target_clip_loss(Tt, Ii, 2.0, mode="bogus") is None


def test_multilabel_contrastive_loss(golden):
    for tau in (1.0, 0.07):
        I = synth.randn(21, 40, 64).requires_grad_(True)
        T = synth.randn(22, 16, 64)
        y = synth.labels(23, 40, 16, density=0.2)
        loss = R.multilabel_contrastive_loss(I, T, y, tau)
        loss.backward()
        close(loss.detach(), golden[f"mlbce_loss_tau{tau}"])
        close(I.grad, golden[f"mlbce_dI_tau{tau}"], atol=1e-6)
        # hand-written d/ds chained through the normalisation == autograd
        In = R.l2_normalize(I.detach())
        Tn = R.l2_normalize(T)
        gs = R.multilabel_contrastive_grad_scores((In @ Tn.T) / tau, y)
        dIn = (gs @ Tn) / tau
        nrm = I.detach().norm(dim=1, keepdim=True)
        dI = (dIn - In * (In * dIn).sum(1, keepdim=True)) / nrm
        close(dI, golden[f"mlbce_dI_tau{tau}"], rtol=1e-4, atol=1e-6)
    close(R.multilabel_contrastive_loss(synth.randn(21, 40, 64), synth.randn(22, 16, 64),
                                        synth.labels(24, 40, 12, density=0.2), 1.0), golden["mlbce_loss_padded"])
    close(R.multilabel_contrastive_loss(synth.randn(21, 40, 64), synth.randn(22, 16, 64),
                                        torch.zeros(40, 16), 1.0), golden["mlbce_loss_nolabels"])


def test_fc_adapter(golden):
    w = synth.uniform(31, -0.125, 0.125, 16, 64).requires_grad_(True)
    b = synth.uniform(32, -0.125, 0.125, 16).requires_grad_(True)
    x = synth.randn(33, 40, 64).requires_grad_(True)
    y = synth.labels(34, 40, 16, density=0.2)
    loss = R.fc_adapter_bce(x, w, b, y)
    loss.backward()
    close(loss.detach(), golden["fc_loss"])
    close(x.grad, golden["fc_dx"], atol=1e-7)
    close(w.grad, golden["fc_dw"], atol=1e-7)
    close(b.grad, golden["fc_db"], atol=1e-7)
    assert np.array_equal(R.fc_adapter_predict(x.detach(), w.detach(), b.detach()).numpy(), golden["fc_pred"])


def test_predict_multilabel(golden):
    I = synth.randn(41, 40, 64)
    T = synth.unit_rows(42, 16, 64)
    assert np.array_equal(R.predict_multilabel(I, T, 0.5).numpy(), golden["predict_multilabel"])
    assert np.array_equal(R.predict_multilabel(I, T, 0.7).numpy(), golden["predict_multilabel_thr0.7"])


def test_zero_shot_modes(golden):
    X = synth.randn(51, 200, 64)
    T16 = synth.unit_rows(52, 16, 64)
    idx, vals = R.zero_shot_softmax_topk(X, T16, 3, 0.07)
    assert np.array_equal(idx.numpy(), golden["z1_idx"])
    close(vals, golden["z1_vals"])
    mask, probs, am = R.zero_shot_sigmoid_threshold(X, T16, 0.5, 0.5)
    assert np.array_equal(mask.numpy(), golden["z2_mask"])
    assert np.array_equal(am.numpy(), golden["z2_argmax"])
    mask_pl, _, _ = R.zero_shot_sigmoid_threshold(X, T16, torch.linspace(0.45, 0.6, 16), 0.5)
    assert np.array_equal(mask_pl.numpy(), golden["z2_mask_perlabel"])
    am3, m3 = R.zero_shot_cosine_argmax(X, T16)
    assert np.array_equal(am3.numpy(), golden["z3_argmax"])
    assert np.array_equal(m3.numpy(), golden["z3_mask"])
    P = synth.unit_rows(53, 28, 64).reshape(14, 2, 64)
    amn, mn, q = R.zero_shot_posneg(X, P, 0.07, 0.5)
    assert np.array_equal(amn.numpy(), golden["zn_argmax"])
    assert np.array_equal(mn.numpy(), golden["zn_mask"])
    close(q, golden["zn_q"])


def test_fusion_asl_attention(golden):
    fp = {"w0": synth.uniform(61, -0.03, 0.03, 512, 1024), "b0": synth.uniform(62, -0.03, 0.03, 512),
          "w3": synth.uniform(63, -0.04, 0.04, 512, 512), "b3": synth.uniform(64, -0.04, 0.04, 512)}
    close(R.multi_view_fusion(synth.randn(65, 6, 512), synth.randn(66, 6, 512), fp), golden["fusion_y"], atol=1e-5)
    lg = synth.randn(71, 40, 16) * 3
    y = synth.labels(72, 40, 16, density=0.2)
    close(R.multilabel_asymmetric_loss(lg, y), golden["asl_mean"])
    close(R.multilabel_asymmetric_loss(lg, y, gamma_pos=1, gamma_neg=2, clip=0.1, reduction="sum"), golden["asl_sum_g1"])
    ap = {"wi": synth.uniform(81, -0.04, 0.04, 512, 512), "bi": synth.uniform(82, -0.04, 0.04, 512),
          "wt": synth.uniform(83, -0.04, 0.04, 512, 512), "bt": synth.uniform(84, -0.04, 0.04, 512),
          "wa": synth.uniform(85, -0.04, 0.04, 1, 512), "ba": synth.uniform(86, -0.04, 0.04, 1),
          "wo": synth.uniform(87, -0.04, 0.04, 512, 512), "bo": synth.uniform(88, -0.04, 0.04, 512)}
    enh, w = R.multimodal_attention(synth.randn(89, 6, 512), synth.unit_rows(90, 16, 512), ap)
    close(enh, golden["attn_enh"], atol=1e-5)
    close(w, golden["attn_w"])


def test_head_flops_formula():
    # cfg 2 = 1.00e11, cfg 3 ViT = 3.56e12, ResNet at 32k = 3.69e12
    assert abs(R.head_flops(4096, 512, 2048, 768, 16) / 1.00e11 - 1) < 0.01
    assert abs(R.head_flops(32768, 512, 768, 768, 16) / 3.56e12 - 1) < 0.01
    assert abs(R.head_flops(32768, 512, 2048, 768, 16) / 3.69e12 - 1) < 0.01
