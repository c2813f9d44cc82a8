"""CPU stand-ins for the C-ABI layer, for ONE purpose: driving the reference's UNMODIFIED callers (train_epoch, validate,
predict_zero_shot callers) through b200clip.install() in the build container, which has the reference but no GPU.  They
replace b200clip.ops' autograd Functions with oracle arithmetic (oracle/ref_head.py) so that everything ABOVE the C ABI --
names, signatures, constructor arguments, state_dict keys, return types, autograd connectivity, install()'s patching -- is
exercised by the reference's own code.  The kernels themselves are checked against the same oracle by the `-m gpu` tests.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

import ref_head as R


class _Projection:
    @staticmethod
    def apply(x, w1, b1, w2, b2, gamma, beta, drop_p=0.0, drop_seed=0):
        proj = x @ w1.T + b1
        f = R.gelu_erf(proj) @ w2.T + b2
        if drop_p > 0:
            f = F.dropout(f, drop_p, training=True)
        return F.layer_norm(f + proj, (w1.shape[0],), gamma, beta, 1e-5)


class _Fusion:
    @staticmethod
    def apply(frontal, lateral, w0, b0, w3, b3, drop_p=0.0, drop_seed=0):
        h = torch.relu(torch.cat([frontal, lateral], dim=1) @ w0.T + b0)
        if drop_p > 0:
            h = F.dropout(h, drop_p, training=True)
        return h @ w3.T + b3


class _Mlbce:
    @staticmethod
    def apply(image_features, text_features, labels, temperature):
        loss = R.multilabel_contrastive_loss(image_features, text_features, labels, temperature)
        bad = torch.isnan(loss) | torch.isinf(loss) | (loss > 1000)
        return loss, bad.to(torch.int32)


class _InfoNCE:
    @staticmethod
    def apply(image_features, text_features, temperature):
        return R.contrastive_loss(image_features, text_features, temperature)


class _Attention:
    @staticmethod
    def apply(image_features, text_features, wi, bi, wt, bt, wa, ba, wo, bo):
        return R.multimodal_attention(image_features, text_features, dict(wi=wi, bi=bi, wt=wt, bt=bt, wa=wa, ba=ba, wo=wo, bo=bo))


def _zeroshot_score(x, prompts, *, pair_mode, temperature, thresholds=None, thr_inclusive=False, normalize_x=True, topk=0,
                    value_mode=0, want_argmax=True, want_mask=True, want_scores=False, **_):
    xs = R.l2_normalize(x.double()) if normalize_x else x.double()
    sc = (xs @ prompts.double().T) / temperature
    if pair_mode:
        sc = sc[:, 0::2] - sc[:, 1::2]
    L = sc.shape[1]
    out = {"argmax": sc.argmax(-1).to(torch.uint8), "mask": None, "topk_idx": None, "topk_val": None, "scores": None}
    if want_mask:
        thr = list(thresholds or [0.5] * L)
        thr = thr * L if len(thr) == 1 else thr
        lg = torch.tensor([float("inf") if t >= 1 else (float("-inf") if t <= 0 else torch.logit(torch.tensor(t, dtype=torch.float64)).item())
                           for t in thr], dtype=torch.float64)
        passed = (sc >= lg) if thr_inclusive else (sc > lg)
        out["mask"] = (passed.long() << torch.arange(L)).sum(-1).to(torch.int32)
    if topk:
        vals, idx = sc.topk(topk, dim=-1)
        if value_mode == 1:
            vals = torch.softmax(sc, -1).gather(1, idx)
        elif value_mode == 2:
            vals = torch.sigmoid(vals)
        out["topk_idx"], out["topk_val"] = idx.to(torch.uint8), vals.float()
    if want_scores:
        out["scores"] = sc.float()
    return out


@contextlib.contextmanager
def cpu_ops():
    """Patch b200clip.ops so that the package's modules / losses run on CPU tensors with oracle arithmetic."""
    from b200clip import ops
    saved = {}
    repl = {"ProjectionFn": _Projection, "FusionFn": _Fusion, "MultilabelContrastiveFn": _Mlbce, "InfoNCEFn": _InfoNCE,
            "AttentionFn": _Attention, "zeroshot_score": _zeroshot_score, "cast_bf16": lambda t: t.float(),
            "normalize": lambda t, dim=-1: R.l2_normalize(t), "require_cuda": lambda *a: None,
            "predict_multilabel_raw": lambda i, t, thr, tau: R.predict_multilabel(i, t, thr, tau),
            "new_dropout_seed": lambda: 0}
    for k, v in repl.items():
        saved[k] = getattr(ops, k)
        setattr(ops, k, v)
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
