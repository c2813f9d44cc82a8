"""Step edges around the head on the device: the multi-label metrics report, the in-loop
accuracy counters and the per-disease prompt-mean pooling.  Same names / return types as the reference."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import ops
from ._lib import check, load, ptr, require_cuda, stream_ptr

_KEYS = ("sample_acc", "label_acc", "hamming_score", "exact_match", "top1_acc", "top3_acc", "f1_score")


def multilabel_metrics_device(predictions: torch.Tensor, labels: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    """7 + C doubles on the device (no host sync): the seven metrics of calculate_multilabel_metrics in dict order, then the
    per-class accuracies in % (the `class_acc_dict` values of 0426/train.py:445-447)."""
    require_cuda(predictions, labels)
    p, y = ops._f32c(predictions), ops._f32c(labels)
    if p.dim() != 2 or p.shape != y.shape or p.shape[1] > 32:
        raise RuntimeError("calculate_multilabel_metrics: predictions and labels must both be [B, C] with C <= 32")
    B, C = p.shape
    lib = load()
    out = torch.empty((7 + C,), dtype=torch.float64, device=p.device)
    ws = torch.empty(max(int(lib.b200clip_multilabel_metrics_workspace_bytes(B)), 256), dtype=torch.uint8, device=p.device)
    check(lib.b200clip_multilabel_metrics(ptr(p), p.stride(0), ptr(y), y.stride(0), B, C, float(threshold), ptr(out), ptr(ws),
                                          ws.numel(), stream_ptr()), "multilabel_metrics")
    return out


@torch.no_grad()
def calculate_multilabel_metrics(predictions, labels) -> Dict[str, float]:
    """0426/train.py:251-302: one kernel pass and ONE device-to-host copy instead of ~20 torch kernels and 7 `.item()` syncs."""
    vals = multilabel_metrics_device(predictions, labels).cpu().tolist()
    return dict(zip(_KEYS, vals[:7]))


@torch.no_grad()
def inloop_accuracy(image_features, text_features, labels, disease_list: Sequence[str], temperature: float = None,
                    threshold: float = 0.5):
    """The prediction/accuracy block of train_epoch and validate (0426/train.py:437-447, :586-591): predictions =
    sigmoid(I T^T / tau) > 0.5 (kernel of predict_multilabel), accuracy = mean per-sample accuracy in %, class_acc_dict =
    per-class accuracy in %.  One host sync (the reference: 1 + C `.item()` calls)."""
    from .modules import MODEL_CONFIG
    tau = MODEL_CONFIG["temperature"] if temperature is None else temperature
    pred = ops.predict_multilabel_raw(image_features, text_features, threshold, tau)
    vals = multilabel_metrics_device(pred, labels, threshold=0.5).cpu().tolist()
    return pred, vals[0], {d: a for d, a in zip(disease_list, vals[7:])}


@torch.no_grad()
def prompt_mean_pool(prompt_features: torch.Tensor, counts: Sequence[int], renormalize: bool = False) -> torch.Tensor:
    """get_text_features_with_findings' pooling (0426/disease_analysis.py:486-497): rows of `prompt_features` [sum(counts), D] are
    the projected prompt embeddings, grouped per disease; returns [len(counts), D] = per-disease mean of the L2-normalised
    rows (re-normalised when `renormalize`)."""
    require_cuda(prompt_features)
    x = ops._f32c(prompt_features)
    counts = [int(c) for c in counts]
    if x.dim() != 2 or sum(counts) != x.shape[0] or min(counts, default=0) < 1:
        raise RuntimeError("prompt_mean_pool: counts must be >= 1 and sum to the number of prompt rows")
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    off_t = torch.tensor(offs, dtype=torch.int32, device=x.device)
    out = torch.empty((len(counts), x.shape[1]), dtype=torch.float32, device=x.device)
    check(load().b200clip_prompt_mean_pool(ptr(x), ptr(off_t), len(counts), x.shape[1], ops.L2_EPS, int(renormalize), ptr(out),
                                           stream_ptr()), "prompt_mean_pool")
    return out
