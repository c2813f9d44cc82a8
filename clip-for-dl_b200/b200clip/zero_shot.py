"""Zero-shot scoring cores.  The reference's predict_zero_shot (0426/disease_analysis.py:291-364)
runs encoders, then scores; these functions replace everything after the projector (normalise, similarities,
softmax/sigmoid, top-k / threshold) and return tensors instead of per-row python lists."""
from __future__ import annotations

from typing import Dict, List, Sequence, Union

import torch

from . import ops
from .modules import MODEL_CONFIG


def _prep(image_features, text_features):
    x = ops.cast_bf16(image_features.reshape(image_features.shape[0], -1))
    p = ops.cast_bf16(text_features.reshape(-1, text_features.shape[-1]))
    return x, p


@torch.no_grad()
def zero_shot_topk(image_features, text_features, top_k: int = 3, temperature: float = MODEL_CONFIG["temperature"]):
    """Z1: softmax(normalize(I) T^T / tau).topk(k) -- 0426/disease_analysis.py:332-351.  Returns (idx uint8 [N,k], prob)."""
    x, p = _prep(image_features, text_features)
    k = min(top_k, p.shape[0])
    out = ops.zeroshot_score(x, p, pair_mode=False, temperature=temperature, topk=k, value_mode=1, want_mask=False)
    return out["topk_idx"], out["topk_val"]


@torch.no_grad()
def zero_shot_threshold(image_features, text_features, threshold: Union[float, Sequence[float]] = 0.5,
                        temperature: float = 0.5, inclusive: bool = True):
    """Z2: sigmoid(normalize(I) T^T / 0.5) >= thr (scalar or per-class) -- multimodal_attention/disease_analysis.py:350-378.
    Returns (mask bits [N] int, argmax uint8 [N])."""
    x, p = _prep(image_features, text_features)
    thr = [float(threshold)] if isinstance(threshold, (int, float)) else [float(t) for t in threshold]
    out = ops.zeroshot_score(x, p, pair_mode=False, temperature=temperature, thresholds=thr, thr_inclusive=inclusive)
    return out["mask"], out["argmax"]


@torch.no_grad()
def zero_shot_posneg(image_features, prompts, temperature: float = MODEL_CONFIG["temperature"], threshold: float = 0.5):
    """north-star shape: prompts [L,2,D] (positive, negative), q_l = softmax([l+, l-])[0]; returns
    (argmax uint8 [N], mask int16 [N] with bit l = q_l > thr)."""
    L = prompts.shape[0]
    x, p = _prep(image_features, prompts.reshape(2 * L, -1))
    out = ops.zeroshot_score(x, p, pair_mode=True, temperature=temperature, thresholds=[threshold])
    return out["argmax"], out["mask"]


def unpack_mask(mask: torch.Tensor, num_labels: int) -> torch.Tensor:
    """bit-packed label set -> bool [N, L]"""
    bits = torch.arange(num_labels, device=mask.device, dtype=torch.int32)
    return ((mask.to(torch.int32).unsqueeze(-1) >> bits) & 1).bool()


def _class_text_features(models: Dict, disease_list, text_features, text_fn):
    """The [C, D] normalised class-text matrix of one predict_zero_shot call.  Order of precedence: an explicit tensor,
    the patched reference module's own get_prediction_text_features (stock-PyTorch BERT + the text projector, exactly what
    the reference does per call, 0426/disease_analysis.py:335-340), a cached models['text_features']."""
    if text_features is not None:
        return text_features
    if text_fn is not None and all(k in models for k in ("tokenizer", "text_model", "text_projector")):
        return text_fn(disease_list, models["tokenizer"], models["text_model"], models["text_projector"])
    if "text_features" in models:
        return models["text_features"]
    raise RuntimeError("predict_zero_shot: need models['tokenizer'/'text_model'/'text_projector'] (reference calling "
                       "convention, with b200clip.install() on the module) or a text_features tensor")


def _image_features(images, models: Dict, device):
    """images -> encoder (stock PyTorch) -> projector (B200 kernels); returns ([N, D] features, is_batch)."""
    for m in models.values():
        if hasattr(m, "eval"):
            m.eval()
    is_batch = isinstance(images, torch.Tensor) and images.dim() == 4
    if not is_batch:
        if isinstance(images, list):
            images = images[0]
        images = images.unsqueeze(0)
    dev = device if device is not None else next(models["image_projector"].parameters()).device
    emb = models["resnet"](images.to(dev))
    return models["image_projector"](emb.reshape(emb.size(0), -1)), is_batch


@torch.no_grad()
def predict_zero_shot(images, models: Dict, disease_list: List[str], top_k: int = 3, prompts=None,
                      use_enhanced_prompts: bool = False, text_features: torch.Tensor = None, _text_fn=None, _device=None):
    """Drop-in for 0426/disease_analysis.py:291-364 (same positional order, same return types).  Encoders stay stock
    PyTorch (models['resnet'], models['text_model']); the projector and everything after it run on the B200 kernels:
    F.normalize -> similarities / tau -> softmax -> topk is ONE kernel and ONE D2H copy for the batch (the reference
    does a topk + two .cpu() per row).  `text_features` [C, D] (normalised) skips the per-call BERT pass."""
    feats, is_batch = _image_features(images, models, _device)
    tf = _class_text_features(models, disease_list, text_features, _text_fn)
    idx, val = zero_shot_topk(feats, tf, min(top_k, len(disease_list)))
    idx_c, val_c = idx.cpu().numpy(), val.cpu().numpy()
    if is_batch:
        return [[disease_list[j] for j in row] for row in idx_c], [row for row in val_c]
    return [{"disease": disease_list[j], "confidence": float(s)} for j, s in zip(idx_c[0], val_c[0])]


@torch.no_grad()
def predict_zero_shot_multimodal(images, models: Dict, disease_list: List[str], threshold=0.5, top_k=None, prompts=None,
                                 use_enhanced_prompts: bool = False, text_features: torch.Tensor = None, _text_fn=None,
                                 _device=None):
    """Drop-in for multimodal_attention/disease_analysis.py:291-421: optional MultiModalAttention, sigmoid(cos / 0.5),
    `threshold` a float or a per-disease dict (diseases missing from the dict never pass, :370-373), `top_k=None` ->
    argmax fallback for empty sets, top-k fill / truncation otherwise (:385-408).  The device produces the scores, the
    pass mask (>= logit(thr), exact) and the arg-max in one kernel; the per-row python lists are assembled from ONE D2H
    copy instead of `.item()` per element."""
    import numpy as np
    feats, is_batch = _image_features(images, models, _device)
    tf = _class_text_features(models, disease_list, text_features, _text_fn)
    if "multimodal_attention" in models:                                  # :344-347
        feats, _ = models["multimodal_attention"](ops.normalize(feats), tf)
    L = len(disease_list)
    if isinstance(threshold, dict):
        thr = [float(threshold[d]) if d in threshold else 2.0 for d in disease_list]     # 2.0 -> logit = +inf: never passes
    else:
        thr = [float(threshold)] * L
    x, p = _prep(feats, tf)
    out = ops.zeroshot_score(x, p, pair_mode=False, temperature=0.5, thresholds=thr, thr_inclusive=True, want_scores=True)
    scores = out["scores"].cpu().numpy()
    mask = out["mask"].cpu().numpy().astype(np.int64) & ((1 << L) - 1)
    prob = (1.0 / (1.0 + np.exp(-scores.astype(np.float32)))).astype(np.float32)
    batch_predictions, batch_scores = [], []
    for i in range(prob.shape[0]):
        pi = prob[i]
        passed = [j for j in range(L) if (mask[i] >> j) & 1]
        predictions = [disease_list[j] for j in passed]
        sc = [float(pi[j]) for j in passed]
        if len(predictions) == 0 or (top_k is not None and len(predictions) < top_k):
            k = top_k if top_k is not None else 1
            order = sorted(range(L), key=lambda j: (-scores[i, j], j))[:k]     # torch.topk: descending, first index on ties
            if predictions:
                have = set(predictions)
                for j in order:
                    if disease_list[j] not in have:
                        predictions.append(disease_list[j])
                        sc.append(float(pi[j]))
                        if len(predictions) >= k:
                            break
            else:
                predictions = [disease_list[j] for j in order]
                sc = [float(pi[j]) for j in order]
        elif top_k is not None and len(predictions) > top_k:
            pairs = sorted(zip(predictions, sc), key=lambda t: t[1], reverse=True)[:top_k]
            predictions, sc = [a for a, _ in pairs], [b for _, b in pairs]
        batch_predictions.append(list(predictions))
        batch_scores.append(list(sc))
    if is_batch:
        return batch_predictions, batch_scores
    return [{"disease": d, "confidence": float(s)} for d, s in zip(batch_predictions[0], batch_scores[0])]


def bind_predict_zero_shot(module):
    """The predict_zero_shot that replaces `module.predict_zero_shot`: picks the variant from the reference function's own
    signature (`threshold` parameter = multimodal_attention) and binds the module's get_prediction_text_features and DEVICE,
    so unmodified callers (0426/zero_shot_predict.py:71-78, multimodal_attention/zero_shot_predict.py:73-157) work."""
    import functools
    import inspect
    ref = getattr(module, "predict_zero_shot", None)
    multimodal = ref is not None and "threshold" in inspect.signature(ref).parameters
    text_fn = getattr(module, "get_prediction_text_features", None)
    device = getattr(module, "DEVICE", None)
    if device is not None and not (isinstance(device, str) and device.startswith("cuda") or getattr(device, "type", "") == "cuda"):
        device = None                                                     # reference configured for CPU: use the projector's device
    impl = predict_zero_shot_multimodal if multimodal else predict_zero_shot
    fn = functools.partial(impl, _text_fn=text_fn, _device=device)
    functools.update_wrapper(fn, impl)
    return fn


# --------------------------------------------------------------------------------------------------------------
# zero-shot post-processing on the device (multimodal_attention/zero_shot_predict.py:66-213)
# --------------------------------------------------------------------------------------------------------------
def dynamic_thresholds(max_scores: torch.Tensor, labels: torch.Tensor, return_f1: bool = False):
    """Per-label thresholds from a validation slice (zero_shot_predict.py:112-159): max_scores [N, L] = per-sample maximum over
    the two views of the sigmoid scores (:96-103), labels [N, L] in {0, 1}.  Returns a float64 CUDA tensor [L] (and the best F1
    per label).  The reference does this with numpy + sklearn in a python loop over labels and 20 grid points."""
    from . import ops as _ops
    from ._lib import check, load, ptr, require_cuda, stream_ptr
    require_cuda(max_scores, labels)
    s, y = _ops._f32c(max_scores), _ops._f32c(labels)
    if s.dim() != 2 or s.shape != y.shape or s.shape[1] > 32:
        raise RuntimeError("dynamic_thresholds: scores and labels must both be [N, L] with L <= 32")
    N, L = s.shape
    lib = load()
    thr = torch.empty((L,), dtype=torch.float64, device=s.device)
    f1 = torch.empty((L,), dtype=torch.float64, device=s.device)
    ws = torch.empty(max(int(lib.b200clip_zs_thresholds_workspace_bytes(N, L)), 256), dtype=torch.uint8, device=s.device)
    check(lib.b200clip_zs_dynamic_thresholds(ptr(s), ptr(y), N, L, ptr(thr), ptr(f1), ptr(ws), ws.numel(), stream_ptr()),
          "zs_dynamic_thresholds")
    return (thr, f1) if return_f1 else thr


def merge_two_views(prob_views: torch.Tensor, thresholds: torch.Tensor, weights=(1.0, 0.8), return_scores: bool = False):
    """Weighted two-view merge (zero_shot_predict.py:183-221): prob_views [N, 2, L] sigmoid scores (frontal, lateral), thresholds
    [L]; per view the labels that pass their threshold (or the single best one), weighted maximum over the views (1.0 / 0.8),
    per-label filter, fallback to the best label.  Returns the {0,1} prediction matrix [N, L] (uint8) on the device."""
    from . import ops as _ops
    from ._lib import check, load, ptr, require_cuda, stream_ptr
    require_cuda(prob_views, thresholds)
    p = _ops._f32c(prob_views)
    if p.dim() != 3 or p.shape[1] != 2 or p.shape[2] > 32 or thresholds.numel() != p.shape[2]:
        raise RuntimeError("merge_two_views: prob_views must be [N, 2, L] with L <= 32 and one threshold per label")
    N, _, L = p.shape
    thr = thresholds.to(torch.float64).contiguous()
    pred = torch.empty((N, L), dtype=torch.uint8, device=p.device)
    merged = torch.empty((N, L), dtype=torch.float32, device=p.device) if return_scores else None
    check(load().b200clip_zs_merge_views(ptr(p), ptr(thr), N, L, float(weights[0]), float(weights[1]), ptr(pred), ptr(merged),
                                         stream_ptr()), "zs_merge_views")
    return (pred, merged) if return_scores else pred
