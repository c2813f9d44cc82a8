"""Oracle (test infrastructure only) for the zero-shot POST-PROCESSING of the multimodal_attention variant -- SURVEY.md 8(f)
rank 3: dynamic per-label thresholds from a validation slice and the weighted two-view merge.  CPU restatement in numpy of
/root/reference/multimodal_attention/zero_shot_predict.py:66-213 (threshold search :112-159, merge :183-213) and of the
per-view prediction lists of multimodal_attention/disease_analysis.py:361-413, with label INDICES instead of disease-name
strings (the reference maps names back through disease_list.index, :218-221).

Parity status: PINNED.  The reference code lives inside `main()` between data loaders and a checkpoint load; round 2 runs
that function unmodified with the data side stubbed (oracle/make_golden_zs_main.py: synthetic loader, stub encoders, the
reference's own projectors / predict_zero_shot / f1_score / evaluate_predictions) and records the thresholds dict it passes to
its second pass and the prediction matrix it hands to evaluate_predictions -> tests/golden/zs_main_golden.npz.
tests/test_oracle_zs_post.py checks this file against that run BIT-EXACTLY (float64 thresholds, {0,1} matrix), and the F1
against scikit-learn.  The CUDA path (csrc/zs_post.cu) is checked against this file and against the same golden in
tests/test_gpu_zs_post.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np


def f1_binary(labels: np.ndarray, preds: np.ndarray) -> float:
    """sklearn.metrics.f1_score(labels, preds, zero_division=0) for binary vectors (positive class 1)."""
    labels, preds = np.asarray(labels).astype(int), np.asarray(preds).astype(int)
    tp = int(np.sum((labels == 1) & (preds == 1)))
    fp = int(np.sum((labels == 0) & (preds == 1)))
    fn = int(np.sum((labels == 1) & (preds == 0)))
    if 2 * tp + fp + fn == 0:
        return 0.0
    return 2.0 * tp / (2.0 * tp + fp + fn)


def view_predictions(prob: np.ndarray, threshold: Union[float, Dict[int, float]], top_k: Optional[int] = None
                     ) -> Tuple[List[int], List[float]]:
    """One view's (label indices, scores) as predict_zero_shot builds them (disease_analysis.py:361-413).
    threshold: scalar (:381-390) or {label index: threshold} (:372-379, labels absent from the dict never pass)."""
    p32 = np.asarray(prob, dtype=np.float32)          # the reference holds the probabilities in a float32 tensor ...
    prob = p32.astype(np.float64)                     # ... and turns the kept ones into python floats (:379, :386)
    L = prob.shape[0]
    # `probabilities[j] >= threshold[disease]` compares a float32 tensor element with a python / numpy scalar: torch keeps the
    # tensor dtype, i.e. the threshold is rounded to float32 before the comparison (:377, :382)
    if isinstance(threshold, dict):
        preds = [j for j in range(L) if j in threshold and p32[j] >= np.float32(threshold[j])]
    else:
        preds = [j for j in range(L) if p32[j] >= np.float32(threshold)]
    scores = [float(prob[j]) for j in preds]
    if len(preds) == 0 or (top_k is not None and len(preds) < top_k):          # :393-410
        k = top_k if top_k is not None else 1
        order = np.argsort(-prob, kind="stable")[:k]                             # torch.topk: descending values
        if preds:
            have = set(preds)
            for idx in order:
                if int(idx) not in have:
                    preds.append(int(idx))
                    scores.append(float(prob[idx]))
                    if len(preds) >= k:
                        break
        else:
            preds = [int(i) for i in order]
            scores = [float(prob[i]) for i in order]
    elif top_k is not None and len(preds) > top_k:                               # :412-416 (stable sort, descending score)
        pairs = sorted(zip(preds, scores), key=lambda x: x[1], reverse=True)[:top_k]
        preds, scores = [p for p, _ in pairs], [s for _, s in pairs]
    return preds, scores


def dynamic_thresholds(max_scores: np.ndarray, labels: np.ndarray, initial: float = 0.3) -> np.ndarray:
    """zero_shot_predict.py:112-159.  max_scores [N, L]: per-sample maximum over the two views of the sigmoid scores (:96-103);
    labels [N, L] in {0, 1}.  Returns the threshold per label."""
    max_scores, labels = np.asarray(max_scores, dtype=np.float64), np.asarray(labels)
    N, L = max_scores.shape
    out = np.full(L, initial, dtype=np.float64)                                  # :66
    if N == 0:
        return out
    for j in range(L):
        s, y = max_scores[:, j], labels[:, j]
        pos, neg = s[y == 1], s[y == 0]
        if len(pos) == 0:                                                        # :121-124
            out[j] = 0.8
            continue
        if len(neg) == 0:                                                        # :127-130
            out[j] = 0.2
            continue
        pos_mean, pos_std, neg_mean, neg_std = np.mean(pos), np.std(pos), np.mean(neg), np.std(neg)   # :133-136
        best_f1, best_thr = 0.0, 0.5                                             # :139-140
        lo, hi = max(0.1, neg_mean - neg_std), min(0.9, pos_mean + pos_std)      # :143-144
        for thr in np.linspace(lo, hi, 20):                                      # :146-151 (strict '>': the first best wins)
            f1 = f1_binary(y, (s >= thr).astype(int))
            if f1 > best_f1:
                best_f1, best_thr = f1, thr
        out[j] = best_thr
    return out


def merge_two_views(view_preds: Sequence[Sequence[int]], view_scores: Sequence[Sequence[float]], thresholds: np.ndarray,
                    weights: Sequence[float] = (1.0, 0.8)) -> Tuple[List[int], List[float]]:
    """zero_shot_predict.py:183-213 for ONE sample: weighted maximum over the views' prediction lists, per-label threshold
    filter, fall back to the single best label.  Dict insertion order (view 0's list, then view 1's new labels) decides ties
    in the fallback exactly as python's max() over dict items does."""
    ds: Dict[int, float] = {}
    for v, (preds, scores) in enumerate(zip(view_preds, view_scores)):
        w = weights[v] if v < len(weights) else weights[-1]                      # :190 (1.0 for the frontal view, 0.8 otherwise)
        for p, s in zip(preds, scores):
            if p not in ds:
                ds[p] = 0
            ds[p] = max(ds[p], s * w)                                            # :192-194
    keep = [(p, s) for p, s in ds.items() if s >= thresholds[p]]                 # :199-202
    if not keep:                                                                 # :205-208
        keep = [max(ds.items(), key=lambda x: x[1])]
    return [p for p, _ in keep], [s for _, s in keep]


def merged_prediction_matrix(prob_views: np.ndarray, thresholds: np.ndarray, top_k: Optional[int] = None) -> np.ndarray:
    """prob_views [N, 2, L] sigmoid scores of the two views -> {0,1} matrix [N, L] (:215-221), thresholds as a per-label array."""
    prob_views = np.asarray(prob_views, dtype=np.float64)
    N, V, L = prob_views.shape
    thr = {j: float(thresholds[j]) for j in range(L)}
    out = np.zeros((N, L))
    for i in range(N):
        vp, vs = zip(*(view_predictions(prob_views[i, v], thr, top_k) for v in range(V)))
        preds, _ = merge_two_views(vp, vs, thresholds)
        out[i, preds] = 1
    return out
