"""CPU oracle: a restatement of the reference's CLIP-head arithmetic in plain PyTorch (fp32 / fp64).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (clip-for-dl_b200/) imports this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the
checker or as the reported CPU baseline -- never as the thing shipped.

Parity status: the reference (cjycarrie/CLIP-FOR-DL) has NO tests, golden vectors or known-answer
fixtures for this path -> "parity unpinned" by the reference's own tests.
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports
the unmodified reference modules from /root/reference and stores their outputs on seeded inputs under
tests/golden/; tests/test_oracle_golden.py checks every function below against those fixtures.

The arithmetic lives in PyTorch (un-pinned dependency: 0426/requirements.txt `torch>=2.5.1`); this
container runs torch 2.11.0 CPU.  Each function cites the reference file:line it follows
(paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# a-P1 / a-P2  projection block                                   0426/train.py:73-96, 98-116
# ----------------------------------------------------------------------------------------------


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (0426/train.py:79)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def projection_forward(x: Tensor, p: Dict[str, Tensor], ln_eps: float = 1e-5,
                       return_intermediates: bool = False):
    """ImageProjection.forward / TextProjection.forward with dropout off (eval, or p=0).

    0426/train.py:84-96:  flatten -> Linear(E,D) -> GELU -> Linear(D,D) -> Dropout -> +residual -> LayerNorm.
    `p` keys: w1[D,E] b1[D] (image_projection / text_projection), w2[D,D] b2[D] (fc), gamma/beta (layer_norm).
    """
    if x.dim() > 2:                                    # :86-88
        x = x.reshape(x.shape[0], -1)
    proj = x @ p["w1"].T + p["b1"]                     # :90 / :110
    h = gelu_erf(proj)                                 # :91
    f = h @ p["w2"].T + p["b2"]                        # :92
    z = f + proj                                       # :94  (dropout :93 is identity here)
    mu = z.mean(dim=-1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)   # LayerNorm uses the biased variance
    rstd = torch.rsqrt(var + ln_eps)
    y = (z - mu) * rstd * p["gamma"] + p["beta"]       # :95
    if return_intermediates:
        return y, {"proj": proj, "h": h, "z": z, "mu": mu, "rstd": rstd}
    return y


# ----------------------------------------------------------------------------------------------
# a-L2  F.normalize                                 0426/train.py:191-192, :971; disease_analysis.py:332
# ----------------------------------------------------------------------------------------------


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


# ----------------------------------------------------------------------------------------------
# a-N  symmetric InfoNCE                                               0426/train.py:154-176
# ----------------------------------------------------------------------------------------------


def contrastive_loss(image_features: Tensor, text_features: Tensor, temperature: float = 1.0) -> Tensor:
    """logits = I T^T / tau (:166); CE against arange in both directions (:173-174); mean of the two (:176)."""
    logits = (image_features @ text_features.T) / temperature
    n = image_features.shape[0]
    tgt = torch.arange(n)
    lse_r = torch.logsumexp(logits, dim=1)
    lse_c = torch.logsumexp(logits, dim=0)
    diag = logits[tgt, tgt]
    return 0.5 * ((lse_r - diag).mean() + (lse_c - diag).mean())


def contrastive_loss_flash(image_features: Tensor, text_features: Tensor, temperature: float,
                           row0: int = 0, shift: Optional[float] = None):
    """The single-exp / fixed-shift restatement the CUDA kernels implement.

    image_features: the LOCAL row block [B_loc, D] starting at global row `row0`;
    text_features: ALL columns [B, D].  Returns the partial statistics a rank contributes:
      r [B_loc]   = sum_j exp(S_ij - m)        (complete for local rows)
      c [B]       = sum_{i local} exp(S_ij - m) (partial column sums, SUM-combined across ranks)
      diag_sum    = sum_{i local} S_ii
    with m = 1/tau (valid because the inputs are L2-normalised so |S| <= 1/tau).
    """
    inv_tau = 1.0 / temperature
    m = inv_tau if shift is None else shift
    S = (image_features @ text_features.T) * inv_tau
    E = torch.exp(S - m)
    r = E.sum(dim=1)
    c = E.sum(dim=0)
    idx = torch.arange(image_features.shape[0])
    diag_sum = S[idx, idx + row0].sum()
    return r, c, diag_sum, m


def flash_loss_from_stats(r: Tensor, c: Tensor, diag_sum: Tensor, m: float, B: int) -> Tensor:
    """loss = m + (sum log r + sum log c)/(2B) - (sum S_ii)/B."""
    return m + (torch.log(r).sum() + torch.log(c).sum()) / (2.0 * B) - diag_sum / B


def contrastive_grads_flash(image_features: Tensor, text_features: Tensor, temperature: float,
                            r: Tensor, c: Tensor, row0: int = 0, shift: Optional[float] = None):
    """Hand-written backward of the flash form for a local row block.

    G_ij = E_ij (1/r_i + 1/c_j)/(2B) - delta_ij/B ; dI = G T / tau ; dT_partial = G^T I / tau.
    `r` are the local rows' row sums, `c` the GLOBAL column sums.
    """
    B = text_features.shape[0]
    inv_tau = 1.0 / temperature
    m = inv_tau if shift is None else shift
    S = (image_features @ text_features.T) * inv_tau
    E = torch.exp(S - m)
    G = E * (1.0 / r[:, None] + 1.0 / c[None, :]) / (2.0 * B)
    idx = torch.arange(image_features.shape[0])
    G[idx, idx + row0] -= 1.0 / B
    dI = (G @ text_features) * inv_tau
    dT = (G.T @ image_features) * inv_tau
    return dI, dT


# ----------------------------------------------------------------------------------------------
# a-S  soft-target CLIP loss                        0426/train.py:118-152, NB02 c22:3-27
# ----------------------------------------------------------------------------------------------


def soft_target_clip_loss(text_projection: Tensor, image_projection: Tensor, temperature: float,
                          mode: str = "eval"):
    logits = (text_projection @ image_projection.T) / temperature                       # :139
    if mode == "eval":                                                                  # :149-150
        return logits
    if mode != "train":                                                                 # :151-153
        return None
    sim_i = image_projection @ image_projection.T                                       # :141
    sim_t = text_projection @ text_projection.T                                         # :142
    targets = torch.softmax((sim_i + sim_t) / 2 * temperature, dim=-1)                  # :143 (not detached)
    texts_loss = (-targets * torch.log_softmax(logits, dim=-1)).sum(1)                  # :144 via :118-125
    images_loss = (-targets.T * torch.log_softmax(logits.T, dim=-1)).sum(1)             # :145
    return ((images_loss + texts_loss) / 2.0).mean()                                    # :146-147


# ----------------------------------------------------------------------------------------------
# a-B  class-balanced multi-label BCE on sigmoid(cos/tau)              0426/train.py:178-230
# ----------------------------------------------------------------------------------------------


def multilabel_contrastive_loss(image_features: Tensor, text_features: Tensor, labels: Tensor,
                                temperature: float = 1.0) -> Tensor:
    In = l2_normalize(image_features)                                                   # :191
    Tn = l2_normalize(text_features)                                                    # :192
    s = (In @ Tn.T) / temperature                                                       # :195
    B, C = labels.shape[0], Tn.shape[0]
    if labels.shape[1] != C:                                                            # :205-210 pad
        padded = torch.zeros(B, C, dtype=labels.dtype)
        padded[:, : labels.shape[1]] = labels
        labels = padded
    sc = s.clamp(-50.0, 50.0)                                                           # :213
    pos = torch.sigmoid(sc)                                                             # :214
    neg = 1 - pos                                                                       # :215
    pos_loss = -(torch.log(pos + 1e-8) * labels).sum() / (labels.sum() + 1e-8)          # :218
    neg_loss = -(torch.log(neg + 1e-8) * (1 - labels)).sum() / ((1 - labels).sum() + 1e-8)  # :219
    loss = (pos_loss + neg_loss) / 2.0                                                  # :221
    if torch.isnan(loss) or torch.isinf(loss) or loss > 1000:                           # :224-228
        return contrastive_loss(In, Tn, temperature)
    return loss


def multilabel_contrastive_grad_scores(s: Tensor, labels: Tensor, eps: float = 1e-8) -> Tensor:
    """d loss / d s for the non-fallback branch (hand backward), s = cos/tau [B,C]."""
    inside = (s.abs() <= 50.0).to(s.dtype)
    p = torch.sigmoid(s.clamp(-50.0, 50.0))
    P = labels.sum() + eps
    N = (1 - labels).sum() + eps
    dpos = -labels * p * (1 - p) / ((p + eps) * P)
    dneg = (1 - labels) * p * (1 - p) / ((1 - p + eps) * N)
    return 0.5 * (dpos + dneg) * inside


# ----------------------------------------------------------------------------------------------
# a-A  FC classification adapter + BCEWithLogits        NB02 c28:50-52, c29:23-25, c30:42-43
# ----------------------------------------------------------------------------------------------


def fc_adapter_logits(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    return x @ weight.T + bias


def bce_with_logits(z: Tensor, y: Tensor) -> Tensor:
    """nn.BCEWithLogitsLoss() (mean): softplus(z) - y z, in the overflow-safe form torch uses."""
    return (torch.clamp_min(z, 0) - z * y + torch.log1p(torch.exp(-z.abs()))).mean()


def fc_adapter_bce(x: Tensor, weight: Tensor, bias: Tensor, labels: Tensor) -> Tensor:
    return bce_with_logits(fc_adapter_logits(x, weight, bias), labels)


def fc_adapter_predict(x: Tensor, weight: Tensor, bias: Tensor, threshold: float = 0.5) -> Tensor:
    """NB02 c30:42-43  sigmoid(classifier(x)) > 0.5"""
    return (torch.sigmoid(fc_adapter_logits(x, weight, bias)) > threshold).float()


# ----------------------------------------------------------------------------------------------
# a-M  in-loop multi-label prediction            0426/train.py:437-447, :869-886
# ----------------------------------------------------------------------------------------------


def predict_multilabel(image_features: Tensor, text_features: Tensor, threshold: float = 0.5,
                       temperature: float = 0.07) -> Tensor:
    sims = (image_features @ text_features.T) / temperature                             # :881 (tau from MODEL_CONFIG)
    return (torch.sigmoid(sims) > threshold).float()                                    # :883-885


def multilabel_batch_metrics(pred: Tensor, labels: Tensor) -> Tuple[Tensor, Tensor]:
    """0426/train.py:441-447: per-sample accuracy mean and per-class accuracy vector."""
    correct = (pred == labels).float()
    return correct.mean(dim=1).mean(), correct.mean(dim=0)


# ----------------------------------------------------------------------------------------------
# a-Z  zero-shot scoring
# ----------------------------------------------------------------------------------------------


def zero_shot_softmax_topk(image_features: Tensor, text_features: Tensor, top_k: int = 3,
                           temperature: float = 0.07):
    """Z1: 0426/disease_analysis.py:329-356.  image_features are projector outputs (un-normalised);
    text_features already normalised [C,D].  Returns (indices [N,k] int64, probs [N,k])."""
    In = l2_normalize(image_features)                                                   # :332
    sims = (In @ text_features.T) / temperature                                         # :343
    probs = torch.softmax(sims, dim=-1)                                                 # :344
    k = min(top_k, text_features.shape[0])                                              # :351
    vals, idx = probs.topk(k, dim=-1)
    return idx, vals


def zero_shot_sigmoid_threshold(image_features: Tensor, text_features: Tensor,
                                threshold: Union[float, Sequence[float]] = 0.5, temperature: float = 0.5):
    """Z2: multimodal_attention/disease_analysis.py:345-413 scoring core (no attention module):
    sigmoid(cos/0.5) >= thr (scalar or per-class).  Returns (mask [N,C] bool, probs [N,C], argmax [N]).
    The reference's python post-processing (top-k fill when nothing passes, :385-408) consumes exactly
    these three things."""
    In = l2_normalize(image_features)
    sims = (In @ text_features.T) / temperature                                         # :350-353
    probs = torch.sigmoid(sims)                                                         # :363
    thr = torch.as_tensor(threshold, dtype=probs.dtype)
    mask = probs >= thr                                                                 # :372 / :378
    return mask, probs, probs.argmax(dim=-1)


def zero_shot_cosine_argmax(image_features: Tensor, text_features: Tensor):
    """Z3: NB02 c41:27-32 argmax of cosine; NB02 c44:24-36 sigmoid(cos) > 0.5."""
    In = l2_normalize(image_features)
    Tn = l2_normalize(text_features)
    sims = In @ Tn.T
    return sims.argmax(dim=-1), torch.sigmoid(sims) > 0.5


def zero_shot_posneg(image_features: Tensor, prompts: Tensor, temperature: float = 0.07,
                     threshold: float = 0.5):
    """north_star shape (defined by this build from reference ops): prompts [L,2,D]
    ordered (positive, negative) per label, already L2-normalised.
      l = normalize(x) P^T / tau ; q_l = softmax([l+, l-])[0] ; set = {l : q_l > thr} ; argmax_l q_l.
    Returns (argmax [N] int64, mask [N,L] bool, q [N,L])."""
    L = prompts.shape[0]
    In = l2_normalize(image_features)
    logits = (In @ prompts.reshape(2 * L, -1).T) / temperature
    q = torch.softmax(logits.reshape(-1, L, 2), dim=-1)[..., 0]
    return q.argmax(dim=-1), q > threshold, q


def zero_shot_lists_topk(image_features: Tensor, text_features: Tensor, disease_list: Sequence[str], top_k: int = 3,
                         temperature: float = 0.07):
    """The python lists 0426/disease_analysis.py:346-356 builds from the probabilities (batch branch): per image the top-k
    disease names and their softmax probabilities."""
    idx, vals = zero_shot_softmax_topk(image_features, text_features, top_k, temperature)
    return [[disease_list[j] for j in row.tolist()] for row in idx], [row.tolist() for row in vals]


def zero_shot_lists_multimodal(image_features: Tensor, text_features: Tensor, disease_list: Sequence[str],
                               threshold: Union[float, Dict[str, float]] = 0.5, top_k: Optional[int] = None,
                               temperature: float = 0.5):
    """multimodal_attention/disease_analysis.py:349-413 after the (optional) attention module: per image, the diseases whose
    sigmoid(cos/0.5) reaches the threshold (:368-383; dict thresholds skip diseases that are not keys, :370-373), topped up
    from the ranking when the set is empty or shorter than top_k (:385-403), truncated to the best top_k when longer (:405-408)."""
    _, probs, _ = zero_shot_sigmoid_threshold(image_features, text_features, 0.0, temperature)
    names, scores = [], []
    for pr in probs:
        if isinstance(threshold, dict):
            keep = [j for j, d in enumerate(disease_list) if d in threshold and bool(pr[j] >= threshold[d])]
        else:
            keep = (pr >= threshold).nonzero().flatten().tolist()
        pn, ps = [disease_list[j] for j in keep], [float(pr[j]) for j in keep]
        if not pn or (top_k is not None and len(pn) < top_k):
            k = 1 if top_k is None else top_k
            vals, idx = pr.topk(k)
            if pn:
                for j, v in zip(idx.tolist(), vals.tolist()):
                    if disease_list[j] not in pn:
                        pn.append(disease_list[j])
                        ps.append(float(v))
                        if len(pn) >= k:
                            break
            else:
                pn, ps = [disease_list[j] for j in idx.tolist()], [float(v) for v in vals]
        elif top_k is not None and len(pn) > top_k:
            order = sorted(range(len(pn)), key=lambda t: ps[t], reverse=True)[:top_k]
            pn, ps = [pn[t] for t in order], [ps[t] for t in order]
        names.append(pn)
        scores.append(ps)
    return names, scores


# ----------------------------------------------------------------------------------------------
# step edges
# ----------------------------------------------------------------------------------------------


def calculate_multilabel_metrics(predictions: Tensor, labels: Tensor) -> Dict[str, float]:
    """0426/train.py:251-302.  Note :277 reduces the top-1 hits with any(dim=0) over the BATCH (so top1_acc is 0 or 100);
    the quirk is part of the reference's behaviour and is kept."""
    hard = (predictions > 0.5).float()                                                   # :261
    same = (hard == labels).float()
    n = labels.shape[0]
    rows = torch.arange(n)
    top1 = predictions.argmax(dim=1)                                                     # :276
    top3 = predictions.topk(k=min(3, predictions.shape[1]), dim=1).indices               # :280
    tp = (hard * labels).sum(dim=1)                                                      # :286
    prec = tp / (hard.sum(dim=1) + 1e-8)                                                 # :288
    rec = tp / (labels.sum(dim=1) + 1e-8)                                                # :289
    f1 = 2 * prec * rec / (prec + rec + 1e-8)                                            # :290
    return {
        "sample_acc": (same.mean(dim=1) * 100).mean().item(),                            # :264
        "label_acc": (same.mean(dim=0) * 100).mean().item(),                             # :267
        "hamming_score": same.mean().item() * 100,                                       # :270
        "exact_match": (hard == labels).all(dim=1).float().mean().item() * 100,          # :273
        "top1_acc": torch.any(labels[rows, top1] == 1, dim=0).float().mean().item() * 100,   # :277
        "top3_acc": torch.any(labels.gather(1, top3) == 1, dim=1).float().mean().item() * 100,   # :281
        "f1_score": f1.mean().item() * 100,                                              # :291
    }


def prompt_mean_pool(prompt_features: Tensor, counts: Sequence[int], renormalize: bool = False) -> Tensor:
    """get_text_features_with_findings' pooling, 0426/disease_analysis.py:490-497 (same arithmetic in get_enhanced_text_features :230-237): per disease, F.normalize the projected prompt
    features and average them (keepdim), concatenated over diseases.  `renormalize` adds the cosine head's re-normalisation."""
    out, o = [], 0
    for c in counts:
        f = l2_normalize(prompt_features[o:o + c]).mean(dim=0, keepdim=True)             # :491-494
        out.append(l2_normalize(f) if renormalize else f)
        o += c
    return torch.cat(out, dim=0)                                                         # :497


# ----------------------------------------------------------------------------------------------
# a-F / a-X   ("next" rows)
# ----------------------------------------------------------------------------------------------


def multi_view_fusion(frontal: Tensor, lateral: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """MultiViewFusion.forward, dropout off.  0426/train.py:988-1000.
    p: w0[D,2D] b0[D] (fusion.0), w3[D,D] b3[D] (fusion.3)."""
    h = torch.relu(torch.cat([frontal, lateral], dim=1) @ p["w0"].T + p["b0"])
    return h @ p["w3"].T + p["b3"]


def multilabel_asymmetric_loss(logits: Tensor, targets: Tensor, gamma_pos: float = 0, gamma_neg: float = 4,
                               clip: float = 0.05, eps: float = 1e-8, reduction: str = "mean") -> Tensor:
    """multimodal_attention/train.py:233-268 (ASL).  Note the negative focusing term uses the UNCLIPPED
    positive probability (:261), while the log term uses the clipped negative probability (:255)."""
    pr = torch.sigmoid(logits)                                                          # :247
    pr_neg = 1 - pr
    if clip is not None and clip > 0:                                                   # :251-252
        pr_neg = (pr_neg + clip).clamp(max=1)
    pos_term = targets * torch.log(pr.clamp(min=eps))                                   # :254
    neg_term = (1 - targets) * torch.log(pr_neg.clamp(min=eps))                         # :255
    if gamma_pos > 0:                                                                   # :257-258
        pos_term = pos_term * (1 - pr) ** gamma_pos
    if gamma_neg > 0:                                                                   # :259-260
        neg_term = neg_term * pr ** gamma_neg
    loss = -(pos_term + neg_term)                                                       # :262
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def multimodal_attention(image_features: Tensor, text_features: Tensor, p: Dict[str, Tensor]):
    """MultiModalAttention.forward, multimodal_attention/train.py:1081-1110 (additive attention over C classes).
    p: wi,bi (image_proj) wt,bt (text_proj) wa[1,D],ba[1] (attention) wo,bo (output_proj)."""
    ip = image_features @ p["wi"].T + p["bi"]                                           # :1092
    tp = text_features @ p["wt"].T + p["bt"]                                            # :1093
    scores = torch.tanh(ip[:, None, :] + tp[None, :, :]) @ p["wa"].T + p["ba"]          # :1101
    w = torch.softmax(scores.squeeze(-1), dim=1)                                        # :1102
    attended = w @ tp                                                                   # :1105
    return (ip + attended) @ p["wo"].T + p["bo"], w                                     # :1108-1110


# ----------------------------------------------------------------------------------------------
# The composite "head alone" step that bench.py times.
# ----------------------------------------------------------------------------------------------


def head_step(x_img: Tensor, x_txt: Tensor, class_text: Tensor, labels: Tensor,
              img_p: Dict[str, Tensor], txt_p: Dict[str, Tensor], fc_w: Tensor, fc_b: Tensor,
              tau_nce: float = 0.07, tau_bce: float = 1.0, quantize=None, nce_fn=None) -> Dict[str, Tensor]:
    """proj(img), proj(txt) -> normalize -> contrastive_loss(tau_nce) + multilabel_contrastive_loss(tau_bce)
    + FC adapter BCEWithLogits; the three losses are summed and back-propagated by the caller."""
    y_img = projection_forward(x_img, img_p)
    y_txt = projection_forward(x_txt, txt_p)
    In, Tn = l2_normalize(y_img), l2_normalize(y_txt)
    if quantize is not None:
        # model of the CUDA path's storage precision (tests only): the normalised embeddings exist as bf16; `quantize` is a
        # straight-through rounding (forward value rounded, gradient passed unchanged), `nce_fn` an InfoNCE whose backward
        # rounds the softmax-gradient tile to bf16 like the kernel does
        In, Tn = quantize(In), quantize(Tn)
        y_img = In * y_img.norm(dim=-1, keepdim=True)
    l_nce = (nce_fn or contrastive_loss)(In, Tn, tau_nce)
    l_bce = multilabel_contrastive_loss(y_img, class_text, labels, tau_bce)
    l_fc = fc_adapter_bce(y_img, fc_w, fc_b, labels)
    return {"loss": l_nce + l_bce + l_fc, "nce": l_nce, "bce": l_bce, "fc": l_fc,
            "y_img": y_img, "y_txt": y_txt}


def head_flops(B: int, D: int, E_img: int, E_txt: int, C: int) -> float:
    """Algorithmic FLOPs per fwd+bwd step, SURVEY.md section 8(d)."""
    return (6.0 * B * B * D + 6.0 * B * (E_img * D + D * D) + 6.0 * B * (E_txt * D + D * D)
            + 6.0 * B * D * C + 6.0 * B * D * C)
