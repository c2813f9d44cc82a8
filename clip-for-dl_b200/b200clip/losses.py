"""Loss / prediction functions with the reference's names and signatures."""
from __future__ import annotations

import logging

import torch

from . import ops
from .modules import MODEL_CONFIG


GENERAL_PATH_MAX_N = 8192


def contrastive_loss(image_features, text_features, temperature=1.0, inputs_normalized=None):
    """0426/train.py:154-176: symmetric cross-entropy of I T^T / tau against arange(B), any inputs.

    Two kernels paths, chosen by what the inputs ARE, never silently wrong:
      * L2-normalised rows (what the reference's only call site passes, :228) -> the flash tensor-core path (bf16 operands,
        no B x B matrix, any B);
      * anything else -> the fp32 general path with true row / column maxima (B <= 8192; un-normalised LayerNorm outputs reach
        |logit| ~ 10^3-10^4 at tau = 0.07, which bf16 operands and a fixed shift cannot represent).
    `inputs_normalized=None` checks the rows on the device (one flag read, like the reference's own host syncs at :198, :224);
    pass True / False to skip the check (True on non-unit rows is a contract violation)."""
    if image_features.shape != text_features.shape:
        # F.cross_entropy(logits[B,C], arange(B)) raises for C < B in the reference too
        raise RuntimeError(f"contrastive_loss: logits must be square, got {tuple(image_features.shape)} x "
                           f"{tuple(text_features.shape)}")
    ops.require_cuda(image_features, text_features)
    if inputs_normalized is None:
        inputs_normalized = ops.rows_are_unit(image_features.detach(), text_features.detach())
    if inputs_normalized:
        return ops.InfoNCEFn.apply(image_features, text_features, float(temperature))
    if image_features.shape[0] > GENERAL_PATH_MAX_N:
        raise RuntimeError(f"contrastive_loss: inputs are not L2-normalised and B={image_features.shape[0]} exceeds the "
                           f"{GENERAL_PATH_MAX_N} rows the fp32 general path covers; normalise the features (b200clip.normalize) "
                           "to use the flash path")
    return ops.InfoNCEGeneralFn.apply(image_features, text_features, float(temperature))


def contrastive_clip_loss_function(text_projection, image_projection, temperature=MODEL_CONFIG["temperature"], mode="eval"):
    """0426/train.py:127-152 (the notebooks' stage-1 loss, NB02 c22): soft targets from both self-similarities, gradients
    flow through the targets.  mode "train" -> scalar loss, "eval" -> logits [B, B], anything else -> None (as the reference)."""
    if mode == "eval":
        return ops.softclip_logits(text_projection, image_projection, float(temperature))
    if mode != "train":
        logging.error("Invalid mode for contrastive loss")
        return None
    return ops.SoftClipFn.apply(text_projection, image_projection, float(temperature))


def multilabel_contrastive_loss(image_features, text_features, labels, temperature=1.0, strict_guard=False):
    """0426/train.py:178-230.  The NaN/Inf/>1000 guard (:224) is evaluated on the device; with strict_guard=True the
    flag is read back (one host sync, like the reference's two) and the reference's fallback is taken."""
    num_classes = text_features.size(0)
    if labels.size(1) != num_classes:                                     # :205-210 (kernel zero-pads)
        logging.error(f"标签格式不正确: shape={labels.shape}, 期望shape=[{labels.size(0)}, {num_classes}]")
        if labels.size(1) > num_classes:
            raise RuntimeError("multilabel_contrastive_loss: more label columns than classes")
    loss, status = ops.MultilabelContrastiveFn.apply(image_features, text_features, labels, float(temperature))
    multilabel_contrastive_loss.last_status = status
    if strict_guard and int(status.item()) != 0:
        logging.error(f"多标签对比损失异常: {loss}")
        return contrastive_loss(ops.normalize(image_features), ops.normalize(text_features), temperature)
    return loss


multilabel_contrastive_loss.last_status = None


def multilabel_asymmetric_loss(logits, targets, gamma_pos=0, gamma_neg=4, clip=0.05, eps=1e-8, reduction='mean'):
    """multimodal_attention/train.py:233-268 (ASL): logits [B, C] (before the sigmoid), targets [B, C] in {0, 1}.
    reduction 'mean' | 'sum' | anything else returns the elementwise loss, like the reference."""
    red = reduction if reduction in ("mean", "sum") else "none"
    return ops.AslFn.apply(logits, targets, float(gamma_pos), float(gamma_neg), float(clip) if clip else 0.0, float(eps), red)


def fc_adapter_bce(x, weight, bias, labels):
    """BCEWithLogitsLoss()(F.linear(x, weight, bias), labels) -- NB02 c29:23-25."""
    return ops.FcBceFn.apply(x, weight, bias, labels)


@torch.no_grad()
def predict_multilabel(image_features, text_features, threshold=0.5):
    """0426/train.py:869-886 (temperature bound from MODEL_CONFIG like the reference)."""
    return ops.predict_multilabel_raw(image_features, text_features, threshold, MODEL_CONFIG["temperature"])
