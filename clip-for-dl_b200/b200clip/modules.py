"""nn.Module surface of the reference head, same constructor signatures and state_dict keys
 so checkpoints written by the reference's save_checkpoint (0426/train.py:846-860) load.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

# 0426/config.py:19-37 (the values the reference binds at import time)
MODEL_CONFIG = {
    "temperature": 0.07,
    "dropout_rate": 0.1,
    "image_embedding_size": 2048,
    "text_embedding_size": 768,
    "shared_embedding_size": 512,
    "num_labels": 16,
}


class _ProjectionBase(nn.Module):
    """Linear(E,D) -> GELU -> Linear(D,D) -> Dropout -> +residual -> LayerNorm (0426/train.py:84-96)."""

    _first = "proj"

    def __init__(self, in_features: int, shared_embedding_size: int, dropout_rate: float = MODEL_CONFIG["dropout_rate"]):
        super().__init__()
        setattr(self, self._first, nn.Linear(in_features, shared_embedding_size))
        self.gelu = nn.GELU()
        self.fc = nn.Linear(shared_embedding_size, shared_embedding_size)
        self.dropout = nn.Dropout(dropout_rate)
        self.layer_norm = nn.LayerNorm(shared_embedding_size)

    def forward(self, embeddings: torch.Tensor) -> torch.Tensor:
        if embeddings.dim() > 2:                                   # 0426/train.py:86-88
            embeddings = embeddings.reshape(embeddings.size(0), -1)
        # nn.Dropout semantics (0426/train.py:81,93): active only in train mode; the mask comes from a counter-based hash of a
        # fresh seed per call (torch's Philox stream cannot be reproduced in a fused epilogue), and
        # `last_dropout_seed` lets a caller / test reconstruct it with ops.dropout_mask
        p = float(self.dropout.p) if self.training else 0.0
        seed = ops.new_dropout_seed() if p > 0 else 0
        self.last_dropout_seed = seed
        first = getattr(self, self._first)
        return ops.ProjectionFn.apply(embeddings, first.weight, first.bias, self.fc.weight, self.fc.bias,
                                      self.layer_norm.weight, self.layer_norm.bias, p, seed)


class ImageProjection(_ProjectionBase):
    """Drop-in for 0426/train.py:73-96 (keys image_projection.*, fc.*, layer_norm.*)."""
    _first = "image_projection"

    def __init__(self, image_embedding_size, shared_embedding_size, dropout_rate: float = MODEL_CONFIG["dropout_rate"]):
        super().__init__(image_embedding_size, shared_embedding_size, dropout_rate)


class TextProjection(_ProjectionBase):
    """Drop-in for 0426/train.py:98-116 (keys text_projection.*, fc.*, layer_norm.*)."""
    _first = "text_projection"

    def __init__(self, text_embedding_size, shared_embedding_size, dropout_rate: float = MODEL_CONFIG["dropout_rate"]):
        super().__init__(text_embedding_size, shared_embedding_size, dropout_rate)


class MultiViewFusion(nn.Module):
    """Drop-in for 0426/train.py:988-1000: cat[frontal, lateral] -> Linear(1024,512) -> ReLU -> Dropout(0.2) -> Linear(512,512).
    Same no-argument constructor and state_dict keys (fusion.0.*, fusion.3.*); the extra keyword arguments only exist for
    other widths.  Dropout follows nn.Dropout semantics (train mode only), fused into the first GEMM's epilogue."""

    def __init__(self, shared_embedding_size: int = MODEL_CONFIG["shared_embedding_size"], dropout_rate: float = 0.2):
        super().__init__()
        d = shared_embedding_size
        self.fusion = nn.Sequential(nn.Linear(2 * d, d), nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(d, d))

    def forward(self, frontal_view: torch.Tensor, lateral_view: torch.Tensor) -> torch.Tensor:
        p = float(self.fusion[2].p) if self.training else 0.0
        seed = ops.new_dropout_seed() if p > 0 else 0
        self.last_dropout_seed = seed
        return ops.FusionFn.apply(frontal_view, lateral_view, self.fusion[0].weight, self.fusion[0].bias, self.fusion[3].weight,
                                  self.fusion[3].bias, p, seed)


class MultiModalAttention(nn.Module):
    """Drop-in for multimodal_attention/train.py:1069-1110: additive attention of each image over the class texts.
    Same no-argument constructor, sub-module names and state_dict keys (image_proj.*, text_proj.*, attention.*,
    output_proj.*); forward(image_features [B,D], text_features [C,D]) -> (enhanced_features [B,D], attn_weights [B,C])."""

    def __init__(self, shared_embedding_size: int = MODEL_CONFIG["shared_embedding_size"]):
        super().__init__()
        d = shared_embedding_size
        self.image_proj = nn.Linear(d, d)
        self.text_proj = nn.Linear(d, d)
        self.attention = nn.Linear(d, 1)
        self.output_proj = nn.Linear(d, d)

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor):
        return ops.AttentionFn.apply(image_features, text_features, self.image_proj.weight, self.image_proj.bias,
                                     self.text_proj.weight, self.text_proj.bias, self.attention.weight, self.attention.bias,
                                     self.output_proj.weight, self.output_proj.bias)


class ClassificationAdapter(nn.Module):
    """The "C-Adapter": nn.Linear(512,16) + BCEWithLogitsLoss (NB02 c28:50-52).  `forward` = logits (as nn.Linear),
    `loss` = fused Linear+BCE, `predict` = sigmoid(logits) > threshold (NB02 c30:42-43).  state_dict keys weight/bias."""

    def __init__(self, in_features: int = 512, num_labels: int = 16):
        super().__init__()
        lin = nn.Linear(in_features, num_labels)
        self.weight = lin.weight
        self.bias = lin.bias

    def forward(self, x):
        return ops.LinearSmallFn.apply(x, self.weight, self.bias)

    def loss(self, x, labels):
        return ops.FcBceFn.apply(x, self.weight, self.bias, labels)

    @torch.no_grad()
    def predict(self, x, threshold: float = 0.5):
        return ops.fc_bce(x, self.weight, self.bias, None, want_pred=True, threshold=threshold)[4]
