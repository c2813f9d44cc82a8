"""b200clip: B200-native CLIP head behind the reference's nn.Module / loss-function surface.

Public names mirror cjycarrie/CLIP-FOR-DL (0426/train.py, disease_analysis.py): ImageProjection, TextProjection,
contrastive_loss, multilabel_contrastive_loss, predict_multilabel, predict_zero_shot; plus ClassificationAdapter
(the notebook's "C-Adapter"), the fused ClipHead and install() which patches a reference module in place.
"""
from .modules import (MODEL_CONFIG, ClassificationAdapter, ImageProjection, MultiModalAttention,  # noqa: F401
                      MultiViewFusion, TextProjection)
from .losses import (contrastive_clip_loss_function, contrastive_loss, fc_adapter_bce,  # noqa: F401
                     multilabel_asymmetric_loss,
                     multilabel_contrastive_loss, predict_multilabel)
from .zero_shot import (dynamic_thresholds, merge_two_views, predict_zero_shot, predict_zero_shot_multimodal,  # noqa: F401
                        unpack_mask, zero_shot_posneg, zero_shot_threshold, zero_shot_topk)
from . import metrics  # noqa: F401
from .metrics import calculate_multilabel_metrics, inloop_accuracy, prompt_mean_pool  # noqa: F401
from .head import ClipHead, ClipHeadFn, GraphedHeadStep  # noqa: F401
from .ops import normalize  # noqa: F401
from .install import install  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__version__ = "0.1.0"
