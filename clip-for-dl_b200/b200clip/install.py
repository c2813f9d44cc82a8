"""Drop-in installation: patch the names an imported reference module (`train`, `disease_analysis`) looks up as module
globals (SURVEY.md section 8b "Who calls it").  Nothing in the reference is edited; see INTEGRATION.md."""
from __future__ import annotations

import types


def install(*modules: types.ModuleType) -> None:
    from . import losses, metrics, modules as mods, zero_shot

    table = {
        "ImageProjection": mods.ImageProjection,
        "TextProjection": mods.TextProjection,
        "MultiViewFusion": mods.MultiViewFusion,
        "MultiModalAttention": mods.MultiModalAttention,
        "contrastive_loss": losses.contrastive_loss,
        "contrastive_clip_loss_function": losses.contrastive_clip_loss_function,
        "multilabel_contrastive_loss": losses.multilabel_contrastive_loss,
        "multilabel_asymmetric_loss": losses.multilabel_asymmetric_loss,
        "predict_multilabel": losses.predict_multilabel,
        "calculate_multilabel_metrics": metrics.calculate_multilabel_metrics,
    }
    for m in modules:
        # predict_zero_shot has two reference signatures (0426/disease_analysis.py:291-298 and
        # multimodal_attention/disease_analysis.py:291-299); bind the one this module defines, together with the module's
        # own get_prediction_text_features / DEVICE, BEFORE the name is overwritten
        pzs = zero_shot.bind_predict_zero_shot(m) if hasattr(m, "predict_zero_shot") else None
        for name, obj in table.items():
            if hasattr(m, name):
                setattr(m, name, obj)
        if pzs is not None:
            m.predict_zero_shot = pzs
