"""Drop-in installation: patch the names an imported reference module (`train`, `disease_analysis`) looks up as module
globals (SURVEY.md section 8b "Who calls it").  Nothing in the reference is edited; see INTEGRATION.md."""
from __future__ import annotations

import types


def install(*modules: types.ModuleType) -> None:
    from . import losses, modules as mods, zero_shot

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
        "predict_zero_shot": zero_shot.predict_zero_shot,
    }
    for m in modules:
        for name, obj in table.items():
            if hasattr(m, name):
                setattr(m, name, obj)
