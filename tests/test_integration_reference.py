"""SURVEY.md section 4, "integration" level: the reference's UNMODIFIED callers run through b200clip.install().

Only possible where the reference exists (the build container, which has no GPU): the C-ABI layer is replaced by oracle
arithmetic (tests/cpu_doubles.py) and everything above it -- the patched names, constructor signatures, state_dict keys,
argument conventions, return types, autograd connectivity -- is exercised by the reference's own train_epoch / validate /
predict_zero_shot call sites.  tests/test_gpu_edges.py drives the same stubs through the real kernels on the B200."""
import logging
import os
import sys

import numpy as np
import pytest
import torch

import b200clip
import stubs
from cpu_doubles import cpu_ops

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/0426"), reason="reference only exists in the build container")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _loader(n_batches, bs, seed):
    import synth
    out = []
    for b in range(n_batches):
        imgs = synth.randn(seed + b, bs, 2, 3, 16, 16)
        labels = synth.labels(seed + 100 + b, bs, 16, density=0.2)
        out.append((imgs, labels, ["no acute findings"] * bs, [["frontal", "lateral"]] * bs))
    return out


def test_unmodified_train_epoch_and_validate_run_through_install(caplog):
    from make_golden import import_reference
    train = import_reference("0426")
    b200clip.install(train)
    assert train.ImageProjection is b200clip.ImageProjection and train.multilabel_contrastive_loss is b200clip.multilabel_contrastive_loss
    with cpu_ops():
        # the `models` dict of train.initialize_models (0426/train.py:888-928), built through the PATCHED module globals
        models = stubs.build_models(train.ImageProjection, train.TextProjection)
        models["view_fusion"] = train.MultiViewFusion()
        params = [p for k in ("image_projector", "text_projector", "view_fusion") for p in models[k].parameters()]
        opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01)
        before = [p.detach().clone() for p in params]
        caplog.set_level(logging.ERROR)
        loss_avg, acc_avg = train.train_epoch(models, _loader(2, 8, 500), opt, 0, stubs.DISEASES)     # unmodified, 0426/train.py:304-497
        assert not [r for r in caplog.records if r.levelno >= logging.ERROR], [r.getMessage() for r in caplog.records]
        assert np.isfinite(loss_avg) and loss_avg > 0 and 0.0 <= acc_avg <= 100.0
        changed = [not torch.equal(a, b.detach()) for a, b in zip(before, params)]
        # the text projector only runs under no_grad in train_epoch (:333); image projector + fusion must have been stepped
        n_img = len(list(models["image_projector"].parameters()))
        assert all(changed[:n_img]) and all(changed[-4:])
        val = train.validate(models, _loader(2, 8, 600), stubs.DISEASES)                              # unmodified, :499-620
        assert not [r for r in caplog.records if r.levelno >= logging.ERROR]
        assert np.isfinite(val[0])


@pytest.mark.parametrize("variant", ["0426", "multimodal_attention"])
def test_unmodified_predict_zero_shot_callers(variant):
    """predict_zero_shot called exactly as 0426/zero_shot_predict.py:71-78 and multimodal_attention/zero_shot_predict.py do:
    through the patched disease_analysis module, with the reference's `models` dict (no 'text_features' entry)."""
    from make_golden import import_reference
    from make_golden_edges import import_disease_analysis
    ref = import_reference(variant)
    da = import_disease_analysis(variant)
    ref_fn = da.predict_zero_shot
    edges = dict(np.load(os.path.join(ROOT, "tests", "golden", "edges_golden.npz")))
    imgs = stubs.images(401, 48)
    b200clip.install(ref, da)
    assert da.predict_zero_shot is not ref_fn
    with cpu_ops():
        models = stubs.build_models(ref.ImageProjection, ref.TextProjection)
        if variant == "0426":
            names, scores = da.predict_zero_shot(imgs, models, stubs.DISEASES, top_k=3, prompts=None, use_enhanced_prompts=True)
            assert [[stubs.DISEASES.index(n) for n in row] for row in names] == edges["z1_idx"].tolist()
            np.testing.assert_allclose(np.array(scores), edges["z1_val"], rtol=1e-4)
            single = da.predict_zero_shot(imgs[5], models, stubs.DISEASES)
            assert [stubs.DISEASES.index(d["disease"]) for d in single] == edges["z1_single_idx"].tolist()
        else:
            for tag, kw in stubs.z2_cases().items():
                names, scores = da.predict_zero_shot(imgs, models, stubs.DISEASES, **kw)
                ref_idx, ref_val = stubs.unpad_lists(edges[f"z2_{tag}_idx"], edges[f"z2_{tag}_val"])
                assert [[stubs.DISEASES.index(n) for n in row] for row in names] == ref_idx, tag
                for a, b in zip(scores, ref_val):
                    np.testing.assert_allclose(a, b, rtol=1e-4)
            single = da.predict_zero_shot(imgs[7], models, stubs.DISEASES, 0.5, 2)          # positional (threshold, top_k)
            assert [stubs.DISEASES.index(d["disease"]) for d in single] == edges["z2_single_idx"].tolist()
