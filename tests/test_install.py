"""Drop-in installation: install() rebinds the names the reference looks up as module globals; module
state_dict keys/shapes equal the reference's so its checkpoints load.  When /root/reference is present (build container)
the real reference module is patched and its own state_dict is loaded into the replacement."""
import os
import sys
import types

import pytest
import torch

import b200clip


def test_install_patches_module_globals():
    fake = types.ModuleType("train")
    for n in ("ImageProjection", "TextProjection", "MultiViewFusion", "contrastive_loss", "multilabel_contrastive_loss",
              "predict_multilabel"):
        setattr(fake, n, object())
    fake.unrelated = 1
    b200clip.install(fake)
    assert fake.ImageProjection is b200clip.ImageProjection and fake.TextProjection is b200clip.TextProjection
    assert fake.contrastive_loss is b200clip.contrastive_loss and fake.MultiViewFusion is b200clip.MultiViewFusion
    assert fake.multilabel_contrastive_loss is b200clip.multilabel_contrastive_loss
    assert fake.predict_multilabel is b200clip.predict_multilabel
    assert fake.unrelated == 1 and not hasattr(fake, "predict_zero_shot")


def test_signatures_match_reference_defaults():
    import inspect
    sig = inspect.signature(b200clip.contrastive_loss)
    # the reference's three parameters, in order, same default; one extra keyword with a default (inputs_normalized) is allowed
    assert list(sig.parameters)[:3] == ["image_features", "text_features", "temperature"] and sig.parameters["temperature"].default == 1.0
    assert all(p.default is not inspect.Parameter.empty for p in list(sig.parameters.values())[3:])
    sig = inspect.signature(b200clip.multilabel_contrastive_loss)
    assert list(sig.parameters)[:4] == ["image_features", "text_features", "labels", "temperature"]
    sig = inspect.signature(b200clip.contrastive_clip_loss_function)
    assert list(sig.parameters) == ["text_projection", "image_projection", "temperature", "mode"]
    assert sig.parameters["temperature"].default == 0.07 and sig.parameters["mode"].default == "eval"
    sig = inspect.signature(b200clip.multilabel_asymmetric_loss)
    assert list(sig.parameters) == ["logits", "targets", "gamma_pos", "gamma_neg", "clip", "eps", "reduction"]
    assert [sig.parameters[k].default for k in ("gamma_pos", "gamma_neg", "clip", "eps", "reduction")] == [0, 4, 0.05, 1e-8, "mean"]
    sig = inspect.signature(b200clip.predict_multilabel)
    assert list(sig.parameters) == ["image_features", "text_features", "threshold"] and sig.parameters["threshold"].default == 0.5
    sig = inspect.signature(b200clip.predict_zero_shot)
    assert list(sig.parameters)[:6] == ["images", "models", "disease_list", "top_k", "prompts", "use_enhanced_prompts"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/0426"), reason="reference only exists in the build container")
def test_reference_checkpoint_keys_load(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from make_golden import import_reference
    ref = import_reference("0426")
    ref_img, ref_txt = ref.ImageProjection(2048, 512), ref.TextProjection(768, 512)
    b200clip.install(ref)
    assert ref.ImageProjection is b200clip.ImageProjection
    mine_img, mine_txt = ref.ImageProjection(2048, 512), ref.TextProjection(768, 512)      # built through the patched globals
    assert {k: tuple(v.shape) for k, v in mine_img.state_dict().items()} == {k: tuple(v.shape) for k, v in ref_img.state_dict().items()}
    mine_img.load_state_dict(ref_img.state_dict())                                          # strict load of a reference checkpoint
    mine_txt.load_state_dict(ref_txt.state_dict())
    assert torch.equal(mine_img.fc.weight, ref_img.fc.weight)
    # MultiViewFusion: same no-argument constructor, same keys, strict load (0426/train.py:988-1000)
    assert ref.MultiViewFusion is b200clip.MultiViewFusion
    mine_fus = ref.MultiViewFusion()
    assert set(mine_fus.state_dict()) == {"fusion.0.weight", "fusion.0.bias", "fusion.3.weight", "fusion.3.bias"}
    assert tuple(mine_fus.state_dict()["fusion.0.weight"].shape) == (512, 1024)


def test_multimodal_attention_state_dict_keys():
    """multimodal_attention/train.py:1069-1080: four nn.Linear sub-modules under these names."""
    m = b200clip.MultiModalAttention()
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == {"image_proj.weight": (512, 512), "image_proj.bias": (512,), "text_proj.weight": (512, 512),
                  "text_proj.bias": (512,), "attention.weight": (1, 512), "attention.bias": (1,),
                  "output_proj.weight": (512, 512), "output_proj.bias": (512,)}
