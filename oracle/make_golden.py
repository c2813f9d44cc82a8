"""Generate tests/golden/head_golden.npz by running the UNMODIFIED reference on seeded inputs.

Run in the build container (where /root/reference exists):   python oracle/make_golden.py
The GPU box has no /root/reference; tests only read the committed .npz.

Only OUTPUTS are stored; inputs are regenerated from oracle/synth.py seeds (numpy RandomState streams).
Reference modules imported (nothing is copied):
  /root/reference/0426/train.py               ImageProjection, TextProjection, MultiViewFusion,
                                              contrastive_loss, contrastive_clip_loss_function,
                                              multilabel_contrastive_loss, predict_multilabel
  /root/reference/multimodal_attention/train.py   multilabel_asymmetric_loss, MultiModalAttention
The zero-shot drivers (0426/disease_analysis.py:291-364 etc.) need pretrained encoders + a tokenizer, so their
scoring cores are reproduced with the very same torch ops the reference calls (F.normalize, @, /tau,
F.softmax, topk, sigmoid, >=) -- see `zero_shot_*` below; each line cites the reference line it mirrors.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "..", "tests", "golden", "head_golden.npz")


def import_reference(variant: str):
    """Import <variant>/train.py from a scratch cwd (config.py mkdirs at import, 0426/config.py:96-98) with
    matplotlib/seaborn stubbed (absent in this image; multimodal_attention/train.py:34)."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    for m in ("train", "config", "prepare_data", "disease_analysis", "visualization"):
        sys.modules.pop(m, None)
    scratch = tempfile.mkdtemp(prefix="refcwd_")
    cwd = os.getcwd()
    os.chdir(scratch)
    sys.path.insert(0, os.path.join(REF, variant))
    try:
        mod = importlib.import_module("train")
    finally:
        sys.path.pop(0)
        os.chdir(cwd)
    return mod


def load_projection(module, first_name: str, p: dict):
    sd = {
        f"{first_name}.weight": p["w1"], f"{first_name}.bias": p["b1"],
        "fc.weight": p["w2"], "fc.bias": p["b2"],
        "layer_norm.weight": p["gamma"], "layer_norm.bias": p["beta"],
    }
    module.load_state_dict(sd)
    module.eval()          # dropout off: parity is defined with dropout off
    return module


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = import_reference("0426")
    out = {}

    # ---- a-P1 / a-P2: projection forward + input/param grads ---------------------------------
    for tag, cls, first, E in (("img", ref.ImageProjection, "image_projection", 96),
                               ("txt", ref.TextProjection, "text_projection", 80)):
        D, B = 64, 24
        p = synth.projection_params(100 if tag == "img" else 200, E, D)
        mod = load_projection(cls(E, D), first, p)
        x = synth.randn(7 if tag == "img" else 8, B, E).requires_grad_(True)
        y = mod(x)
        w = synth.randn(9, B, D)
        (y * w).sum().backward()
        out[f"proj_{tag}_y"] = y.detach().numpy()
        out[f"proj_{tag}_dx"] = x.grad.numpy()
        out[f"proj_{tag}_dw1"] = getattr(mod, first).weight.grad.numpy()
        out[f"proj_{tag}_db1"] = getattr(mod, first).bias.grad.numpy()
        out[f"proj_{tag}_dw2"] = mod.fc.weight.grad.numpy()
        out[f"proj_{tag}_db2"] = mod.fc.bias.grad.numpy()
        out[f"proj_{tag}_dgamma"] = mod.layer_norm.weight.grad.numpy()
        out[f"proj_{tag}_dbeta"] = mod.layer_norm.bias.grad.numpy()
    # 4-D input is flattened (0426/train.py:86-88)
    p = synth.projection_params(100, 96, 64)
    mod = load_projection(ref.ImageProjection(96, 64), "image_projection", p)
    out["proj_img_y_4d"] = mod(synth.randn(7, 24, 96).reshape(24, 6, 4, 4)).detach().numpy()

    # ---- a-N: contrastive_loss, tau 0.07 and 1.0, normalised inputs, + grads --------------------
    for tau in (0.07, 1.0):
        I = synth.unit_rows(11, 48, 64).requires_grad_(True)
        T = synth.unit_rows(12, 48, 64).requires_grad_(True)
        loss = ref.contrastive_loss(I, T, tau)
        loss.backward()
        out[f"nce_loss_tau{tau}"] = loss.detach().numpy()
        out[f"nce_dI_tau{tau}"] = I.grad.numpy()
        out[f"nce_dT_tau{tau}"] = T.grad.numpy()
    # un-normalised inputs (general path)
    I = synth.randn(13, 20, 32)
    T = synth.randn(14, 20, 32)
    out["nce_loss_unnorm"] = ref.contrastive_loss(I, T, 1.0).numpy()

    # ---- a-S: soft-target CLIP loss (train + eval modes) --------------------------------------
    Tt = synth.randn(15, 16, 32).requires_grad_(True)
    Ii = synth.randn(16, 16, 32).requires_grad_(True)
    loss = ref.contrastive_clip_loss_function(Tt, Ii, temperature=2.0, mode="train")
    loss.backward()
    out["soft_loss"] = loss.detach().numpy()
    out["soft_dT"] = Tt.grad.numpy()
    out["soft_dI"] = Ii.grad.numpy()
    out["soft_logits_eval"] = ref.contrastive_clip_loss_function(Tt.detach(), Ii.detach(), temperature=2.0,
                                                                 mode="eval").numpy()

    # ---- a-B: multilabel_contrastive_loss (tau 1.0 as train_epoch :434, and 0.07 as 0425 validate) ---
    for tau in (1.0, 0.07):
        I = synth.randn(21, 40, 64).requires_grad_(True)          # un-normalised: the fn normalises
        T = synth.randn(22, 16, 64)
        y = synth.labels(23, 40, 16, density=0.2)
        loss = ref.multilabel_contrastive_loss(I, T, y, tau)
        loss.backward()
        out[f"mlbce_loss_tau{tau}"] = loss.detach().numpy()
        out[f"mlbce_dI_tau{tau}"] = I.grad.numpy()
    # label padding branch (:205-210): labels narrower than C
    y_narrow = synth.labels(24, 40, 12, density=0.2)
    out["mlbce_loss_padded"] = ref.multilabel_contrastive_loss(synth.randn(21, 40, 64), synth.randn(22, 16, 64),
                                                               y_narrow, 1.0).numpy()
    # all-zero labels: pos term is 0/(0+1e-8)
    out["mlbce_loss_nolabels"] = ref.multilabel_contrastive_loss(synth.randn(21, 40, 64), synth.randn(22, 16, 64),
                                                                 torch.zeros(40, 16), 1.0).numpy()

    # ---- a-A: FC adapter + BCEWithLogits (NB02 c28:50-52) --------------------------------------
    fc = torch.nn.Linear(64, 16)
    fc.load_state_dict({"weight": synth.uniform(31, -0.125, 0.125, 16, 64), "bias": synth.uniform(32, -0.125, 0.125, 16)})
    x = synth.randn(33, 40, 64).requires_grad_(True)
    y = synth.labels(34, 40, 16, density=0.2)
    loss = torch.nn.BCEWithLogitsLoss()(fc(x), y)
    loss.backward()
    out["fc_loss"] = loss.detach().numpy()
    out["fc_dx"] = x.grad.numpy()
    out["fc_dw"] = fc.weight.grad.numpy()
    out["fc_db"] = fc.bias.grad.numpy()
    out["fc_pred"] = (torch.sigmoid(fc(x)) > 0.5).float().detach().numpy()        # NB02 c30:42-43

    # ---- a-M: predict_multilabel (tau bound from MODEL_CONFIG = 0.07) ---------------------------
    I = synth.randn(41, 40, 64)
    T = synth.unit_rows(42, 16, 64)
    out["predict_multilabel"] = ref.predict_multilabel(I, T, threshold=0.5).numpy()
    out["predict_multilabel_thr0.7"] = ref.predict_multilabel(I, T, threshold=0.7).numpy()

    # ---- a-Z: zero-shot scoring cores ------------------------------------------------------------
    X = synth.randn(51, 200, 64)
    T16 = synth.unit_rows(52, 16, 64)
    feats = F.normalize(X, dim=-1)                                   # 0426/disease_analysis.py:332
    sims = (feats @ T16.T) / 0.07                                    # :343
    probs = F.softmax(sims, dim=-1)                                  # :344
    vals, idx = probs.topk(3)                                        # :351 (row loop collapsed)
    out["z1_idx"] = idx.numpy()
    out["z1_vals"] = vals.numpy()
    sims2 = (feats @ T16.T) / 0.5                                    # multimodal_attention/disease_analysis.py:350-353
    pr2 = torch.sigmoid(sims2)                                       # :363
    out["z2_mask"] = (pr2 >= 0.5).numpy()                            # :378
    thr_vec = torch.linspace(0.45, 0.6, 16)
    out["z2_mask_perlabel"] = (pr2 >= thr_vec).numpy()               # :372 (dict thresholds)
    out["z2_argmax"] = pr2.argmax(-1).numpy()                        # :388 top-1 fallback
    cos = feats @ T16.T                                              # NB02 c41:27-32
    out["z3_argmax"] = cos.argmax(-1).numpy()
    out["z3_mask"] = (torch.sigmoid(cos) > 0.5).numpy()              # NB02 c44:24-36
    # north-star 14 x (pos,neg) shape, composed from the same reference ops
    P = synth.unit_rows(53, 28, 64).reshape(14, 2, 64)
    lg = (feats @ P.reshape(28, 64).T) / 0.07
    q = F.softmax(lg.reshape(-1, 14, 2), dim=-1)[..., 0]
    out["zn_argmax"] = q.argmax(-1).numpy()
    out["zn_mask"] = (q > 0.5).numpy()
    out["zn_q"] = q.numpy()

    # ---- a-F: MultiViewFusion (shared size is fixed to 512 by MODEL_CONFIG) -----------------------
    fus = ref.MultiViewFusion().eval()
    fp = {"w0": synth.uniform(61, -0.03, 0.03, 512, 1024), "b0": synth.uniform(62, -0.03, 0.03, 512),
          "w3": synth.uniform(63, -0.04, 0.04, 512, 512), "b3": synth.uniform(64, -0.04, 0.04, 512)}
    fus.load_state_dict({"fusion.0.weight": fp["w0"], "fusion.0.bias": fp["b0"],
                         "fusion.3.weight": fp["w3"], "fusion.3.bias": fp["b3"]})
    out["fusion_y"] = fus(synth.randn(65, 6, 512), synth.randn(66, 6, 512)).detach().numpy()

    # ---- a-X: ASL + MultiModalAttention (multimodal_attention variant) ---------------------------
    refm = import_reference("multimodal_attention")
    lg = synth.randn(71, 40, 16) * 3
    y = synth.labels(72, 40, 16, density=0.2)
    out["asl_mean"] = refm.multilabel_asymmetric_loss(lg, y).numpy()
    out["asl_sum_g1"] = refm.multilabel_asymmetric_loss(lg, y, gamma_pos=1, gamma_neg=2, clip=0.1,
                                                        reduction="sum").numpy()
    att = refm.MultiModalAttention().eval()
    ap = {"wi": synth.uniform(81, -0.04, 0.04, 512, 512), "bi": synth.uniform(82, -0.04, 0.04, 512),
          "wt": synth.uniform(83, -0.04, 0.04, 512, 512), "bt": synth.uniform(84, -0.04, 0.04, 512),
          "wa": synth.uniform(85, -0.04, 0.04, 1, 512), "ba": synth.uniform(86, -0.04, 0.04, 1),
          "wo": synth.uniform(87, -0.04, 0.04, 512, 512), "bo": synth.uniform(88, -0.04, 0.04, 512)}
    att.load_state_dict({"image_proj.weight": ap["wi"], "image_proj.bias": ap["bi"],
                         "text_proj.weight": ap["wt"], "text_proj.bias": ap["bt"],
                         "attention.weight": ap["wa"], "attention.bias": ap["ba"],
                         "output_proj.weight": ap["wo"], "output_proj.bias": ap["bo"]})
    enh, w = att(synth.randn(89, 6, 512), synth.unit_rows(90, 16, 512))
    out["attn_enh"] = enh.detach().numpy()
    out["attn_w"] = w.detach().numpy()

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
