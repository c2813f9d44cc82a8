"""Generate tests/golden/edges_golden.npz by running the UNMODIFIED reference on seeded inputs (round 2 fixtures).

Run in the build container (where /root/reference exists):   python oracle/make_golden_edges.py
What is pinned here (all of it produced by the reference's own functions, nothing restated):
  * predict_zero_shot of 0426/disease_analysis.py:291-364 and of multimodal_attention/disease_analysis.py:291-421, called
    exactly as their callers do (0426/zero_shot_predict.py:71-78) with a `models` dict of stub encoders (tests/stubs.py) and
    the reference's own ImageProjection / TextProjection / MultiModalAttention -> the returned python lists;
  * calculate_multilabel_metrics (0426/train.py:251-302);
  * get_text_features_with_findings (0426/disease_analysis.py:449-497) -> the pooled prompt features.
The projector outputs / text features the lists were computed from are stored too, so the oracle's scoring cores can be
checked against the lists on the CPU and the CUDA kernels can be fed identical inputs.
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
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import stubs  # noqa: E402
import synth  # noqa: E402
from make_golden import REF, import_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "edges_golden.npz")


def import_disease_analysis(variant: str):
    """<variant>/disease_analysis.py next to its own train/config modules (same scratch-cwd / stub rules as import_reference)."""
    import_reference(variant)                       # leaves config etc. of this variant in sys.modules
    sys.modules.pop("disease_analysis", None)
    scratch = tempfile.mkdtemp(prefix="refcwd_")
    cwd = os.getcwd()
    os.chdir(scratch)
    sys.path.insert(0, os.path.join(REF, variant))
    try:
        return importlib.import_module("disease_analysis")
    finally:
        sys.path.pop(0)
        os.chdir(cwd)


def pad_lists(names, scores, disease_list, width):
    """ragged python lists -> (idx [N,width] int64 padded with -1, val [N,width] f32 padded with 0)"""
    idx = -np.ones((len(names), width), dtype=np.int64)
    val = np.zeros((len(names), width), dtype=np.float32)
    for i, (ns, ss) in enumerate(zip(names, scores)):
        for j, (n, s) in enumerate(zip(ns, ss)):
            idx[i, j] = disease_list.index(n)
            val[i, j] = s
    return idx, val


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    out = {}
    dl = stubs.DISEASES
    imgs = stubs.images(401, 48)

    # ---- 0426: predict_zero_shot (softmax top-k) ----------------------------------------------------------------
    ref = import_reference("0426")
    da = import_disease_analysis("0426")
    models = stubs.build_models(ref.ImageProjection, ref.TextProjection)
    with torch.no_grad():
        emb = models["resnet"](imgs)
        out["feats"] = models["image_projector"](emb.view(emb.size(0), -1)).numpy()          # projector output (un-normalised)
        out["text"] = da.get_prediction_text_features(dl, models["tokenizer"], models["text_model"], models["text_projector"]).numpy()
    names, scores = da.predict_zero_shot(imgs, models, dl, top_k=3, prompts=None, use_enhanced_prompts=True)   # 0426/zero_shot_predict.py:71-78
    out["z1_idx"], out["z1_val"] = pad_lists(names, [list(s) for s in scores], dl, 3)
    single = da.predict_zero_shot(imgs[5], models, dl)                                       # single-image branch (:357-364)
    out["z1_single_idx"] = np.array([dl.index(d["disease"]) for d in single])
    out["z1_single_val"] = np.array([d["confidence"] for d in single], dtype=np.float32)

    # ---- 0426: get_text_features_with_findings pooling (:449-497), ragged prompt counts ----------------------------------
    prompts = {d: [f"Chest film with {d}.", f"Findings suggest {d}.", f"{d} is present."][:1 + (i % 3)] for i, d in enumerate(dl[:10])}
    pooled = da.get_text_features_with_findings(dl[:12], models["tokenizer"], models["text_model"], models["text_projector"], prompts, "cpu")
    out["pool_out"] = pooled.numpy()
    rows, counts = [], []
    with torch.no_grad():
        for d in dl[:12]:
            ps = prompts.get(d, [f"This is a chest X-ray showing {d}."])                      # :474
            inp = models["tokenizer"](ps)
            rows.append(models["text_projector"](models["text_model"](**inp).last_hidden_state[:, 0, :]))
            counts.append(len(ps))
    out["pool_in"] = torch.cat(rows).numpy()
    out["pool_counts"] = np.array(counts)

    # ---- 0426: calculate_multilabel_metrics (:251-302) -------------------------------------------------------------
    for tag, (pseed, lseed, n, dens) in {"a": (411, 412, 200, 0.2), "b": (413, 414, 37, 0.0524)}.items():
        pred = torch.sigmoid(synth.randn(pseed, n, 16) * 2.0)
        lab = synth.labels(lseed, n, 16, density=dens)
        m = ref.calculate_multilabel_metrics(pred, lab)
        out[f"metrics_{tag}"] = np.array([m[k] for k in ("sample_acc", "label_acc", "hamming_score", "exact_match", "top1_acc",
                                                          "top3_acc", "f1_score")], dtype=np.float64)
        hard = (pred > 0.5).float()
        out[f"metrics_{tag}_class_acc"] = ((hard == lab).float().mean(dim=0) * 100).numpy()    # 0426/train.py:445-447 on the same matrix

    # ---- multimodal_attention: predict_zero_shot (sigmoid / thresholds / top-k) -------------------------------------
    refm = import_reference("multimodal_attention")
    dam = import_disease_analysis("multimodal_attention")
    mm = stubs.build_models(refm.ImageProjection, refm.TextProjection)
    cases = stubs.z2_cases()
    for tag, kw in cases.items():
        names, scores = dam.predict_zero_shot(imgs, mm, dl, **kw)
        out[f"z2_{tag}_idx"], out[f"z2_{tag}_val"] = pad_lists(names, scores, dl, 16)
    single = dam.predict_zero_shot(imgs[7], mm, dl, 0.5, 2)                                    # positional (threshold, top_k), :291-299
    out["z2_single_idx"] = np.array([dl.index(d["disease"]) for d in single])
    out["z2_single_val"] = np.array([d["confidence"] for d in single], dtype=np.float32)
    # with the attention module in the dict (:344-347)
    mma = stubs.build_models(refm.ImageProjection, refm.TextProjection, attention_cls=refm.MultiModalAttention)
    with torch.no_grad():
        tf = dam.get_prediction_text_features(dl, mma["tokenizer"], mma["text_model"], mma["text_projector"])
        enh, _ = mma["multimodal_attention"](F.normalize(torch.from_numpy(out["feats"]), dim=-1), tf)
        out["feats_attn"] = enh.numpy()
    names, scores = dam.predict_zero_shot(imgs, mma, dl, threshold=0.5, top_k=2)
    out["z2_attn_idx"], out["z2_attn_val"] = pad_lists(names, scores, dl, 16)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
