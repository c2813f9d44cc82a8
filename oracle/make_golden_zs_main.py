"""Generate tests/golden/zs_main_golden.npz by running the UNMODIFIED `main()` of the reference's
multimodal_attention/zero_shot_predict.py (:14-261) -- the dynamic per-label threshold search (:66-159) and the weighted
two-view merge (:161-224) live inside that function between data loaders and a checkpoint load, so they are pinned by
executing the whole function with the DATA side stubbed:

  * `load_data` (prepare_data.py; needs the MIMIC-CXR files)            -> a list of synthetic (images, labels, findings,
    view_types) batches, images [bs, 2, 3, 224, 224] as the real loader yields them (two views per study);
  * `initialize_models` (train.py; downloads ResNet-50 / Bio_ClinicalBERT) -> tests/stubs.py encoders + the reference's OWN
    ImageProjection / TextProjection;
  * `analyze_disease_distribution` / `create_rich_prompts` (data statistics of the report table) -> prompts = None, i.e. the
    stock get_prediction_text_features text side; `visualize_predictions` (matplotlib) -> no-op;
  * the checkpoint file is an empty {'models': {}} written to a scratch directory.

Everything on the scoring path runs unmodified: the reference's predict_zero_shot, the threshold search with scikit-learn's
f1_score, the merge, the prediction matrix, evaluate_predictions.  Observation points (wrappers that only record and forward):
predict_zero_shot's `threshold` argument of the second pass IS the thresholds dict (full precision), and
evaluate_predictions receives the final prediction matrix.  For every call the wrapper also asks the reference function for
the unfiltered per-view probabilities (threshold=0.0, as the first pass does), which become the fixture's inputs.

Run in the build container (where /root/reference exists):   python oracle/make_golden_zs_main.py
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import stubs  # noqa: E402
import synth  # noqa: E402
from make_golden import REF, import_reference  # noqa: E402
from make_golden_edges import import_disease_analysis  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "zs_main_golden.npz")
VARIANT = "multimodal_attention"
N_BATCHES, BS = 12, 8


def study_images(seed: int, n: int) -> torch.Tensor:
    """[n, 2, 3, 224, 224]: 16x16 noise blown up 14x so the stub encoder's 4x4 average pooling keeps per-image variance"""
    low = synth.randn(seed, n * 2, 3, 16, 16)
    return low.repeat_interleave(14, dim=-1).repeat_interleave(14, dim=-2).reshape(n, 2, 3, 224, 224).contiguous()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    logging.disable(logging.CRITICAL)
    dl = list(stubs.DISEASES)
    L = len(dl)
    refm = import_reference(VARIANT)
    dam = import_disease_analysis(VARIANT)
    models = stubs.build_models(refm.ImageProjection, refm.TextProjection)

    # synthetic validation set; labels follow the scores (so the F1 search has something to find) with noise, plus the two
    # degenerate columns the reference special-cases: a label with no positives and one with no negatives in the 25 % slice
    n = N_BATCHES * BS
    images = study_images(701, n)
    names, scores = dam.predict_zero_shot(images.view(-1, 3, 224, 224), models, dl, threshold=0.0)
    prob = np.array(scores, dtype=np.float64).reshape(n, 2, L)
    mx = prob.max(axis=1)
    noise = synth.randn(702, n, L).numpy() * mx.std(axis=0, keepdims=True) * 0.7
    labels = ((mx + noise) > np.quantile(mx, 0.65, axis=0, keepdims=True)).astype(np.float32)
    n_thr = (N_BATCHES // 4) * BS
    labels[:n_thr, 3] = 0.0                     # no positive sample in the threshold slice -> 0.8   (:121-124)
    labels[:n_thr, 9] = 1.0                     # no negative sample                         -> 0.2   (:127-130)
    loader = [(images[b * BS:(b + 1) * BS], torch.from_numpy(labels[b * BS:(b + 1) * BS]), [""] * BS, [("PA", "LATERAL")] * BS)
              for b in range(N_BATCHES)]

    # ---- import the script module next to its own siblings, data-side modules stubbed -------------------------------
    import types
    prep = types.ModuleType("prepare_data")
    prep.load_data = lambda: (None, loader, dl, None)
    vis = types.ModuleType("visualization")
    vis.visualize_predictions = lambda *a, **k: None
    sys.modules["prepare_data"], sys.modules["visualization"] = prep, vis
    sys.modules.pop("zero_shot_predict", None)
    scratch = tempfile.mkdtemp(prefix="refzs_")
    cwd = os.getcwd()
    os.chdir(scratch)
    sys.path.insert(0, os.path.join(REF, VARIANT))
    try:
        zsp = importlib.import_module("zero_shot_predict")
    finally:
        sys.path.pop(0)
    try:
        zsp.LOG_CONFIG = {"log_dir": os.path.join(scratch, "logs"), "checkpoint_dir": os.path.join(scratch, "ck")}
        os.makedirs(zsp.LOG_CONFIG["checkpoint_dir"], exist_ok=True)
        torch.save({"models": {}}, os.path.join(zsp.LOG_CONFIG["checkpoint_dir"], "model_best.pth"))
        zsp.initialize_models = lambda device: models
        zsp.analyze_disease_distribution = lambda df: {}
        zsp.create_rich_prompts = lambda stats: None

        seen = {"thr_args": [], "probs": [], "pred_matrix": None, "true": None}
        real_predict, real_eval = zsp.predict_zero_shot, zsp.evaluate_predictions

        def predict_recorder(imgs, mdl, dlist, threshold=0.5, top_k=None, prompts=None, use_enhanced_prompts=False):
            seen["thr_args"].append(threshold)
            if isinstance(threshold, dict):                                    # second pass: also record the raw probabilities
                _, sc = real_predict(imgs, mdl, dlist, threshold=0.0, prompts=prompts, use_enhanced_prompts=use_enhanced_prompts)
                seen["probs"].append(np.array(sc, dtype=np.float64))
            return real_predict(imgs, mdl, dlist, threshold=threshold, top_k=top_k, prompts=prompts,
                                use_enhanced_prompts=use_enhanced_prompts)

        def eval_recorder(predictions, true_labels, dlist):
            seen["pred_matrix"], seen["true"] = np.array(predictions), np.array(true_labels)
            return real_eval(predictions, true_labels, dlist)

        zsp.predict_zero_shot, zsp.evaluate_predictions = predict_recorder, eval_recorder
        zsp.main()                                                             # the unmodified function, end to end
    finally:
        os.chdir(cwd)

    first = [t for t in seen["thr_args"] if not isinstance(t, dict)]
    second = [t for t in seen["thr_args"] if isinstance(t, dict)]
    assert len(first) == N_BATCHES // 4 and all(t == 0.0 for t in first) and len(second) == N_BATCHES
    thr = np.array([second[0][d] for d in dl], dtype=np.float64)
    prob2 = np.concatenate(seen["probs"]).reshape(n, 2, L)
    assert np.array_equal(prob2, prob)                                         # eval mode: the two passes see the same scores
    assert np.array_equal(seen["true"], labels)
    p32 = prob.astype(np.float32)
    assert np.array_equal(p32.astype(np.float64), prob)                        # python floats of float32 tensor elements
    out = {"prob_views": p32, "labels": labels, "n_threshold_studies": np.array(n_thr), "thresholds": thr,
           "pred_matrix": seen["pred_matrix"].astype(np.float64)}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    kinds = {"0.8": int((thr == 0.8).sum()), "0.2": int((thr == 0.2).sum()), "0.5": int((thr == 0.5).sum())}
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.1f} KiB; thresholds {np.round(thr, 3).tolist()} special {kinds}; "
          f"positives per study {seen['pred_matrix'].sum(1).mean():.2f}")


if __name__ == "__main__":
    main()
