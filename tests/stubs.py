"""Tiny random-init stand-ins for the encoders / tokenizer the reference's callers expect in their `models` dict
(`resnet`, `tokenizer`, `text_model`; 0426/train.py:888-928).  Used by oracle/make_golden_edges.py (to drive the UNMODIFIED
reference functions in the build container) and by the tests (to drive the b200clip drop-ins with the very same stubs).
Weights come from oracle/synth.py numpy streams, so they are identical on every machine."""
from __future__ import annotations

import zlib

import torch
import torch.nn as nn

import synth

E_IMG, E_TXT, D, VOCAB, SEQ = 128, 64, 512, 997, 8
DISEASES = ["Cardiomegaly", "Pulmonary Atelectasis", "Pleural Effusion", "Nodule", "Infiltrate", "Emphysema", "Thickening",
            "Hernia", "Pulmonary Edema", "Pneumonia", "Consolidation", "Pneumothorax", "Fibrosis", "Mass", "Granuloma", "Normal"]


class StubBatch(dict):
    """What a HuggingFace tokenizer returns, as far as the reference uses it: a mapping with .to(device)."""

    def to(self, device):
        return StubBatch({k: v.to(device) for k, v in self.items()})


class StubTokenizer:
    """Deterministic word-hash tokenizer: prompts -> input_ids [n, SEQ], attention_mask (padding='max_length')."""

    def __call__(self, prompts, return_tensors="pt", padding="max_length", max_length=128, truncation=True):
        if isinstance(prompts, str):
            prompts = [prompts]
        ids = torch.zeros((len(prompts), SEQ), dtype=torch.long)
        mask = torch.zeros((len(prompts), SEQ), dtype=torch.long)
        for i, text in enumerate(prompts):
            toks = [1 + zlib.crc32(w.encode()) % (VOCAB - 1) for w in text.lower().split()][:SEQ - 1]
            toks = [2 + zlib.crc32(text.encode()) % (VOCAB - 2)] + toks          # a "[CLS]" slot that depends on the whole prompt
            ids[i, :len(toks)] = torch.tensor(toks[:SEQ])
            mask[i, :len(toks)] = 1
        return StubBatch(input_ids=ids, attention_mask=mask)


class _Out:
    def __init__(self, h):
        self.last_hidden_state = h


class StubTextModel(nn.Module):
    """Embedding lookup + one mixing layer standing in for Bio_ClinicalBERT: forward(**inputs).last_hidden_state [n, SEQ, E_TXT]."""

    def __init__(self):
        super().__init__()
        self.emb = nn.Embedding(VOCAB, E_TXT)
        self.mix = nn.Linear(E_TXT, E_TXT)
        with torch.no_grad():
            self.emb.weight.copy_(synth.randn(301, VOCAB, E_TXT))
            self.mix.weight.copy_(synth.uniform(302, -0.2, 0.2, E_TXT, E_TXT))
            self.mix.bias.copy_(synth.uniform(303, -0.2, 0.2, E_TXT))

    def forward(self, input_ids=None, attention_mask=None, **_):
        h = self.emb(input_ids)
        m = attention_mask.unsqueeze(-1).to(h.dtype)
        pooled = (h * m).sum(1, keepdim=True) / m.sum(1, keepdim=True).clamp_min(1.0)
        return _Out(torch.tanh(self.mix(h + 0.25 * pooled)))


class StubResNet(nn.Module):
    """[N,3,H,W] -> [N, E_IMG, 1, 1] like torchvision resnet50 with fc = Identity (the callers flatten with .view(N, -1))."""

    def __init__(self):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.fc = nn.Linear(48, E_IMG)
        with torch.no_grad():
            self.fc.weight.copy_(synth.uniform(311, -0.5, 0.5, E_IMG, 48))
            self.fc.bias.copy_(synth.uniform(312, -0.5, 0.5, E_IMG))

    def forward(self, x):
        return self.fc(self.pool(x).flatten(1)).reshape(x.shape[0], E_IMG, 1, 1)


def images(seed: int, n: int) -> torch.Tensor:
    return synth.randn(seed, n, 3, 16, 16)


def load_projection(module, first_name: str, p: dict):
    module.load_state_dict({f"{first_name}.weight": p["w1"], f"{first_name}.bias": p["b1"], "fc.weight": p["w2"], "fc.bias": p["b2"],
                            "layer_norm.weight": p["gamma"], "layer_norm.bias": p["beta"]})
    return module.eval()


def build_models(image_projection_cls, text_projection_cls, device="cpu", attention_cls=None):
    """The `models` dict of 0426/train.py:888-928 with stub encoders and the given projector classes (the reference's or
    b200clip's -- same constructor signature, same state_dict keys)."""
    models = {
        "resnet": StubResNet().to(device).eval(),
        "tokenizer": StubTokenizer(),
        "text_model": StubTextModel().to(device).eval(),
        "image_projector": load_projection(image_projection_cls(E_IMG, D), "image_projection", synth.projection_params(321, E_IMG, D)).to(device),
        "text_projector": load_projection(text_projection_cls(E_TXT, D), "text_projection", synth.projection_params(322, E_TXT, D)).to(device),
    }
    if attention_cls is not None:
        att = attention_cls()
        att.load_state_dict({"image_proj.weight": synth.uniform(331, -0.04, 0.04, D, D), "image_proj.bias": synth.uniform(332, -0.04, 0.04, D),
                             "text_proj.weight": synth.uniform(333, -0.04, 0.04, D, D), "text_proj.bias": synth.uniform(334, -0.04, 0.04, D),
                             "attention.weight": synth.uniform(335, -0.04, 0.04, 1, D), "attention.bias": synth.uniform(336, -0.04, 0.04, 1),
                             "output_proj.weight": synth.uniform(337, -0.04, 0.04, D, D), "output_proj.bias": synth.uniform(338, -0.04, 0.04, D)})
        models["multimodal_attention"] = att.to(device).eval()
    return models


def z2_cases():
    """keyword arguments of the multimodal predict_zero_shot calls pinned in tests/golden/edges_golden.npz"""
    per = {d: 0.5 + 0.01 * (i % 5) for i, d in enumerate(DISEASES) if i % 4 != 3}
    return {"thr05": dict(threshold=0.5), "thr_hi_top2": dict(threshold=0.62, top_k=2), "thr_lo_top3": dict(threshold=0.45, top_k=3),
            "dict": dict(threshold=dict(per)), "dict_top2": dict(threshold=dict(per), top_k=2)}


def unpad_lists(idx, val):
    """inverse of make_golden_edges.pad_lists: padded arrays -> ragged (index lists, score lists)"""
    names, scores = [], []
    for r, v in zip(idx, val):
        k = int((r >= 0).sum())
        names.append([int(j) for j in r[:k]])
        scores.append([float(x) for x in v[:k]])
    return names, scores


def prediction_text_features(diseases, tokenizer, text_model, text_projector, device="cpu"):
    """What get_prediction_text_features (0426/disease_analysis.py:152-198) computes, for the GPU box where the reference is
    absent: one fixed prompt per disease ("Normal" has its own sentence, :176-179), CLS token -> text projector -> F.normalize.
    tests/test_integration_reference.py checks the b200clip drop-in against the reference's OWN function; this restatement is
    pinned by tests/golden/edges_golden.npz['text'] (test_gpu_edges.py / test_oracle_edges.py)."""
    prompts = ["This is a normal chest X-ray without any significant findings." if d == "Normal" else f"This chest X-ray shows {d}."
               for d in diseases]
    inputs = tokenizer(prompts, return_tensors="pt", padding="max_length", max_length=128, truncation=True).to(device)
    with torch.no_grad():
        cls = text_model(**inputs).last_hidden_state[:, 0, :]
        return torch.nn.functional.normalize(text_projector(cls), dim=-1)
