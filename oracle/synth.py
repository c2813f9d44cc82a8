"""Seeded synthetic inputs shared by the oracle, the golden-vector generator, the tests and bench.py.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Everything is drawn from
``numpy.random.RandomState`` (MT19937; numpy guarantees the stream is stable across versions), so the
golden fixtures under tests/golden/ only need to store OUTPUTS: inputs are regenerated bit-identically
from the seed on any machine.

Shapes/densities follow SURVEY.md section 8(d): label density Bernoulli(0.0524) is the Indiana
chest-X-ray dataset density (3229 positives / 3851x16), tau=0.07 is 0426/config.py:26.
"""
from __future__ import annotations

import numpy as np
import torch

LABEL_DENSITY = 0.0524


def randn(seed: int, *shape: int) -> torch.Tensor:
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def uniform(seed: int, lo: float, hi: float, *shape: int) -> torch.Tensor:
    return torch.from_numpy(np.random.RandomState(seed).uniform(lo, hi, shape).astype(np.float32))


def labels(seed: int, rows: int, classes: int, density: float = LABEL_DENSITY) -> torch.Tensor:
    u = np.random.RandomState(seed).uniform(0.0, 1.0, (rows, classes))
    return torch.from_numpy((u < density).astype(np.float32))


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and come back to fp32: the oracle is evaluated on the SAME bf16-rounded inputs
    the CUDA path consumes (parity tolerances)."""
    return x.to(torch.bfloat16).to(torch.float32)


def projection_params(seed: int, in_features: int, shared: int) -> dict:
    """nn.Linear-like init (uniform +-1/sqrt(fan_in)), LayerNorm gamma/beta perturbed away from (1,0)
    so that parity tests exercise them.  Names mirror the reference state_dict
    (0426/train.py:78-82): first linear, fc, layer_norm."""
    k1 = 1.0 / np.sqrt(in_features)
    k2 = 1.0 / np.sqrt(shared)
    return {
        "w1": uniform(seed + 1, -k1, k1, shared, in_features),
        "b1": uniform(seed + 2, -k1, k1, shared),
        "w2": uniform(seed + 3, -k2, k2, shared, shared),
        "b2": uniform(seed + 4, -k2, k2, shared),
        "gamma": 1.0 + 0.1 * randn(seed + 5, shared),
        "beta": 0.1 * randn(seed + 6, shared),
    }


def unit_rows(seed: int, rows: int, cols: int) -> torch.Tensor:
    x = randn(seed, rows, cols)
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
