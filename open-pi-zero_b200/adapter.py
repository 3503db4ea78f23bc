"""Caller-side normalisation of the reference's environment adapters: `BaseEnvAdapter`'s four formulas
(`src/agent/env_adapter/base.py:8-49`, same names and arguments, numpy) and their folding into the device path --
`SimplerAdapter.preprocess` normalises the raw proprio (`simpler.py:76-90`), `postprocess` de-normalises every action
dimension except the gripper (`simpler.py:102-125`).  Both are per-dimension affine maps, so `PiZero.set_io_normalization`
hands them to the kernels as (scale, shift) vectors: the proprio map runs inside the kernel that first reads the proprio,
the action map after the sampler's final clip.  Tokenisation, image resizing and the simulator-specific pose conversions
stay on the host (tokenizer files / simulator absent here)."""
from __future__ import annotations

import numpy as np


class BaseEnvAdapter:
    """base.py:4-49, verbatim semantics."""

    def normalize_bound(self, data, data_min, data_max, clip_min: float = -1, clip_max: float = 1, eps: float = 1e-8):
        ndata = 2 * (data - data_min) / (data_max - data_min + eps) - 1
        return np.clip(ndata, clip_min, clip_max)

    def denormalize_bound(self, data, data_min, data_max, clip_min: float = -1, clip_max: float = 1, eps=1e-8):
        clip_range = clip_max - clip_min
        return (data - clip_min) / clip_range * (data_max - data_min) + data_min

    def normalize_gaussian(self, data, mean, std, eps: float = 1e-8):
        return (data - mean) / (std + eps)

    def denormalize_gaussian(self, data, mean, std, eps: float = 1e-8):
        return data * (std + eps) + mean


def proprio_affine(stats: dict, kind: str, eps: float = 1e-8):
    """(scale, shift, clip) with normalise(x) = clip(x * scale + shift): `bound` uses p01 / p99, `gaussian` mean / std."""
    if kind == "bound":
        lo, hi = np.asarray(stats["p01"], np.float64), np.asarray(stats["p99"], np.float64)
        scale = 2.0 / (hi - lo + eps)
        return scale, -lo * scale - 1.0, True
    if kind == "gaussian":
        mean, std = np.asarray(stats["mean"], np.float64), np.asarray(stats["std"], np.float64)
        scale = 1.0 / (std + eps)
        return scale, -mean * scale, False
    raise ValueError(f"unknown normalization type {kind!r}")


def action_affine(stats: dict, kind: str, eps: float = 1e-8):
    """(scale, shift) with denormalise(a) = a * scale + shift for every dimension but the last (the gripper action is not
    normalised in the training data, simpler.py:102): identity there."""
    if kind == "bound":
        lo, hi = np.asarray(stats["p01"], np.float64), np.asarray(stats["p99"], np.float64)
        scale, shift = (hi - lo) / 2.0, (hi - lo) / 2.0 + lo          # (a + 1) / 2 * (hi - lo) + lo
    elif kind == "gaussian":
        mean, std = np.asarray(stats["mean"], np.float64), np.asarray(stats["std"], np.float64)
        scale, shift = std + eps, mean
    else:
        raise ValueError(f"unknown normalization type {kind!r}")
    scale, shift = scale.copy(), shift.copy()
    scale[-1], shift[-1] = 1.0, 0.0
    return scale, shift
