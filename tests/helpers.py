"""Shared helpers for the parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import open_pi_zero_b200 as pz  # noqa: E402

# kernel-compatible small shape: real attention geometry (8 heads x 256, MQA),
# SigLIP head_dim 72, everything else shrunk
SMALL = pz.make_dims(
    vocab_size=1024, image_token_index=1000, image_size=56, num_image_tokens=16,
    max_image_text_tokens=24, num_layers=3, vlm_hidden=256, vlm_inter=512, act_hidden=128,
    act_inter=256, vit_hidden=144, vit_inter=256, vit_layers=2, vit_heads=2)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """relative Frobenius error ||a-b|| / ||b||"""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b) -> float:
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def valid_rows(t: torch.Tensor, valid_len) -> torch.Tensor:
    """concatenate rows [0, valid_len[b]) of each sample of a [B, S, D] tensor"""
    return torch.cat([t[b, : int(valid_len[b])] for b in range(t.shape[0])], 0)
