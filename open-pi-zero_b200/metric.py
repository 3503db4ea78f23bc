"""Validation metrics around `infer_action` (`src/utils/metric.py:6-21`, used by `TrainAgent.run` at `train.py:418-447`):
the share of action vectors whose every dimension is within a threshold of the ground truth, the L1 loss, and their mean
over the data-parallel ranks (two `all_reduce` calls, train.py:446-447).  Host-side logic on tiny tensors; stays on the
device of its inputs (no host synchronisation)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def get_action_accuracy(gt: torch.Tensor, pred: torch.Tensor, thresholds: Sequence[float] = (0.1, 0.2)) -> torch.Tensor:
    """metric.py:6-21, vectorised over the thresholds: `[len(thresholds)]` accuracies."""
    diff = torch.abs(gt - pred).reshape(-1, gt.shape[-1])                          # [B * horizon, action_dim]
    th = torch.as_tensor(list(thresholds), dtype=diff.dtype, device=diff.device)
    inside = (diff[None] < th[:, None, None]).float().mean(dim=2) >= 1.0           # every dimension under the threshold
    return inside.float().mean(dim=1)


def eval_stats(preds: List[torch.Tensor], gts: List[torch.Tensor], thresholds: Sequence[float] = (0.1, 0.2), group=None):
    """Mean accuracy / L1 over a list of validation batches and over the ranks (train.py:420-449)."""
    import torch.distributed as dist
    acc = torch.stack([get_action_accuracy(g, p, thresholds) for p, g in zip(preds, gts)]).mean(0)
    l1 = torch.stack([torch.nn.functional.l1_loss(p, g) for p, g in zip(preds, gts)]).mean()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(l1, op=dist.ReduceOp.SUM, group=group)
        acc /= dist.get_world_size(group)
        l1 /= dist.get_world_size(group)
    return acc, l1
