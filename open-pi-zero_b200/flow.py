"""Host-side pieces of the flow-matching training step around `PiZero.forward`
(SURVEY 8f-1): the timestep sampler of `TrainAgent.sample_fm_time`
(`src/agent/train.py:217-247`).  Pure host logic on tiny tensors; the loss
itself runs through `pz_flow_matching_loss`.
"""
from __future__ import annotations

import torch


class FlowTimeSampler:
    """`flow_sampling` in {"uniform", "beta"} (train.py:217-226).

    beta (pi0 paper): t = (1 - flow_sig_min) * (1 - z), z ~ Beta(flow_alpha, flow_beta)   (train.py:244-246)
    uniform: stratified over the batch, t_i = (u + i / bsz) mod (1 - 1e-5), u ~ U(0, 1)   (train.py:240-243)
    """

    def __init__(self, flow_sampling: str = "beta", flow_alpha: float = 1.5, flow_beta: float = 1.0,
                 flow_sig_min: float = 0.001):
        assert flow_sampling in ["uniform", "beta"], f"Invalid flow matching timestep sampling mode: {flow_sampling}"
        self.flow_sampling = flow_sampling
        if flow_sampling == "beta":
            self.flow_t_max = 1 - flow_sig_min
            self.flow_beta_dist = torch.distributions.Beta(flow_alpha, flow_beta)

    @classmethod
    def from_cfg(cls, cfg) -> "FlowTimeSampler":
        g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        return cls(g("flow_sampling", "beta"), g("flow_alpha", 1.5), g("flow_beta", 1), g("flow_sig_min", 0.001))

    def sample_fm_time(self, bsz: int) -> torch.Tensor:
        if self.flow_sampling == "uniform":
            eps = 1e-5
            return (torch.rand(1) + torch.arange(bsz) / bsz) % (1 - eps)
        z = self.flow_beta_dist.sample((bsz,))
        return self.flow_t_max * (1 - z)
