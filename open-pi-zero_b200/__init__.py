"""open-pi-zero_b200 -- B200-native `PiZero.infer_action`.

The directory name follows the project layout (`open-pi-zero_b200/`); import it
as `open_pi_zero_b200` (the sibling alias package adds this directory to its
search path).
"""
from .config import BRIDGE_DIMS, PI0_PAPER_DIMS, AttrDict, cfg_from_dims, dims_from_cfg, make_dims
from .synth import init_state_dict, make_inputs, state_dict_spec

__all__ = ["BRIDGE_DIMS", "PI0_PAPER_DIMS", "AttrDict", "cfg_from_dims", "dims_from_cfg",
           "make_dims", "init_state_dict", "make_inputs", "state_dict_spec", "PiZero",
           "PiZeroInference", "JointModel", "KVCache", "TextKVCache", "FlowTimeSampler", "GradBuffer", "FusedAdamW",
           "OverlappedAllReduce", "ModelAveraging", "CosineAnnealingWarmupRestarts", "flow_matching_step", "VLAProcessor"]


def __getattr__(name):   # lazy: importing the package must not need torch.cuda / the .so
    if name in ("GradBuffer", "FusedAdamW", "OverlappedAllReduce", "ModelAveraging", "CosineAnnealingWarmupRestarts",
                "flow_matching_step", "allreduce_gradients"):
        from . import train
        return getattr(train, name)
    if name in ("PiZero", "PiZeroInference", "JointModel", "KVCache", "TextKVCache", "PzError"):
        from . import pizero
        return getattr(pizero, name)
    if name == "VLAProcessor":
        from .processing import VLAProcessor
        return VLAProcessor
    if name == "FlowTimeSampler":
        from .flow import FlowTimeSampler
        return FlowTimeSampler
    raise AttributeError(name)
