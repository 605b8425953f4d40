"""marl_dpp_b200 -- B200-native (sm_100a) batched simulator + tabular learner for the MARL-DPP
hot path (reference: DosepackAIR/MARL-DPP src/env_gym.py, src/layout.py, src/algorithms/qlearning.py).

Host code is Python + PyTorch (device memory, streams, torch.distributed); the compute is
hand-written CUDA behind the C ABI in include/mdpp.h.  No CPU fallback.
"""
from .layout import layout_3, BLOCK_NAMES  # noqa: F401
from .tables import Scenario, StaticGraph, learner_params, epsilon_schedule  # noqa: F401

__all__ = ["layout_3", "BLOCK_NAMES", "Scenario", "StaticGraph", "learner_params", "epsilon_schedule"]
