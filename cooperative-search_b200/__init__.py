"""coopsearch_b200 -- B200-native batched env-step hot path of WZN1ng/Cooperative-Search.

Public surface (mirrors the reference's env protocol):
    VecFlightEasyEnv, VecFlightEnv, VecSearchEnv, VecSimpleSpreadEnv, SingleEnvAdapter, load_targets
    dist.shard_range / dist.allreduce_stats for the one-process-per-GPU launch
"""
from ._lib import CoopSearchError, load as load_library  # noqa: F401
from .policy import BatchedRNNAgents  # noqa: F401
from .vec_flight import VecFlightEasyEnv, VecFlightEnv, HostStepper, DeviceStepper, generate_episodes, load_targets  # noqa: F401
from .adapter import SingleEnvAdapter  # noqa: F401
from .replay import DeviceReplayBuffer, collect_replay_stats  # noqa: F401
from .vec_spread import VecSimpleSpreadEnv  # noqa: F401
from . import dist  # noqa: F401

try:
    from .vec_search import VecSearchEnv  # noqa: F401
except ImportError:  # pragma: no cover
    pass

__all__ = ["VecFlightEasyEnv", "VecFlightEnv", "VecSearchEnv", "VecSimpleSpreadEnv", "SingleEnvAdapter", "HostStepper", "DeviceStepper", "BatchedRNNAgents", "generate_episodes", "load_targets", "DeviceReplayBuffer", "collect_replay_stats",
           "CoopSearchError", "load_library", "dist"]
