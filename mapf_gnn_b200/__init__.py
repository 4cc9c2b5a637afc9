"""mapf_gnn_b200 -- B200-native batched decentralized-policy step of MAPF-GNN.

Python/PyTorch API of the reference (``GraphEnv``, ``GCNLayer``, ``MessagePassingLayer``, the three
``Network`` classes) re-backed by hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/mapf_b200.h``.  Importing this package loads (and, when the sources changed, rebuilds)
``lib/libmapf_b200.so``; there is no CPU or PyTorch fallback for any compute.
"""
from . import _lib

_lib.load()

from . import ops  # noqa: E402,F401
from .env import BatchedGraphEnv, GraphEnv, create_goals, create_obstacles, d2_limit_from_range  # noqa: E402,F401
from .layers import GCNLayer, MessagePassingLayer  # noqa: E402,F401
from .network import BaselineNetwork, GNNNetwork, MessageNetwork, weights_init  # noqa: E402,F401
from .rollout import HostSteppedRollout, PipelinedHostRollout, Rollout, build_network  # noqa: E402,F401
from .scenarios import make_scenarios, make_scenarios_device  # noqa: E402,F401
from .evaluate import evaluate  # noqa: E402,F401
from . import cbs  # noqa: E402,F401

__all__ = ["BatchedGraphEnv", "GraphEnv", "create_goals", "create_obstacles", "d2_limit_from_range", "GCNLayer",
           "MessagePassingLayer", "GNNNetwork", "MessageNetwork", "BaselineNetwork", "weights_init", "Rollout", "HostSteppedRollout", "PipelinedHostRollout",
           "build_network", "make_scenarios", "make_scenarios_device", "evaluate", "cbs"]
