"""muax_b200 — B200-native batched MuZero search behind the muax agent API (MuZero.act / muax.policy)."""
from . import nn, random, utils  # noqa: F401
from .model import MuZero  # noqa: F401
from .nn import MZNetwork, MZNetworkParams, create_muzero_network  # noqa: F401
from .policy import (GumbelMuZeroPolicy, MuZeroPolicy, Policy, PolicyOutput, RecurrentFnOutput,  # noqa: F401
                     RootFnOutput, StochasticMuZeroPolicy, qtransform_by_parent_and_siblings,
                     qtransform_completed_by_mix_value)
from .search import SearchEngine  # noqa: F401
from .actor import BatchedPNStep, CartPoleVec, TrajectoryStore, Transitions, VectorActor  # noqa: F401
from .train import fit, test  # noqa: F401

__all__ = ["MuZero", "MZNetwork", "MZNetworkParams", "create_muzero_network", "Policy", "MuZeroPolicy",
           "GumbelMuZeroPolicy", "StochasticMuZeroPolicy", "PolicyOutput", "RootFnOutput", "RecurrentFnOutput",
           "SearchEngine", "fit", "test", "BatchedPNStep", "TrajectoryStore", "Transitions", "VectorActor", "CartPoleVec", "nn", "random",
           "utils"]
