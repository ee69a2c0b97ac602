"""B200-native batched Splendor environment + MCTS self-play engine behind the reference's Game / MCTS API.

    from azg_b200 import SplendorGame, SplendorEnv      (import alias: azg_b200.py at the repo root)

The compute path is hand-written sm_100a CUDA in csrc/, reached through the C ABI in include/splendor_b200.h.
"""
from . import _native
from .engine import SplendorEnv, rows
from .game import Board, SplendorGame, action_size, observation_size
from .mcts import MCTS, MCTSArena
from .nnet import FusedSplendorNNet, SplendorNNetB200
from .selfplay import SelfPlayEngine
from .arena import BatchedArena
from . import examples, multigpu, nnet, trainbatch
from .trainbatch import TrainBatcher

__all__ = ["SplendorEnv", "SplendorGame", "Board", "MCTS", "MCTSArena", "SplendorNNetB200", "FusedSplendorNNet", "SelfPlayEngine", "BatchedArena", "TrainBatcher", "nnet", "examples", "multigpu", "trainbatch", "observation_size", "action_size", "rows", "_native"]
