"""TEST INFRASTRUCTURE - a deterministic stand-in for the network ("fixed NN outputs", BASELINE north_star):
policy and values are a pure function of the state bytes. All probabilities are multiples of 2^-13 that sum to
exactly 1 and all values are multiples of 1/64, so every float32 sum / normalisation over them is exact in any
order - the reference (Numba, fastmath), the C oracle and the CUDA tree then agree bit for bit on Ps.
Used by oracle/refgen/gen_mcts_golden.py (inside the reference's MCTS), tests/ and bench.py's CPU arm.
"""
import numpy as np

M64 = (1 << 64) - 1
TOTAL = 8192
DIR_TOTAL = 4096


def fnv1a64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & M64
    return h


def mix64(x: int) -> int:
    x &= M64
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def predict(state, valids, n_players):
    """-> (Ps float32[406], v float32[n]) for one state (int8[R,7]) and its bool[406] mask"""
    h = fnv1a64(np.ascontiguousarray(state, dtype=np.int8).tobytes())
    idx = np.flatnonzero(valids)
    w = [1 + (mix64(h + int(a) * 0x9E3779B97F4A7C15) >> 58) for a in idx]   # 1..64
    w[0] += TOTAL - sum(w)
    ps = np.zeros(len(valids), dtype=np.float32)
    ps[idx] = np.array(w, dtype=np.float32) / np.float32(TOTAL)
    v = np.array([(((h >> (8 * p + 3)) & 0x7F) - 64) / 64.0 for p in range(n_players)], dtype=np.float32)
    return ps, v


def dirichlet(seed, m):
    """dyadic stand-in for rng.dirichlet([alpha]*m): multiples of 2^-12 summing to exactly 1 (float64)"""
    w = [1 + (mix64(seed * 0x9E3779B97F4A7C15 + i) >> 59) for i in range(m)]   # 1..32
    w[0] += DIR_TOTAL - sum(w)
    return np.array(w, dtype=np.float64) / DIR_TOTAL


class FakeNNet:
    """NeuralNet.predict surface (GenericNNetWrapper.py:141) for the reference's MCTS"""

    def __init__(self, n_players):
        self.n = n_players
        self.calls = 0

    def predict(self, board, valid_actions):
        self.calls += 1
        return predict(board, valid_actions, self.n)
