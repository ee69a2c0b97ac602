"""TEST INFRASTRUCTURE - pieces that let the reference's own callers (Arena.py, Coach.py; unmodified) run once over the
reference's Game / MCTS and once over the B200 mirrors on identical inputs:

  * FakeWrapper  - the NeuralNet surface Coach / MCTS need (predict, args, re-constructible as cls(game, args)) around the
                   fixed network of oracle/fakenn.py ("fixed NN outputs")
  * SeqRng       - stands in for MCTS.rng (the reference uses an unseeded np.random.default_rng()): a fixed sequence of
                   playout-cap coins and dyadic Dirichlet vectors
  * recording / replaying Game subclasses - chance is the only thing the two sides cannot share (Numba's MT19937 inside
    njit vs Philox): the reference run records the initial deal and every revealed card, the mirror run replays them
    through getNextState(..., reveal=code)

Used by oracle/refgen/gen_callers_golden.py (to freeze tests/golden/callers_*.npz) and by tests/test_ref_callers.py.
"""
import numpy as np

from oracle import fakenn


class FakeWrapper:
    def __init__(self, game, args):
        self.game, self.args, self.n = game, args, game.num_players
        self.calls = 0

    def predict(self, board, valid_actions):
        self.calls += 1
        return fakenn.predict(board, valid_actions, self.n)


class SeqRng:
    def __init__(self, seed):
        self.seed, self.k = int(seed), 0
        self.dirs = []

    def random(self):
        self.k += 1
        return (fakenn.mix64(self.seed * 0x9E3779B97F4A7C15 + self.k) >> 11) / float(1 << 53)

    def dirichlet(self, alphas):
        self.k += 1
        d = fakenn.dirichlet(self.seed * 131 + self.k, len(alphas))
        self.dirs.append(d)
        return d


def deck_diff(before, after):
    """colour*8+idx of the card drawn between two states, or -1"""
    out = -1
    for t in range(3):
        rb, ra = before[25 + 2 * t + 1, :5].astype(np.uint8), after[25 + 2 * t + 1, :5].astype(np.uint8)
        for c in range(5):
            d = int(rb[c]) & ~int(ra[c])
            if d:
                assert out == -1 and bin(d).count("1") == 1
                out = c * 8 + (7 - d.bit_length() + 1)
    return out


def recording(game_cls):
    """subclass of a Game that logs (initial board, [(action, revealed card code)]) of everything played through it"""
    class Recording(game_cls):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.inits, self.log = [], []

        def getInitBoard(self):
            b = super().getInitBoard()
            self.inits.append(np.array(b, dtype=np.int8).copy())
            return b

        def getNextState(self, board, player, action, deterministic=False):
            before = np.array(board, dtype=np.int8).copy()
            nb, nxt = super().getNextState(board, player, action, deterministic)
            self.log.append((int(action), deck_diff(before, nb)))
            return nb, nxt
    return Recording


def replaying(game_cls):
    """subclass of the mirror Game that starts from recorded boards and replays recorded reveals (the actions are the
    caller's own: the test compares them with the recorded ones afterwards)"""
    class Replaying(game_cls):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.inits, self.script, self.pos, self.init_pos, self.actions = [], [], 0, 0, []

        def getInitBoard(self):
            b = np.array(self.inits[self.init_pos], dtype=np.int8).copy()
            self.init_pos += 1
            self.board.copy_state(b, False)
            return self.board.get_state()

        def getNextState(self, board, player, action, deterministic=False):
            code = self.script[self.pos][1] if self.pos < len(self.script) else -1
            self.pos += 1
            self.actions.append(int(action))
            if code < 0:   # nothing was drawn on the reference side (empty deck, or a move that reveals nothing)
                return super().getNextState(board, player, action, deterministic=True)
            return super().getNextState(board, player, action, reveal=code)
    return Replaying


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


def first_best_numpy():
    """a stand-in for the `np` global of a module whose np.random.choice(bests) tie-break must be reproducible without
    the global generator: picks the FIRST of the candidates"""
    class _Rand:
        def __getattr__(self, name):
            return getattr(np.random, name)

        @staticmethod
        def choice(a, *args, **kw):
            if args or kw:
                return np.random.choice(a, *args, **kw)
            return a[0]

    class _Np:
        random = _Rand()

        def __getattr__(self, name):
            return getattr(np, name)
    return _Np()
