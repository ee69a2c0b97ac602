#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - freezes runs of the reference's own CALLERS of the hot path into tests/golden/callers_*.npz:
`Arena.playGames` (Arena.py:175-227, 1-2-2-1 seat order, pit.py:91 players) and `Coach.executeEpisode` (Coach.py:50-100),
both unmodified, over the patched reference Game and the reference's own MCTS.py, with the fixed network of oracle/fakenn.py,
an injected MCTS.rng (oracle/callers_harness.SeqRng) and seeded chance. Run in the build container:

    python oracle/refgen/gen_callers_golden.py

tests/test_ref_callers.py replays the recorded deals / reveals through the B200 mirrors (same callers, `azg_b200.SplendorGame`
and `azg_b200.MCTS`) and through the batched engines (`BatchedArena`, `SelfPlayEngine`) and compares actions, results and
example tuples.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.realpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import build_patched_ref  # noqa: E402
from oracle import callers_harness as ch  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")

ARENA_ARGS = [   # player1, player2 (pit.py:54-61 shape); the same numbers are used by the tests
    dict(numMCTSSims=48, cpuct=1.0, fpu=0.0, prob_fullMCTS=1.0, ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False),
    dict(numMCTSSims=32, cpuct=2.0, fpu=0.2, prob_fullMCTS=1.0, ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False),
]
COACH_ARGS = dict(numMCTSSims=40, prob_fullMCTS=0.5, ratio_fullMCTS=4, forced_playouts=True, cpuct=1.25, fpu=0.1, no_mem_optim=False,
                  temperature=[1.25, 0.8], dirichletAlpha=0.3, tempThreshold=0, no_compression=True)


def _seed(seed):
    from numba import njit

    @njit
    def seed_numba(k):
        np.random.seed(k)
    np.random.seed(seed)
    seed_numba(seed)


def arena_case(n, seed, games, first_best):
    import Arena as ArenaMod
    import MCTS as MctsMod
    from splendor.SplendorGame import SplendorGame
    if first_best:
        MctsMod.np = ch.first_best_numpy()       # ties of the most visited action go to the first one (else np.random.choice)
    else:
        MctsMod.np = np
    game = ch.recording(SplendorGame)(n)
    nets = [ch.FakeWrapper(game, None), ch.FakeWrapper(game, None)]
    mcts = [MctsMod.MCTS(game, nets[i], ch.dotdict(ARENA_ARGS[i])) for i in range(2)]
    players = [(lambda x, m=m: int(np.argmax(m.getActionProb(x, temp=0, force_full_search=True)[0]))) for m in mcts]   # pit.py:91
    arena = ArenaMod.Arena(players[0], players[1], players[1] if n == 3 else None, game, ch.dotdict(lag=False, record_dir=None), no_record=True)
    results, starts = [], []
    orig = arena.playGame

    def play(**kw):
        starts.append(len(game.log))
        r = orig(**kw)
        results.append([float(r[0]), float(r[1]), float(r[2])])
        return r
    arena.playGame = play
    _seed(seed)
    one, two, draws = arena.playGames(games)
    MctsMod.np = np
    starts.append(len(game.log))
    L = max(starts[i + 1] - starts[i] for i in range(games))
    actions = np.full((games, L), -1, dtype=np.int16); reveals = np.full((games, L), -1, dtype=np.int16)
    for g in range(games):
        seg = game.log[starts[g]:starts[g + 1]]
        actions[g, :len(seg)] = [a for a, _ in seg]; reveals[g, :len(seg)] = [c for _, c in seg]
    name = f"callers_arena_n{n}" + ("_firstbest" if first_best else "")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), n=np.int32(n), seed=np.int64(seed), first_best=np.int32(first_best),
                        inits=np.array(game.inits), actions=actions, reveals=reveals, results=np.array(results),
                        totals=np.array([one, two, draws]), nn_calls=np.array([nets[0].calls, nets[1].calls]))
    print(name, "games", games, "moves", [starts[i + 1] - starts[i] for i in range(games)], "one/two/draws", one, two, draws, "results", results)


def coach_case(n, seed):
    import Coach as CoachMod
    import MCTS as MctsMod
    from splendor.SplendorGame import SplendorGame
    MctsMod.np = np
    game = ch.recording(SplendorGame)(n)
    args = ch.dotdict(COACH_ARGS)
    coach = CoachMod.Coach(game, ch.FakeWrapper(game, None), args)
    rng = ch.SeqRng(seed)
    coins = []
    orig_random = rng.random

    def random():
        c = orig_random()
        coins.append(c)
        return c
    rng.random = random
    coach.mcts.rng = rng
    _seed(seed)
    ex = coach.executeEpisode()
    E = len(ex)
    dirs = np.zeros((len(rng.dirs), 406))
    for i, d in enumerate(rng.dirs):
        dirs[i, :len(d)] = d
    np.savez_compressed(os.path.join(GOLD, f"callers_coach_n{n}.npz"), n=np.int32(n), seed=np.int64(seed), init=game.inits[0],
                        actions=np.array([a for a, _ in game.log], dtype=np.int16), reveals=np.array([c for _, c in game.log], dtype=np.int16),
                        coins=np.array(coins), dirs=dirs, dir_len=np.array([len(d) for d in rng.dirs]),
                        board=np.array([e[0] for e in ex], dtype=np.int8), pi=np.array([e[1] for e in ex], dtype=np.float32),
                        winner=np.array([e[2] for e in ex], dtype=np.float32), scdiff=np.array([e[3] for e in ex], dtype=np.int64),
                        valids=np.array([e[4] for e in ex], dtype=np.bool_), surprise=np.array([e[5] for e in ex], dtype=np.float64),
                        nn_calls=np.int64(coach.nnet.calls))
    print(f"callers_coach_n{n}", "moves", len(game.log), "full searches", int(sum(c < args.prob_fullMCTS for c in coins)), "examples", E)


def main():
    build_patched_ref.import_ref(callers=True)
    import warnings
    warnings.filterwarnings("ignore")
    arena_case(2, 11, 4, False)
    arena_case(2, 12, 4, True)
    arena_case(3, 13, 2, True)
    coach_case(2, 21)
    coach_case(3, 22)


if __name__ == "__main__":
    main()
