"""The reference's own CALLERS of the hot path against the B200 mirrors and the batched engines.

Fixtures tests/golden/callers_*.npz = runs of the unmodified Arena.py / Coach.py over the patched reference Game and the
reference's own MCTS.py (oracle/refgen/gen_callers_golden.py: fixed network, injected MCTS.rng, seeded chance, every deal and
revealed card recorded). Four comparisons, each replaying the recorded deals / reveals:

  1. Arena.playGames (Arena.py:64-227, unmodified) driving azg_b200.SplendorGame + azg_b200.MCTS      -> same moves, same results
  2. Coach.executeEpisode (Coach.py:50-100, unmodified) driving the same mirrors                        -> same example tuples
  3. BatchedArena (all games as lanes, two tree arenas)                                                 -> same moves, same totals   [N2 pin]
  4. SelfPlayEngine + ExampleBuffer + expand_symmetries (all lanes per move, examples on the device)    -> same example tuples       [N1 pin]

1 and 2 need the callers' source files: built from /root/reference in the build container, or the travelling copy
oracle/_ref/pyref (`python oracle/refgen/build_patched_ref.py --travel`; git-ignored) on the GPU box; they skip without it.
3 and 4 only need the fixtures.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import callers_harness as ch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle", "refgen"))
import build_patched_ref  # noqa: E402
import gen_callers_golden as gcg  # noqa: E402  (only its argument tables are used here)


def _ref_callers():
    d = build_patched_ref.find_ref(callers=True)
    if d is None:
        pytest.skip("no copy of the reference's callers (Arena.py / Coach.py): run oracle/refgen/build_patched_ref.py --travel in the build container")
    return d


def _np(x):
    return x.cpu().numpy()


# ---------------------------------------------------------------------------------------------- 1. Arena.py over the mirrors
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["callers_arena_n2", "callers_arena_n2_firstbest", "callers_arena_n3_firstbest"])
def test_reference_arena_drives_the_mirrors(golden_dir, name):
    _ref_callers()
    import azg_b200 as az
    import Arena as ArenaMod
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, games = int(g["n"]), g["inits"].shape[0]
    game = ch.replaying(az.SplendorGame)(n)
    game.inits = list(g["inits"])
    game.script = [(int(a), int(c)) for gi in range(games) for a, c in zip(g["actions"][gi], g["reveals"][gi]) if a >= 0]
    nets = [ch.FakeWrapper(game, None), ch.FakeWrapper(game, None)]
    mcts = [az.MCTS(game, nets[i], ch.dotdict(gcg.ARENA_ARGS[i])) for i in range(2)]
    players = [(lambda x, m=m: int(np.argmax(m.getActionProb(x, temp=0, force_full_search=True)[0]))) for m in mcts]   # pit.py:91
    saved_np, saved_mcts = az.mcts.np, ArenaMod.MCTS
    try:
        if int(g["first_best"]):
            az.mcts.np = ch.first_best_numpy()
        ArenaMod.MCTS = az.MCTS            # the binding a maintainer changes: `from MCTS import MCTS` (Arena.py:9)
        arena = ArenaMod.Arena(players[0], players[1], players[1] if n == 3 else None, game, ch.dotdict(lag=False, record_dir=None), no_record=True)
        results = []
        orig = arena.playGame
        arena.playGame = lambda **kw: (results.append(orig(**kw)), results[-1])[1]
        np.random.seed(int(g["seed"]))
        one, two, draws = arena.playGames(games)
    finally:
        az.mcts.np, ArenaMod.MCTS = saved_np, saved_mcts
    ref_actions = [int(a) for gi in range(games) for a in g["actions"][gi] if a >= 0]
    assert game.actions == ref_actions
    assert [one, two, draws] == [int(x) for x in g["totals"]]
    assert np.array_equal(np.array([[float(x) for x in r] for r in results]), g["results"])
    assert [nets[0].calls, nets[1].calls] == [int(x) for x in g["nn_calls"]]


# ---------------------------------------------------------------------------------------------- 2. Coach.py over the mirrors
@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3])
def test_reference_coach_episode_on_the_mirrors(golden_dir, n):
    _ref_callers()
    import azg_b200 as az
    import Coach as CoachMod
    g = np.load(os.path.join(golden_dir, f"callers_coach_n{n}.npz"))
    game = ch.replaying(az.SplendorGame)(n)
    game.inits = [g["init"]]
    game.script = [(int(a), int(c)) for a, c in zip(g["actions"], g["reveals"])]
    saved = CoachMod.MCTS
    try:
        CoachMod.MCTS = az.MCTS            # `from MCTS import MCTS` (Coach.py:14)
        coach = CoachMod.Coach(game, ch.FakeWrapper(game, None), ch.dotdict(gcg.COACH_ARGS))
        coach.mcts.rng = ch.SeqRng(int(g["seed"]))
        np.random.seed(int(g["seed"]))
        ex = coach.executeEpisode()
    finally:
        CoachMod.MCTS = saved
    assert game.actions == [int(a) for a in g["actions"]]
    assert len(ex) == g["board"].shape[0] and coach.nnet.calls == int(g["nn_calls"])
    assert np.array_equal(np.array([e[0] for e in ex]), g["board"])
    assert np.array_equal(np.array([e[4] for e in ex]), g["valids"])
    assert np.array_equal(np.array([e[2] for e in ex], dtype=np.float32), g["winner"])
    assert np.array_equal(np.array([e[3] for e in ex]), g["scdiff"])
    assert np.abs(np.array([e[1] for e in ex], dtype=np.float32) - g["pi"]).max() < 1e-6
    assert np.abs(np.array([e[5] for e in ex]) - g["surprise"]).max() < 1e-6


# ---------------------------------------------------------------------------------------------- 3. BatchedArena vs Arena.playGames
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["callers_arena_n2_firstbest", "callers_arena_n3_firstbest"])
def test_batched_arena_reproduces_reference_playgames(golden_dir, name):
    """N2 pin: every recorded game is a lane, all lanes move together, each player owns one tree arena; the most visited action
    with the first-best tie-break (the fixture was recorded with the same rule)"""
    import azg_b200 as az
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, games = int(g["n"]), g["inits"].shape[0]
    helper = az.MCTSArena(n, 1, node_cap=64)
    fixed = lambda s, v: helper.fixed_net(s, v)
    a = gcg.ARENA_ARGS
    pit = az.BatchedArena(n, [fixed, fixed], num_sims=[a[0]["numMCTSSims"], a[1]["numMCTSSims"]], cpuct=[a[0]["cpuct"], a[1]["cpuct"]],
                          fpu=[a[0]["fpu"], a[1]["fpu"]], forced_playouts=[a[0]["forced_playouts"], a[1]["forced_playouts"]], node_cap=8192)
    one, two, draws, d = pit.play_games(games, init_boards=g["inits"], reveals=g["reveals"], record_actions=True)
    acts = _np(d["actions"])
    for gi in range(games):
        ref = [int(x) for x in g["actions"][gi] if x >= 0]
        assert [int(x) for x in acts[gi, :len(ref)]] == ref, (gi,)
        assert int(d["moves"][gi]) == len(ref)
    assert [one, two, draws] == [int(x) for x in g["totals"]]
    assert np.array_equal(_np(d["result_seat0"]).astype(np.float64), g["results"][:, 0])


# ---------------------------------------------------------------------------------------------- 4. SelfPlayEngine vs Coach.executeEpisode
@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3])
def test_selfplay_engine_reproduces_reference_episode(golden_dir, n):
    """N1 pin: the recorded episode replayed in three lanes at once (its playout-cap coins, Dirichlet vectors, chosen actions and
    revealed cards); the examples the engine assembles on the device - boards, policy targets, legal masks, winners and score
    differences rolled into each mover's frame, all symmetric variants in order - must be the tuples Coach returned"""
    import azg_b200 as az
    g = np.load(os.path.join(golden_dir, f"callers_coach_n{n}.npz"))
    A, T = gcg.COACH_ARGS, 3
    eng = az.SelfPlayEngine(n, T, None, A["numMCTSSims"], seed=1, cpuct=A["cpuct"], fpu=A["fpu"], ratio_full=A["ratio_fullMCTS"],
                            forced_playouts=A["forced_playouts"], dirichlet_noise=True, dirichlet_alpha=A["dirichletAlpha"],
                            temperature0=A["temperature"][0], node_cap=8192, record_examples=True)
    eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v)
    dev = eng.device
    eng.env.set_states(torch.from_numpy(np.repeat(g["init"][None], T, 0)))
    full = g["coins"] < A["prob_fullMCTS"]
    k = 0
    for mv in range(len(g["actions"])):
        is_full = torch.full((T,), bool(full[mv]), dtype=torch.bool, device=dev)
        dirv = None
        if full[mv]:
            dirv = torch.from_numpy(np.repeat(g["dirs"][k][None], T, 0)).to(dev).contiguous(); k += 1
        code = int(g["reveals"][mv])
        eng.play_move(1.0, is_full=is_full, dir_values=dirv, forced_actions=torch.full((T,), int(g["actions"][mv]), dtype=torch.int16),
                      reveals=torch.full((T,), 255 if code < 0 else code, dtype=torch.uint8))
    assert int(eng.games_finished.item()) == T
    ex = eng.drain_examples(symmetries=True)
    E = g["board"].shape[0]
    assert ex["board"].shape[0] == T * E
    for t in range(T):       # finished lanes hand their examples over lane by lane
        sl = slice(t * E, (t + 1) * E)
        assert np.array_equal(_np(ex["board"][sl]), g["board"])
        assert np.array_equal(_np(ex["valids"][sl]).astype(bool), g["valids"])
        assert np.array_equal(_np(ex["winner"][sl]), g["winner"])
        assert np.array_equal(_np(ex["scdiff"][sl]).astype(np.int64), g["scdiff"])
        assert np.abs(_np(ex["pi"][sl]) - g["pi"]).max() < 1e-6
        assert np.abs(_np(ex["surprise"][sl]).astype(np.float64) - g["surprise"]).max() < 1e-6
    # ... and in Coach's own container format: tuples, zlib + pickle, checkpoint.examples (Coach.py:91-100,167-208)
    one = {k: v[:E] for k, v in ex.items()}
    tup = az.examples.to_coach_format(one, compress=False)
    assert isinstance(tup[0], tuple) and len(tup[0]) == 6 and tup[0][0].shape == g["board"][0].shape and tup[0][4].dtype == np.bool_


# ---------------------------------------------------------------------------------------------- CPU: container formats, printBoard
def _fixture_examples(golden_dir, n=2, count=40):
    g = np.load(os.path.join(golden_dir, f"callers_coach_n{n}.npz"))
    return dict(board=torch.from_numpy(g["board"][:count]), pi=torch.from_numpy(g["pi"][:count]), winner=torch.from_numpy(g["winner"][:count]),
                scdiff=torch.from_numpy(g["scdiff"][:count].astype(np.int32)), valids=torch.from_numpy(g["valids"][:count].astype(np.uint8)),
                surprise=torch.from_numpy(g["surprise"][:count].astype(np.float32)))


def test_checkpoint_examples_round_trip(golden_dir, tmp_path):
    """Coach.saveTrainExamples / loadTrainExamples (Coach.py:167-208): history of iterations, compressed or not, trimmed"""
    from collections import deque
    from azg_b200 import examples as exm
    ex = _fixture_examples(golden_dir)
    comp = exm.to_coach_format(ex, compress=True)
    hist = [deque(comp[:25], maxlen=1000), deque(comp[25:], maxlen=1000)]
    path = exm.save_train_examples(hist, str(tmp_path / "ckpt"))
    assert os.path.basename(path) == "checkpoint.examples"
    back = exm.load_train_examples(path, no_compression=True)
    assert len(back) == 2 and len(back[0]) == 25 and type(back[0][0]) is tuple
    again = exm.from_coach_format(list(back[0]) + list(back[1]))
    for k in exm.FIELDS:
        assert torch.equal(again[k], ex[k]), k
    trimmed = exm.load_train_examples(path, no_compression=False, num_iters_history=1, maxlen_of_queue=10)
    assert len(trimmed) == 1 and len(trimmed[0]) == 10 and isinstance(trimmed[0][0], bytes)


@pytest.mark.ref
def test_reference_coach_loads_our_checkpoint_examples(golden_dir, tmp_path):
    """the file the engine writes is read by the reference's own Coach.loadTrainExamples (Coach.py:175-208), unmodified"""
    _ref_callers()
    import Coach as CoachMod
    from azg_b200 import examples as exm
    ex = _fixture_examples(golden_dir)
    hist = [exm.to_coach_format(ex, compress=True)]
    folder = str(tmp_path / "run")
    exm.save_train_examples(hist, folder)

    class _G:
        num_players = 2
    coach = CoachMod.Coach.__new__(CoachMod.Coach)
    coach.args = ch.dotdict(load_folder_file=os.path.join(folder, "best.pt"), no_compression=True, numItersHistory=5, maxlenOfQueue=1000)
    coach.trainExamplesHistory = []
    coach.loadTrainExamples()
    got = coach.trainExamplesHistory
    assert len(got) == 1 and len(got[0]) == 40 and type(got[0][0]) is tuple
    assert np.array_equal(got[0][3][0], _np(ex["board"][3])) and np.allclose(got[0][3][1], _np(ex["pi"][3]))


def test_print_board_text(golden_dir):
    """printBoard (SplendorGame.py:72-75): plain-text rendering of the same content as print_board"""
    from azg_b200 import game
    g = np.load(os.path.join(golden_dir, "callers_coach_n3.npz"))
    txt = game.board_to_text(g["board"][300], 3)
    lines = txt.splitlines()
    assert lines[0].startswith("Round ") and lines[1].startswith("Nobles:") and sum(l.startswith("Tier ") for l in lines) == 3
    assert sum(l.startswith("Player ") for l in lines) == 3 and any(l.startswith("Bank:") for l in lines)
    st = g["board"][300]
    assert f"{int(st[0, 0])}W" in [l for l in lines if l.startswith("Bank:")][0]
