"""GPU parity tests of the MCTS tree arena, through the C ABI (include/splendor_b200.h, spl_mcts_*).

Bar (BASELINE.json north_star): MCTS visit counts identical, Q within 1e-6 (we hold 1e-12) for fixed NN outputs.
  * golden: tests/golden/mcts_*.npz - outputs of the reference's own MCTS.py (cross-move tree reuse, transpositions,
    forced playouts, Dirichlet noise, playout cap, 2/3/4 players), network = the fixed dyadic function
  * live: random mid-game roots vs the C search oracle, many trees per launch, ragged budgets
  * the drop-in `MCTS` class driven like Coach/Arena with a host `predict`
  * the leaf evaluator on the device vs the reference network's recorded outputs
"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SCEN = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mcts_*.npz")))


def _azg():
    import azg_b200
    return azg_b200


def _np(x):
    return x.cpu().numpy()


@pytest.mark.parametrize("name", SCEN)
def test_arena_reproduces_reference_mcts(golden_dir, name):
    az = _azg()
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    T = 3   # the same search in three trees at once: they must not disturb each other
    temp0 = float(g["temp0"]) if "temp0" in g else 1.0          # args.temperature[0]: root softmax before the noise
    ar = az.MCTSArena(n, T, node_cap=4096, cpuct=cpuct, fpu=fpu, temperature0=temp0)
    dev = ar.device
    for i in range(len(g["ns"])):
        full = bool(g["full"][i])
        nb = sims if full else sims // ratio
        fl = (1 if (full and forced) else 0) | (2 if (full and noise) else 0)
        roots = torch.from_numpy(np.repeat(g["root"][i][None], T, 0)).to(dev)
        simt = torch.full((T,), nb, dtype=torch.int32, device=dev)
        flt = torch.full((T,), fl, dtype=torch.uint8, device=dev)
        dirv = torch.from_numpy(np.repeat(g["dir"][i][None], T, 0)).to(dev).contiguous() if noise else None
        ar.search(roots, simt, lambda s, v: ar.fixed_net(s, v), flt, dirv)
        ar.check_status()
        st = ar.root_stats()
        probs, q = ar.policy(1.0)
        for t in range(T):
            assert np.array_equal(_np(st["nsa"][t]).astype(np.int64), g["nsa"][i]), (name, i, t, "visit counts")
            assert int(st["ns"][t]) == g["ns"][i] and int(st["nodes"][t]) == g["nodes"][i] and int(st["nn_calls"][t]) == g["nn_calls"][i]
            assert np.allclose(_np(st["qsa"][t]), g["qsa"][i], rtol=0, atol=1e-12)
            assert _np(st["qs"])[t] == g["qs"][i]
            assert np.allclose(_np(probs[t]), g["probs"][i], rtol=0, atol=1e-12)
            assert np.allclose(_np(q[t]), g["q"][i], rtol=0, atol=1e-12)
            if "ps" in g:      # root priors after the temperature softmax + noise: 2 float32 ulps (the reference's fastmath sums)
                assert np.abs(_np(st["ps"][t]) - g["ps"][i]).max() <= 2.4e-7


@pytest.mark.parametrize("n", [2, 3, 4])
def test_arena_vs_search_oracle_many_trees(n):
    """64 different mid-game roots per launch, ragged budgets and flags, three consecutive moves with tree reuse and
    a pool small enough to force the ply-based cleaning"""
    az = _azg()
    from oracle import pyoracle as po
    T, rng = 64, np.random.default_rng(100 + n)
    kw = dict(cpuct=1.7, fpu=0.15)
    # n = 3 also exercises a descend call that yields after 4 edges, n = 4 two (descend, rules, attach) passes per wave
    ar = az.MCTSArena(n, T, node_cap=900 if n == 2 else 1600, max_levels=4 if n == 3 else 0, rounds=2 if n == 4 else 1, **kw)
    dev = ar.device
    boards, oracles, budgets, flags = [], [], [], []
    for t in range(T):
        b = po.Board(n); b.init_philox(4242, t)
        for _ in range(int(rng.integers(0, 40 * n))):
            if b.check_end_game().any():
                break
            v = b.valid_moves(0)
            b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 4242, t, 0); b.swap_players(1)
        if b.check_end_game().any():
            b.init_philox(4242, 1000 + t)
        boards.append(b)
        sims = int(rng.integers(40, 260))
        fl = int(rng.integers(0, 4))
        budgets.append(sims); flags.append(fl)
        oracles.append(po.MCTSOracle(n, sims, forced_playouts=bool(fl & 1), dirichlet_noise=bool(fl & 2), **kw))
    cleaned = 0
    for mv in range(4):
        roots = torch.from_numpy(np.stack([b.state for b in boards])).to(dev)
        dirs = np.zeros((T, 406))
        for t, b in enumerate(boards):
            k = int(b.valid_moves(0).sum())
            dirs[t, :k] = rng.dirichlet([0.3] * k)
        ar.search(roots, torch.tensor(budgets, dtype=torch.int32, device=dev), lambda s, v: ar.fixed_net(s, v),
                  torch.tensor(flags, dtype=torch.uint8, device=dev), torch.from_numpy(dirs).to(dev))
        ar.check_status()
        st = ar.root_stats()
        probs, q = ar.policy(1.0)
        nsa, qsa, ns, qs, nnc = _np(st["nsa"]), _np(st["qsa"]), _np(st["ns"]), _np(st["qs"]), _np(st["nn_calls"])
        assert int(st["resets"].max()) == 0
        cleaned = int(st["cleanings"].max())
        for t, b in enumerate(boards):
            if b.check_end_game().any():
                continue
            o = oracles[t].get_action_prob(b.state, temp=1.0, full_search=True, dir_values=dirs[t])
            assert np.array_equal(o["nsa"], nsa[t].astype(np.int64)), (mv, t, "visit counts")
            assert o["ns"] == ns[t] and oracles[t].nn_calls == nnc[t]
            assert np.allclose(o["qsa"], qsa[t], rtol=0, atol=1e-12) and o["qs"] == qs[t]
            assert np.allclose(o["probs"], _np(probs[t]), rtol=0, atol=1e-12) and np.allclose(o["q"], _np(q[t]), rtol=0, atol=1e-12)
            a = int(np.argmax(o["nsa"]))
            b.make_move(a, 0, -2 if (t + mv) % 3 == 0 else -1, 4242, t, 0); b.swap_players(1)
    assert cleaned > 0 or n > 2      # (n=2 fills the pool; wider games revisit more and stay below it)


def test_fixed_net_kernel_matches_oracle():
    az = _azg()
    from oracle import pyoracle as po
    for n in (2, 3, 4):
        ar = az.MCTSArena(n, 1, node_cap=64)
        states, valids = [], []
        b = po.Board(n); b.init_philox(5, n)
        rng = np.random.default_rng(n)
        for _ in range(50):
            v = b.valid_moves(0)
            states.append(b.state.copy()); valids.append(v.astype(np.uint8))
            b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -1); b.swap_players(1)
        s = torch.from_numpy(np.stack(states)).to(ar.device); va = torch.from_numpy(np.stack(valids)).to(ar.device)
        pi, v = ar.fixed_net(s, va)
        for i in range(len(states)):
            p0, v0 = po.fake_predict(states[i], valids[i], n)
            assert np.array_equal(_np(pi[i]), p0) and np.array_equal(_np(v[i]), v0)


class _FakeRng:
    def __init__(self, coin, seed):
        self.coin, self.seed, self.last_dir = coin, seed, None

    def random(self):
        return self.coin

    def dirichlet(self, alphas):
        from oracle import fakenn
        self.seed += 1
        self.last_dir = fakenn.dirichlet(self.seed, len(alphas))
        return self.last_dir


class _DotDict(dict):
    def __getattr__(self, name):
        return self[name]


@pytest.mark.parametrize("name", ["b_n2_forced_noise", "e_n2_cap", "g_n3_late", "h_n2_temp"])
def test_mcts_class_is_a_drop_in(golden_dir, name):
    """the reference-surface class driven exactly as oracle/refgen/gen_mcts_golden.py drove the reference's MCTS:
    same args object, a host `predict` network, an injected rng - same returned probs / q / flags"""
    az = _azg()
    from oracle import fakenn
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    game = az.SplendorGame(n)
    args = _DotDict(numMCTSSims=sims, prob_fullMCTS=prob_full, ratio_fullMCTS=ratio, forced_playouts=bool(forced), cpuct=cpuct, fpu=fpu,
                    no_mem_optim=True, temperature=[float(g["temp0"]) if "temp0" in g else 1.0, 1.0], dirichletAlpha=0.3)
    nnet = fakenn.FakeNNet(n)
    mcts = az.MCTS(game, nnet, args, dirichlet_noise=bool(noise))
    mcts.rng = _FakeRng(0.5, seed=sum(map(ord, name)))
    for i in range(len(g["ns"])):
        probs, q, full = mcts.getActionProb(g["root"][i], temp=1, force_full_search=bool(force))
        assert isinstance(probs, list) and len(probs) == 406 and isinstance(q, list) and len(q) == n and isinstance(full, bool)
        assert full == bool(g["full"][i])
        assert np.allclose(probs, g["probs"][i], rtol=0, atol=1e-12) and np.allclose(q, g["q"][i], rtol=0, atol=1e-12)
        assert nnet.calls == g["nn_calls"][i]
        st = mcts.root_stats()
        assert np.array_equal(st["nsa"].astype(np.int64), g["nsa"][i])
    probs, q, full = mcts.getActionProb(g["root"][0], temp=0, force_full_search=True)
    assert sum(probs) == 1 and max(probs) == 1
    v = mcts.search(g["root"][1])
    assert v.shape == (n,) and v.dtype == np.float32
    az.MCTS.reset_all_search_trees()
    assert len(mcts.nodes_data) == 0


@pytest.mark.parametrize("n", [2, 3, 4])
def test_device_network_matches_reference_network(golden_dir, n):
    az = _azg()
    g = np.load(os.path.join(golden_dir, f"nnet_n{n}.npz"))
    net = az.SplendorNNetB200(n, state_dict=az.nnet.random_state_dict(n, int(g["seed"])), dtype=torch.float32)
    torch.backends.cuda.matmul.allow_tf32 = False
    pi, v = net(torch.from_numpy(g["state"]).to(net.device), torch.from_numpy(g["valids"]).to(net.device))
    assert np.abs(_np(pi) - g["pi"]).max() < 2e-5 and np.abs(_np(v) - g["v"]).max() < 2e-5     # float32 tolerance
    p1, v1 = net.predict(g["state"][3], g["valids"][3])
    assert np.abs(p1 - g["pi"][3]).max() < 2e-5 and np.abs(v1 - g["v"][3]).max() < 2e-5
    net16 = az.SplendorNNetB200(n, state_dict=az.nnet.random_state_dict(n, int(g["seed"])), dtype=torch.bfloat16)
    pi16, v16 = net16(torch.from_numpy(g["state"]).to(net.device), torch.from_numpy(g["valids"]).to(net.device))
    assert np.abs(_np(pi16) - g["pi"]).max() < 0.08 and np.abs(_np(v16) - g["v"]).max() < 0.08   # bf16 fast mode: loose


def test_search_with_the_real_network_and_dirichlet_sampler():
    """4096-tree search with the torch network and the on-device Dirichlet sampler: budgets are spent exactly, visit
    counts add up, the noise has the right first moment"""
    az = _azg()
    n, T, sims = 2, 512, 64
    env = az.SplendorEnv(n, T, seed=9)
    env.reset(); env.rollout(30, rotate=True)
    roots = env.states()
    net = az.SplendorNNetB200(n, seed=3)
    ar = az.MCTSArena(n, T, node_cap=256, cpuct=1.25, fpu=0.2, dirichlet_alpha=0.3, seed=77)
    simt = torch.full((T,), sims, dtype=torch.int32, device=ar.device)
    flt = torch.full((T,), 3, dtype=torch.uint8, device=ar.device)
    ar.search(roots, simt, net, flt, None)
    ar.check_status()
    st = ar.root_stats()
    ns, nsa = _np(st["ns"]), _np(st["nsa"])
    assert (ns == sims - 1).all() and (nsa.sum(1) == ns).all() and (_np(st["sims_done"]) == sims).all()
    ps = _np(st["ps"])
    assert np.allclose(ps.sum(1), 1.0, atol=1e-5) and (ps >= 0).all()
    probs, q = ar.policy(1.0)
    assert np.allclose(_np(probs).sum(1), 1.0, atol=1e-9)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_fused_network_kernel(golden_dir, n):
    """one-launch evaluator (csrc/spl_nnet.cu, bf16 products / fp32 accumulation) vs the reference network's recorded
    float32 outputs and vs the torch bf16 path; ragged batch sizes. Tolerance: bf16 inference, 0.05 absolute."""
    az = _azg()
    g = np.load(os.path.join(golden_dir, f"nnet_n{n}.npz"))
    sd = az.nnet.random_state_dict(n, int(g["seed"]))
    net = az.FusedSplendorNNet(n, state_dict=sd)
    st = torch.from_numpy(g["state"]).to(net.device); va = torch.from_numpy(g["valids"].astype(np.uint8)).to(net.device)
    pi, v = net(st, va)
    torch.cuda.synchronize()
    pi, v = _np(pi).copy(), _np(v).copy()
    assert np.isfinite(pi).all() and np.isfinite(v).all()
    assert np.allclose(pi.sum(1), 1.0, atol=1e-4) and (pi[~g["valids"]] == 0).all()
    e_pi, e_v = np.abs(pi - g["pi"]).max(), np.abs(v - g["v"]).max()
    assert e_pi < 0.05 and e_v < 0.05, (e_pi, e_v)
    assert np.abs(pi - g["pi"]).mean() < 2e-4          # a layout slip would be far above this
    for B in (1, 15, 17, 33):                          # ragged tails of the 16-leaf tiles
        p2, v2 = net(st[:B].contiguous(), va[:B].contiguous())
        torch.cuda.synchronize()
        assert np.array_equal(_np(p2), pi[:B]) and np.array_equal(_np(v2), v[:B])
    p1, v1 = net.predict(g["state"][5], g["valids"][5])
    assert np.array_equal(p1, pi[5]) and np.array_equal(v1, v[5])


@pytest.mark.parametrize("n", [2, 3])
def test_selfplay_examples_like_coach(n):
    """batched self-play with example recording, driven until games finish; every recorded example is checked against the
    rules oracle (legal mask of its board, policy on legal moves only) and against Coach's end-of-game bookkeeping:
    the winner / score-difference vectors are the final result seen from the example's mover."""
    az = _azg()
    from oracle import pyoracle as po
    T, sims = 96, 12
    eng = az.SelfPlayEngine(n, T, None, sims, seed=5, cpuct=1.0, node_cap=512, prob_full=0.6, ratio_full=3, record_examples=True)
    eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v)
    eng.env.rollout(50 * n - 20, rotate=True)                 # late positions: games end within a few dozen moves
    eng.arena.reset(); eng.examples.cur_player.zero_()        # (seat labels restart here; only relative seats matter)
    finished = 0
    for mv in range(60):
        probs, q, is_full, ended = eng.play_move()
        finished = int(eng.games_finished.item())
        if finished >= 20 and mv >= 25:
            break
    assert finished >= 5
    ex = eng.drain_examples(symmetries=False)
    E = ex["board"].shape[0]
    assert E > 0
    boards, pi, winner, scdiff, valids = [_np(ex[k]) for k in ("board", "pi", "winner", "scdiff", "valids")]
    for e in range(E):
        b = po.Board(n).set_state(boards[e])
        assert np.array_equal(b.valid_moves(0), valids[e].astype(bool))
        assert abs(pi[e].sum() - 1.0) < 1e-5 and (pi[e][~valids[e].astype(bool)] == 0).all()
        assert scdiff[e][0] == 0 and set(np.unique(winner[e])) <= {1.0, -1.0, np.float32(0.01)}
        assert (winner[e] == 1.0).sum() + (winner[e] == np.float32(0.01)).sum() >= 1
    exs = eng.examples   # symmetric variants: identity first, counts per example in 10..1+9+2n
    ex2 = az.examples.expand_symmetries(eng.env, ex)
    assert ex2["board"].shape[0] >= 10 * E and ex2["board"].shape[0] <= (10 + 2 * n) * E
    lst = az.examples.to_coach_format(ex, compress=True)
    assert len(lst) == E


@pytest.mark.parametrize("n", [2, 3])
def test_batched_arena_like_playgames(n):
    """configs[3] of BASELINE.json in small: a pit of two random-init networks with playout-cap randomisation, every game a
    lane. Checks Arena.playGames' bookkeeping (1-2-2-1 seat order, seat-0 result credited to whoever sat there) and that a
    finished lane's result agrees with an oracle replay of nothing but the final state."""
    az = _azg()
    T = 24
    nets = [az.FusedSplendorNNet(n, seed=1), az.FusedSplendorNNet(n, seed=2)]
    pit = az.BatchedArena(n, nets, num_sims=10, seed=3, prob_full=0.5, ratio_full=5, forced_playouts=True, node_cap=256)
    one, two, draws, d = pit.play_games(T)
    assert one + two + draws == T and d["unfinished"] == 0
    r0, ovt = _np(d["result_seat0"]), _np(d["one_vs_two"])
    assert list(ovt[:8]) == [True, False, False, True, True, False, False, True]           # Arena.py:199
    assert set(np.unique(r0)) <= {1.0, -1.0, np.float32(0.01)}
    assert one == int(((r0 == 1) & ovt).sum() + ((r0 == -1) & ~ovt).sum())
    assert two == int(((r0 == -1) & ovt).sum() + ((r0 == 1) & ~ovt).sum())
    assert (_np(d["moves"]) >= 20).all() and (_np(d["moves"]) <= 62 * n + n).all()
    assert d["total_sims"] > 0


def test_selfplay_async_moves():
    """lanes advancing on their own (SelfPlayEngine.tick): every completed move spent exactly its budget, games finish and
    restart, recorded examples stay consistent with the rules oracle; a descend call that yields after 3 edges is
    result-neutral for the search itself (covered by the arena tests) and must not break the bookkeeping here"""
    az = _azg()
    from oracle import pyoracle as po
    n, T, sims = 2, 128, 24
    eng = az.SelfPlayEngine(n, T, None, sims, seed=11, cpuct=1.0, node_cap=512, record_examples=True, max_levels=3, graph_waves=0,
                            gc_reachable=True, clean_every=3, clean_percent=10)    # also: the between-waves cleaning, often
    eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v)
    eng.env.rollout(70, rotate=True)
    eng.start_async()
    for _ in range(60):
        eng.tick(8)
    moves, done_sims = int(eng.moves_completed.item()), int(eng.sims_completed.item())
    assert moves > 4 * T and done_sims == moves * sims          # prob_full = 1: every search has the full budget
    assert int(eng.games_finished.item()) >= 10
    ex = eng.drain_examples(symmetries=False)
    E = ex["board"].shape[0]
    assert E > 50
    boards, pi, winner, scdiff, valids = [_np(ex[k]) for k in ("board", "pi", "winner", "scdiff", "valids")]
    for e in range(0, E, 7):
        b = po.Board(n).set_state(boards[e])
        assert np.array_equal(b.valid_moves(0), valids[e].astype(bool))
        assert abs(pi[e].sum() - 1.0) < 1e-5 and (pi[e][~valids[e].astype(bool)] == 0).all()
        assert scdiff[e][0] == 0 and ((winner[e] == 1.0) | (winner[e] == np.float32(0.01))).any()
    st = eng.arena.root_stats(want_arrays=False)
    assert int(st["status"].max()) == 0 and int(st["cleanings"].sum()) > T and int(st["resets"].sum()) == 0


def test_cleaning_between_waves_is_result_neutral_on_the_device(golden_dir):
    """spl_mcts_clean (exact mode) after every few waves of a fixture search - simulations in flight get re-based - must not
    change a single visit count (the host-simulator test does the same with a one-lane warp)"""
    az = _azg()
    name = "b_n2_forced_noise"
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    T = 4
    ar = az.MCTSArena(n, T, node_cap=4096, cpuct=cpuct, fpu=fpu, max_levels=5)
    dev = ar.device
    for i in range(len(g["ns"])):
        roots = torch.from_numpy(np.repeat(g["root"][i][None], T, 0)).to(dev)
        simt = torch.full((T,), sims, dtype=torch.int32, device=dev)
        flt = torch.full((T,), 3, dtype=torch.uint8, device=dev)
        dirv = torch.from_numpy(np.repeat(g["dir"][i][None], T, 0)).to(dev).contiguous()
        ar.begin(roots, simt, flt, None, dirv)
        ar.select()
        for w in range(sims):
            pi, v = ar.fixed_net()
            ar.expand_select(pi, v, dirv)
            if w % 5 == 2:
                ar.clean(0)                      # threshold 0: every tree compacts, whatever its search is doing
        ar.finish(lambda s, v: ar.fixed_net(s, v), dirv)
        ar.check_status()
        st = ar.root_stats()
        for t in range(T):
            assert np.array_equal(_np(st["nsa"][t]).astype(np.int64), g["nsa"][i]), (i, t)
            assert np.allclose(_np(st["qsa"][t]), g["qsa"][i], rtol=0, atol=1e-12)
        assert int(st["cleanings"].min()) > 10 * (i + 1)


@pytest.mark.parametrize("n,T", [(2, 200), (3, 67)])
def test_overlapped_wave_equals_classic_wave(n, T):
    """spl_mcts_wave_nnet (network on a side stream next to the attach kernel, input rows taken from the rules kernel's
    staging rows) must give bit-identical trees to select -> spl_nnet_forward -> expand: same visit counts, same Q, same
    node counts; ragged tree count, cleaning between waves, several moves with tree reuse, plain launches and a captured graph"""
    az = _azg()
    sims = 48
    env = az.SplendorEnv(n, T, seed=5)
    env.reset(); env.rollout(25 * n, rotate=True)
    net = az.FusedSplendorNNet(n, seed=4)
    net2 = az.FusedSplendorNNet(n, seed=4)
    a = az.MCTSArena(n, T, node_cap=1024, cpuct=1.1, fpu=0.1, max_levels=6)
    b = az.MCTSArena(n, T, node_cap=1024, cpuct=1.1, fpu=0.1, max_levels=6)
    simt = torch.full((T,), sims, dtype=torch.int32, device=a.device)
    simt[::3] = sims // 4                                  # ragged budgets (playout cap)
    graph = None
    for move in range(3):
        roots = env.states().clone()
        a.search(roots, simt, net)                         # classic, lock-step
        b.begin(roots, simt)
        b.select()
        if move < 2:
            for w in range(sims + 8):
                b.wave_nnet(net2)
                if w % 7 == 3:
                    b.clean(0)
        else:                                              # the same call sequence captured once and replayed
            b.wave_nnet(net2)                              # (first call evaluates the classic leaf rows: outside the capture)
            side = torch.cuda.Stream(b.device)
            side.wait_stream(torch.cuda.current_stream(b.device))
            with torch.cuda.stream(side):
                b.wave_nnet(net2)
            torch.cuda.current_stream(b.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(4):
                    b.wave_nnet(net2)
            for _ in range((sims + 8) // 4):
                graph.replay()
        b.finish(net2)
        a.check_status(); b.check_status()
        sa, sb = a.root_stats(), b.root_stats()
        assert torch.equal(sa["nsa"], sb["nsa"]) and torch.equal(sa["qsa"], sb["qsa"]) and torch.equal(sa["ps"], sb["ps"]), move
        assert torch.equal(sa["sims_done"], sb["sims_done"]) and torch.equal(sa["nn_calls"], sb["nn_calls"]) and torch.equal(sa["ns"], sb["ns"])
        assert int(sa["sims_done"].min()) >= sims // 4
        probs, _ = a.policy(1.0)
        env.step(probs.argmax(1).to(torch.int16), player=0, chance="philox", rotate=True)


@pytest.mark.parametrize("n,T", [(2, 7000), (3, 200)])
def test_wave_with_two_rounds_is_only_scheduling(n, T):
    """rounds = 2 in spl_mcts_wave_nnet (a tree whose descent crossed a transposition or a terminal node gets another attach / descend /
    rules pass before the network runs; the lock-step e2e leg of bench.py): bit-identical trees, fewer waves. 7000 trees take the branch
    with the separate rules kernel, 200 trees the one that otherwise fuses the rules step into the descent"""
    az = _azg()
    sims = 40
    env = az.SplendorEnv(n, T, seed=9)
    env.reset(); env.rollout(30 * n, rotate=True)       # late positions: terminal nodes and transpositions are frequent
    net = az.FusedSplendorNNet(n, seed=4)
    a = az.MCTSArena(n, T, node_cap=512, cpuct=1.1, fpu=0.1, pool_nodes=256)
    b = az.MCTSArena(n, T, node_cap=512, cpuct=1.1, fpu=0.1, pool_nodes=256, rounds=2)
    simt = torch.full((T,), sims, dtype=torch.int32, device=a.device)
    waves = []
    for move in range(2):
        roots = env.states().clone()
        for ar in (a, b):
            l0 = ar.launches
            ar.search(roots, simt, net)
            ar.check_status()
            waves.append((ar.launches - l0) / ar.wave_nnet_launches)
        sa, sb = a.root_stats(), b.root_stats()
        assert torch.equal(sa["nsa"], sb["nsa"]) and torch.equal(sa["qsa"], sb["qsa"]) and torch.equal(sa["ps"], sb["ps"]), move
        assert torch.equal(sa["sims_done"], sb["sims_done"]) and torch.equal(sa["nn_calls"], sb["nn_calls"]) and torch.equal(sa["ns"], sb["ns"])
        assert torch.equal(sa["nodes"], sb["nodes"])
        probs, _ = a.policy(1.0)
        env.step(probs.argmax(1).to(torch.int16), player=0, chance="philox", rotate=True)
    assert waves[1] <= waves[0] and waves[3] <= waves[2], waves


def test_train_batches_from_selfplay_examples_stay_on_the_device():
    """N4: the examples batched self-play produced feed the training mini-batch assembly without leaving the GPU"""
    az = _azg()
    n, T = 2, 96
    eng = az.SelfPlayEngine(n, T, None, 12, seed=21, node_cap=256, record_examples=True)
    eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v)
    eng.env.rollout(70, rotate=True)
    eng.examples.cur_player.zero_()
    for _ in range(40):
        eng.play_move()
    ex = eng.drain_examples(symmetries=True)
    E = int(ex["board"].shape[0])
    assert E >= 64
    b = az.TrainBatcher(ex, 64, seed=5)
    out = b.batch()
    for k in ("boards", "valid_actions", "target_pis", "target_vs", "target_scdiffs"):
        assert out[k].is_cuda and out[k].shape[0] == 64
    assert out["boards"].dtype == torch.float32 and out["valid_actions"].dtype == torch.bool and tuple(out["target_scdiffs"].shape) == (64, 31, n)
    ids = out["ids"]
    assert torch.equal(out["boards"], ex["board"][ids].float()) and float(out["target_scdiffs"].sum()) == 64 * n
    sd = (ex["scdiff"][ids].long() + 15).clamp(0, 30)
    assert bool((out["target_scdiffs"].argmax(1) == sd).all())


def test_selfplay_async_graph_replay_path():
    """the configuration bench.py times: captured waves (spl_mcts_wave_nnet under a CUDA graph: programmatic dependent launches,
    attach on the side stream) + the per-tick move logic replayed as a second graph. Invariants: every completed move spent
    exactly its budget, games finish and restart, no tree overflows, the network evaluates about one row per simulation."""
    az = _azg()
    n, T, sims = 2, 256, 32
    net = az.FusedSplendorNNet(n, seed=2)
    eng = az.SelfPlayEngine(n, T, net, sims, seed=13, node_cap=8 * sims, gc_reachable=True, graph_waves=8, max_levels=16,
                            clean_every=6, clean_percent=45, tick_graph=True)
    assert eng.overlap_nnet
    eng.env.rollout(60, rotate=True)
    eng.start_async()
    for _ in range(80):
        eng.tick(8)
    assert eng._tick_graph is not None and eng._graph is not None          # both graphs were captured and replayed
    moves, done_sims = int(eng.moves_completed.item()), int(eng.sims_completed.item())
    assert moves > 8 * T and done_sims == moves * sims
    assert int(eng.games_finished.item()) >= 5
    st = eng.arena.root_stats(want_arrays=False)
    assert int(st["status"].max()) == 0 and int(st["truncated"].sum()) == 0
    rows = float(st["nn_calls"].sum()) / max(1, done_sims + int(eng.sims_in_flight().item()))
    assert 0.5 < rows < 1.2
    # the boards the engine holds are consistent game states: the rules oracle accepts them and agrees on the legal moves
    from oracle import pyoracle as po
    eng.env.step(None, player=0, store_state=False, want_ended=False, want_status=False)
    boards, valids = _np(eng.env.states()), _np(eng.env.valids())
    for i in range(0, T, 17):
        b = po.Board(n).set_state(boards[i])
        assert np.array_equal(b.valid_moves(0), valids[i].astype(bool))


def test_sample_moves_follows_the_policy():
    """spl_mcts_sample_moves = getActionProb's tail + the caller's np.random.choice (Coach.py:75-86) in one launch: over many
    trees that hold the SAME search (same root, fixed network) but different game ids, the drawn actions follow the probabilities
    spl_mcts_policy returns; unfinished trees give -1; temp 0 gives the most visited action; the counters add up; forced
    playouts prune the same way"""
    az = _azg()
    n, T, sims = 2, 4096, 200
    env = az.SplendorEnv(n, 1, seed=3)
    env.reset(); env.rollout(20, rotate=True)
    root = env.states()[0:1]
    for forced in (0, 1):
        ar = az.MCTSArena(n, T, node_cap=512, cpuct=1.5, fpu=0.1, seed=99)
        roots = root.expand(T, -1, -1).contiguous()
        simt = torch.full((T,), sims, dtype=torch.int32, device=ar.device)
        simt[T // 2:] = sims + 50                                   # the second half will still be searching
        flt = torch.full((T,), forced, dtype=torch.uint8, device=ar.device)
        ar.begin(roots, simt, flt)
        for _ in range(sims):
            ar.wave(lambda s, v: ar.fixed_net(s, v))
        ar.drain_nnet()
        st = ar.root_stats(want_arrays=False)
        assert int(st["sims_done"][: T // 2].min()) == sims
        probs, _ = ar.policy(1.0)
        counters = torch.zeros(2, dtype=torch.int64, device=ar.device)
        episodes = torch.zeros(T, dtype=torch.int32, device=ar.device)
        acts, fin = ar.sample_moves(1.0, episodes, counters=counters)
        acts, fin = _np(acts).astype(np.int64), _np(fin).astype(bool)
        assert fin[: T // 2].all() and not fin[T // 2:].any() and (acts[T // 2:] == -1).all()
        assert [int(x) for x in counters.cpu()] == [sims * (T // 2), T // 2]
        p = _np(probs[0])
        assert (p[acts[: T // 2]] > 0).all()                        # only actions with weight are ever drawn
        freq = np.bincount(acts[: T // 2], minlength=406) / (T // 2)
        assert 0.5 * np.abs(freq - p).sum() < 0.06, 0.5 * np.abs(freq - p).sum()      # total variation, 2048 draws
        a2, _ = ar.sample_moves(1.0, episodes)                      # same key -> same draws; another episode -> other draws
        assert np.array_equal(_np(a2).astype(np.int64), acts)
        a3, _ = ar.sample_moves(1.0, episodes + 1)
        if p.max() < 0.8:                                           # (forced playouts can prune the policy down to one action)
            assert (_np(a3).astype(np.int64)[: T // 2] != acts[: T // 2]).mean() > 0.2
        a0, _ = ar.sample_moves(0.0, episodes)
        nsa = _np(ar.root_stats()["nsa"][0])
        if not forced:
            assert (_np(a0).astype(np.int64)[: T // 2] == int(nsa.argmax())).all()


@pytest.mark.parametrize("K", [2, 4])
def test_virtual_loss_leaf_batching(K):
    """leaves_per_tree > 1 (north star: virtual-loss leaf batching) - an explicit NON-parity option: several simulations of a tree
    per wave, each in flight as a lost visit on the edges it walked. Checked here: every search spends exactly its budget, visit
    counts add up, no virtual visit is left behind, a search needs about 1/K of the waves - and it still finds what the sequential
    search finds (same most visited move in most trees, visit distributions close)."""
    az = _azg()
    n, T, sims = 2, 256, 200
    env = az.SplendorEnv(n, T, seed=17)
    env.reset(); env.rollout(24, rotate=True)
    roots = env.states()
    simt = torch.full((T,), sims, dtype=torch.int32, device=roots.device)
    simt[::5] = 37                                                     # ragged budgets, not multiples of K
    def run(ar):
        ar.begin(roots, simt)
        waves, early = 0, 0
        while True:
            ar.counters.zero_()
            ar.select(count=True)
            leaves, left = [int(x) for x in ar.counters.cpu()]
            if left == 0:
                return waves, early
            if 4 <= waves < 30:
                early += leaves                                        # rows handed to the network per wave while every tree is still searching
            pi, v = ar.fixed_net()
            ar.expand(pi, v)
            waves += 1
    ref = az.MCTSArena(n, T, node_cap=1024, cpuct=1.5, fpu=0.2)
    waves1, early1 = run(ref)                                          # the sequential search, driven the same way
    ar = az.MCTSArena(n, T, node_cap=1024, cpuct=1.5, fpu=0.2, leaves_per_tree=K)
    assert ar.leaf_states.shape[0] == T * K
    waves, early = run(ar)
    ar.check_status()
    a, b = ref.root_stats(), ar.root_stats()
    assert torch.equal(b["sims_done"], simt)
    nsa = b["nsa"]
    assert int(nsa.max()) < (1 << 24) and torch.equal(nsa.sum(1), b["ns"])          # no virtual visit left in any counter
    assert torch.equal(b["ns"], a["ns"])                                            # the same number of real visits as the sequential search
    probs, _ = ar.policy(1.0)
    assert np.allclose(_np(probs).sum(1), 1.0, atol=1e-9)
    assert waves < waves1 and waves >= sims // K and early > (1.5 if K == 2 else 2.2) * early1, (waves, waves1, early, early1)
    pa, pb = _np(a["nsa"]).astype(np.float64), _np(nsa).astype(np.float64)
    pa, pb = pa / pa.sum(1, keepdims=True), pb / pb.sum(1, keepdims=True)
    tv = 0.5 * np.abs(pa - pb).sum(1)
    agree = (pa.argmax(1) == pb.argmax(1)).mean()
    print(f"leaves_per_tree {K}: {waves} waves for {sims} simulations ({waves1} with one leaf per tree; network rows per wave x{early / early1:.2f}), top move agrees in {agree:.2%} of the trees, mean TV distance {tv.mean():.3f}")
    assert agree > 0.6 and tv.mean() < 0.3
    # a second move on the same trees (reuse), then the whole thing once more through the fused network and the one-call wave
    env.step(probs.argmax(1).to(torch.int16), player=0, chance="philox", rotate=True)
    ar.search(env.states(), simt, lambda s, v: ar.fixed_net(s, v))
    ar.check_status()
    assert torch.equal(ar.root_stats()["sims_done"], simt)
    net = az.FusedSplendorNNet(n, seed=4)
    ar.search(env.states(), simt, net)
    ar.check_status()
    st = ar.root_stats()
    assert torch.equal(st["sims_done"], simt) and int(st["nsa"].max()) < (1 << 24)


def test_selfplay_async_with_virtual_loss():
    az = _azg()
    n, T, sims = 2, 256, 64
    net = az.FusedSplendorNNet(n, seed=2)
    eng = az.SelfPlayEngine(n, T, net, sims, seed=13, node_cap=16 * sims, graph_waves=8, max_levels=16, tick_graph=True, leaves_per_tree=2)
    eng.env.rollout(40, rotate=True)
    eng.start_async()
    for _ in range(60):
        eng.tick(8)
    moves, done_sims = int(eng.moves_completed.item()), int(eng.sims_completed.item())
    assert moves > 6 * T and done_sims == moves * sims
    st = eng.arena.root_stats(want_arrays=False)
    assert int(st["status"].max()) == 0 and int(st["truncated"].sum()) == 0
