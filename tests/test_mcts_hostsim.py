"""Tree-arena logic (csrc/spl_mcts.cuh) compiled for the host with a one-lane warp vs the golden fixtures produced by
the reference's own MCTS.py and vs the C search oracle. This is how the CUDA tree code's *semantics* are checked in the
GPU-less container; tests/test_gpu_mcts.py repeats it through the real kernels and the C ABI.
Bar: visit counts, node count and network-call count exact; Qsa within 1e-12 (the spec asks 1e-6); Qs exact.
"""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.hostsim import sim as hs

SCEN = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mcts_*.npz")))


def test_fixed_network_matches():
    rng = np.random.default_rng(5)
    for n in (2, 3, 4):
        b = po.Board(n); b.init_philox(11, n)
        for _ in range(40):
            v = b.valid_moves(0)
            p0, v0 = po.fake_predict(b.state, v, n)
            p1, v1 = hs.fixed_net(b.state, v, n)
            assert np.array_equal(p0, p1) and np.array_equal(v0, v1)
            b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -1); b.swap_players(1)


def _check(out, g, i, name, nodes_exact=True):
    assert out["status"] == 0
    assert np.array_equal(out["nsa"], g["nsa"][i]), (name, i, "visit counts")
    assert out["ns"] == g["ns"][i] and out["nn_calls"] == g["nn_calls"][i]
    if nodes_exact:
        assert out["nodes"] == g["nodes"][i]
    assert np.allclose(out["qsa"], g["qsa"][i], rtol=0, atol=1e-12), (name, i, np.abs(out["qsa"] - g["qsa"][i]).max())
    assert out["qs"] == g["qs"][i]
    assert np.allclose(out["probs"], g["probs"][i], rtol=0, atol=1e-12)
    assert np.allclose(out["q"], g["q"][i], rtol=0, atol=1e-12)
    if "ps" in g:
        # the root's priors after the temperature softmax and the noise. The reference compiles `softmax` / `normalise` with
        # fastmath=True (MCTS.py:238,244): their sums may be re-associated, so the last float32 bit is not defined by the
        # source; we sum in action order and hold the priors to 2 float32 ulps (visit counts and Q above stay exact)
        assert np.abs(out["ps"] - g["ps"][i]).max() <= 2.4e-7, (name, i, np.abs(out["ps"] - g["ps"][i]).max())


@pytest.mark.parametrize("name", SCEN)
def test_tree_reproduces_reference_mcts(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    temp0 = float(g["temp0"]) if "temp0" in g else 1.0          # args.temperature[0]: root softmax before the noise
    m = hs.TreeSim(n, sims, cpuct=cpuct, fpu=fpu, forced_playouts=bool(forced), dirichlet_noise=bool(noise), ratio_full=ratio,
                   temperature0=temp0)
    for i in range(len(g["ns"])):
        d = g["dir"][i] if g["dir_len"][i] > 0 else (np.zeros(406) if noise else None)
        out = m.get_action_prob(g["root"][i], temp=1.0, full_search=bool(g["full"][i]), dir_values=d)
        _check(out, g, i, name)
        assert out["compactions"] == 0 and out["resets"] == 0


@pytest.mark.parametrize("n,seed,cap", [(2, 1, 8192), (2, 2, 8192), (3, 3, 8192), (4, 4, 8192), (2, 5, 2600), (3, 6, 4200), (4, 7, 3400)])
def test_tree_vs_search_oracle_random_games(n, seed, cap):
    """longer searches than the fixtures, mid-game roots with Philox reveals between moves (fresh roots + reuse).
    A revealed card retires the whole tree (its nodes carry a deck no later state has) and the small node limits force the
    copying cleaning (ply + deck rule) before several moves: both must be result-neutral (the search oracle never cleans)."""
    rng = np.random.default_rng(seed)
    sims = 500
    kw = dict(cpuct=1.0 + 0.5 * seed, fpu=0.1 * (seed % 3), forced_playouts=bool(seed & 1), dirichlet_noise=bool(seed & 2), ratio_full=4)
    mo = po.MCTSOracle(n, sims, **kw)
    mt = hs.TreeSim(n, sims, cap=cap, **kw)
    b = po.Board(n); b.init_philox(77, seed)
    for _ in range(12 + 4 * n):    # random opening
        v = b.valid_moves(0)
        b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 77, seed, 0); b.swap_players(1)
    for mv in range(10):
        if b.check_end_game().any():
            break
        full = bool(rng.random() < 0.7)
        k = int(b.valid_moves(0).sum())
        d = np.zeros(406); d[:k] = rng.dirichlet([0.3] * k)
        o = mo.get_action_prob(b.state, temp=1.0, full_search=full, dir_values=d)
        t = mt.get_action_prob(b.state, temp=1.0, full_search=full, dir_values=d)
        assert t["status"] == 0
        assert np.array_equal(o["nsa"], t["nsa"]), (mv, "visit counts")
        assert o["ns"] == t["ns"] and mo.nn_calls == t["nn_calls"] and t["resets"] == 0
        # exact cleanings only drop what the reference's dictionary can never return again: live + dropped = its size
        assert mo.num_nodes == t["nodes"] + t["dropped"], (mv, mo.num_nodes, t["nodes"], t["dropped"])
        assert np.allclose(o["qsa"], t["qsa"], rtol=0, atol=1e-12) and o["qs"] == t["qs"]
        assert np.allclose(o["probs"], t["probs"], rtol=0, atol=1e-12) and np.allclose(o["q"], t["q"], rtol=0, atol=1e-12)
        a = int(np.argmax(o["nsa"]))
        b.make_move(a, 0, -2 if mv % 2 else -1, 77, seed, 0); b.swap_players(1)
    if cap < 8192:
        assert t["compactions"] > 0


@pytest.mark.parametrize("n,seed", [(2, 11), (3, 12), (4, 13)])
def test_reachable_gc_keeps_a_consistent_tree(n, seed):
    """production cleaning (keep only what the new root reaches) with a pool of a few moves: every search must still be a
    complete, self-consistent search (budget spent, counts add up, no lossy reset) and the carried-over subtree must
    keep its statistics (root Ns before the move's simulations >= the chosen child's visit count - 1 when no card was revealed)."""
    rng = np.random.default_rng(seed)
    sims, CAP = 400, 2400
    mt = hs.TreeSim(n, sims, cap=CAP, cpuct=1.5, fpu=0.2, gc_reachable=True)
    b = po.Board(n); b.init_philox(99, seed)
    for _ in range(10 + 4 * n):
        v = b.valid_moves(0)
        b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 99, seed, 0); b.swap_players(1)
    prev_child_n, carried = None, 0
    for mv in range(14):
        if b.check_end_game().any():
            break
        t = mt.get_action_prob(b.state, temp=1.0, full_search=True)
        assert t["status"] == 0 and t["resets"] == 0 and t["nodes"] <= CAP
        assert t["nsa"].sum() == t["ns"] and abs(t["probs"].sum() - 1.0) < 1e-12
        if prev_child_n is not None and prev_child_n > 1:       # reused root: its old visits are still there
            assert t["ns"] >= sims + prev_child_n - 1     # (>: a transposition can feed the node from other parents)
            carried += 1
        else:
            assert t["ns"] in (sims - 1, sims)
        a = int(np.argmax(t["nsa"]))
        before = b.state.copy()
        b.make_move(a, 0, -1, 99, seed, 0)          # no reveal: the new root is a node of the tree
        prev_child_n = int(t["nsa"][a])
        b.swap_players(1)
    assert carried >= 5 and t["compactions"] >= 1


@pytest.mark.parametrize("name,every", [("a_n2_plain", 7), ("b_n2_forced_noise", 13), ("c_n3_fpu", 5), ("d_n4_forced", 11), ("f_n2_late", 3)])
def test_cleaning_in_the_middle_of_a_search_is_result_neutral(golden_dir, name, every):
    """the periodic cleaning (exact mode: keep ply >= root ply) compacts the pools while a simulation is in flight - after a
    descent stopped at a pending edge, after an attach produced a leaf, between yields - and re-bases root / cur / leaf /
    pending edge / path. The search must still reproduce the reference's fixtures exactly."""
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    m = hs.TreeSim(n, sims, cpuct=cpuct, fpu=fpu, forced_playouts=bool(forced), dirichlet_noise=bool(noise), ratio_full=ratio)
    m.set_clean(every)
    for i in range(len(g["ns"])):
        d = g["dir"][i] if g["dir_len"][i] > 0 else (np.zeros(406) if noise else None)
        out = m.get_action_prob(g["root"][i], temp=1.0, full_search=bool(g["full"][i]), dir_values=d)
        _check(out, g, i, name, nodes_exact=False)
    assert out["compactions"] > 50


def test_reachable_cleaning_in_the_middle_of_a_search_keeps_the_tree_consistent():
    n, sims = 2, 300
    rng = np.random.default_rng(3)
    mt = hs.TreeSim(n, sims, cap=4096, cpuct=1.5, fpu=0.2)
    mt.set_clean(9, gc_reachable=True)
    b = po.Board(n); b.init_philox(5, 1)
    for _ in range(20):
        v = b.valid_moves(0)
        b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 5, 1, 0); b.swap_players(1)
    prev = None
    for mv in range(8):
        t = mt.get_action_prob(b.state, temp=1.0, full_search=True)
        assert t["status"] == 0 and t["nsa"].sum() == t["ns"] and abs(t["probs"].sum() - 1.0) < 1e-12
        if prev is not None and prev > 1:
            assert t["ns"] >= sims + prev - 1            # the carried-over subtree kept its statistics
        a = int(np.argmax(t["nsa"])); prev = int(t["nsa"][a])
        b.make_move(a, 0, -1, 5, 1, 0); b.swap_players(1)


@pytest.mark.parametrize("n", [2, 3])
def test_page_pool_accounting(n):
    """every page is either in the free ring or owned by the tree; cleanings (exact, in the middle of searches) give the old
    pages back; reset returns everything"""
    rng = np.random.default_rng(20 + n)
    sims = 220
    mt = hs.TreeSim(n, sims, cap=1500, cpuct=1.3, fpu=0.1)
    free0, total = mt.free_pages()
    assert free0 == total
    mt.set_clean(17)
    b = po.Board(n); b.init_philox(8, n)
    for _ in range(18):
        v = b.valid_moves(0)
        b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 8, n, 0); b.swap_players(1)
    lo = total
    for mv in range(8):
        t = mt.get_action_prob(b.state, temp=1.0, full_search=True)
        assert t["status"] == 0 and t["resets"] == 0
        free, _ = mt.free_pages()
        assert free + t["pages"] == total, (free, t["pages"], total)
        lo = min(lo, free)
        a = int(np.argmax(t["nsa"]))
        b.make_move(a, 0, -2 if mv % 3 == 2 else -1, 8, n, 0); b.swap_players(1)
    assert lo < total and t["compactions"] > 0
    mt.reset()
    assert mt.free_pages() == (total, total)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_compact_state_codec_is_lossless(golden_dir, n):
    """node records hold an 80 / 96 / 128-byte form of the state: every state of the reference's own trajectories (all rotations,
    deterministic and revealing moves, n = 3 / 4 with the mis-rotated noble rows) comes back byte for byte, different states get
    different compact forms, and states the form cannot hold are refused instead of aliased"""
    g = np.load(os.path.join(golden_dir, f"traj_n{n}.npz"))
    states = list(g["state"][:: max(1, len(g["state"]) // 400)]) + list(g["init_state"])
    b = po.Board(n); b.init_philox(3, n)
    rng = np.random.default_rng(n)
    for _ in range(60 * n):                                    # canonical-resident play: rotations after every move
        if b.check_end_game().any():
            break
        v = b.valid_moves(0)
        nxt = b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2 if rng.random() < 0.7 else -1, 3, n, 0); b.swap_players(nxt)
        states.append(b.state.copy())
    seen = {}
    for st in states:
        ok, back, cst = hs.codec_roundtrip(st, n)
        assert ok and np.array_equal(back, st)
        key = cst.tobytes()
        assert key not in seen or np.array_equal(seen[key], st)
        seen[key] = st
    bad = states[5].copy(); bad[3, 2] += 1                     # a cost row that is no card of the tables
    assert not hs.codec_roundtrip(bad, n)[0]
    bad = states[5].copy(); bad[25, 0] += 1                    # a deck count that is not the popcount of its mask
    assert not hs.codec_roundtrip(bad, n)[0]
    bad = states[5].copy(); bad[32 + n, 6] = 1                 # a cell the rules never write
    assert not hs.codec_roundtrip(bad, n)[0]
