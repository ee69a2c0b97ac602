"""Device rules core (csrc/spl_rules.cuh) compiled for the host vs the CPU oracle and the golden fixtures.

This is how the CUDA code's *logic* is checked in the GPU-less container; the -m gpu tests repeat the
comparison through the real kernels and the C-ABI. Bar: bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.hostsim import sim as hs


@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_trajectories(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"traj_n{n}.npz"))
    off = g["offsets"]
    for gi in range(len(off) - 1):
        flags = hs.F_GIVEBACK | hs.F_REFCOMPAT | (hs.F_RESERVE if g["reserve"][gi] else 0)
        s = hs.Sim(n, limit=int(g["token_limit"][gi]), flags=flags)
        s.init_explicit(g["deals"][gi], g["nobles"][gi])
        assert np.array_equal(s.state, g["init_state"][gi])
        for i in range(off[gi], off[gi + 1]):
            player = int(g["player"][i])
            assert np.array_equal(np.packbits(s.valid_moves(player), bitorder="little"), g["mask"][i]), (gi, i - off[gi])
            reveal = -1 if g["det"][i] else int(g["reveal"][i])
            assert s.make_move(int(g["action"][i]), player, reveal) == (player + 1) % n
            assert np.array_equal(s.state, g["state"][i]), (gi, i - off[gi], int(g["action"][i]))
            assert np.array_equal(s.check_end_game(), g["ended"][i])
            assert [s.get_score(p) for p in range(n)] == list(g["score"][i])


@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_synthetic(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"synth_n{n}.npz"))
    for i in range(len(g["state"])):
        s = hs.Sim(n).set_state(g["state"][i])
        assert np.array_equal(np.packbits(s.valid_moves(int(g["player"][i])), bitorder="little"), g["mask"][i]), i
        assert np.array_equal(s.check_end_game(), g["ended"][i]), i
        for k in range(1, n):
            r = hs.Sim(n).set_state(g["state"][i]); r.swap_players(k)
            assert np.array_equal(r.state, g["rot"][i][k - 1]), (i, k)


@pytest.mark.parametrize("n,compat", [(2, True), (3, True), (4, True), (3, False), (4, False)])
def test_philox_games_vs_oracle(n, compat):
    """full Philox-driven random games: init, pick, reveal, mask, end, rotation all agree with the oracle"""
    flags = hs.F_RESERVE | hs.F_GIVEBACK | (hs.F_REFCOMPAT if compat else 0)
    seed = 0xC0FFEE1234 + n
    plies = 0
    for game in range(60):
        o = po.Board(n, ref_compat=compat); o.init_philox(seed, game, 3)
        s = hs.Sim(n, flags=flags); s.init_philox(seed, game, 3)
        assert np.array_equal(o.state, s.state)
        player = 0
        while True:
            vo = o.valid_moves(player)
            words = s.valid_words(player)
            assert np.array_equal(vo, hs.unpack_mask(words)), (game, plies)
            ply = o.get_round()
            a = po.philox_pick(vo, seed, game, 3, ply)
            assert a == hs.pick_random(words, seed, game, 3, ply)
            no = o.make_move(a, player, -2, seed, game, 3)
            ns = s.make_move(a, player, -2, seed, game, 3)
            assert no == ns and np.array_equal(o.state, s.state), (game, plies, a)
            eo, es = o.check_end_game(), s.check_end_game()
            assert np.array_equal(eo, es)
            # canonical rotation of a copy
            oc, sc = o.copy(), hs.Sim(n, flags=flags).set_state(s.state)
            oc.swap_players(no); sc.swap_players(ns)
            assert np.array_equal(oc.state, sc.state)
            player = no
            plies += 1
            if eo.any():
                break
    assert plies > 3000


def test_deterministic_tree_steps_vs_oracle():
    """the in-tree MCTS step: make_move(a, 0, deterministic=True) + swap_players(1) chains (MCTS.py:222-237)"""
    rng = np.random.default_rng(5)
    for n in (2, 3, 4):
        for game in range(25):
            o = po.Board(n); o.init_philox(99, game)
            s = hs.Sim(n).set_state(o.state)
            for ply in range(80):
                v = o.valid_moves(0)
                assert np.array_equal(v, s.valid_moves(0))
                a = int(rng.choice(np.flatnonzero(v)))
                o.make_move(a, 0, -1); s.make_move(a, 0, -1)
                o.swap_players(1); s.swap_players(1)
                assert np.array_equal(o.state, s.state), (n, game, ply, a)
                assert np.array_equal(o.check_end_game(), s.check_end_game())
                if o.check_end_game().any():
                    break


def test_generated_fail_masks_match_their_source_tables():
    """SPL_COMBO_WITH / SPL_GIVE3_NEEDS (the compile-time masks spl_valid_mask ORs together for what a player or the bank lacks) are
    derived tables: re-derive them from SPL_COMBO_BITS / SPL_GIVE3 as they stand in the generated header"""
    import os
    import re
    path = os.path.join(os.path.dirname(__file__), "..", "alphazero-general-ori_b200", "csrc", "spl_tables.cuh")
    src = open(path).read()

    def table(name):
        m = re.search(name + r"(?:\[\d+\])+ = \{([^}]*)\}", src)
        assert m, name
        return [int(x.strip().rstrip("ul"), 0) for x in m.group(1).split(",")]
    combo, give3 = table("SPL_COMBO_BITS"), table("SPL_GIVE3")
    assert len(combo) == 25 and len(give3) == 40
    assert table("SPL_COMBO_WITH") == [sum(1 << i for i in range(25) if (combo[i] >> c) & 1) for c in range(5)]
    needs = table("SPL_GIVE3_NEEDS")
    for c in range(5):
        for lvl in range(3):
            assert needs[3 * c + lvl] == sum(1 << i for i in range(40) if ((give3[i] >> (4 * c)) & 15) == lvl + 1)
    # every give-3 pattern hands over exactly three gems
    assert all(sum((p >> (4 * c)) & 15 for c in range(5)) == 3 for p in give3)
