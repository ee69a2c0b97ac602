"""Parity of the sm_100a environment kernels, called through the C ABI (include/splendor_b200.h), against
(a) the golden fixtures frozen from the reference and (b) the CPU oracle on the same seeded inputs.
Bar: bit-exact (all rule arithmetic is int8/integer; end vectors are exact float constants).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402


def _az():
    import azg_b200
    return azg_b200


def unpack_words(m):
    """uint32[13, L] mask planes -> bool[L, 406]"""
    m = np.ascontiguousarray(m.T).astype(np.uint32)
    return np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")[:, :406].astype(np.bool_)


def t_i16(x, dev):
    return torch.tensor(np.asarray(x, dtype=np.int16), device=dev)


def t_u8(x, dev):
    return torch.tensor(np.asarray(x, dtype=np.uint8), device=dev)


@pytest.mark.parametrize("use_tma", [True, False])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_trajectories(golden_dir, n, use_tma):
    """every game of the reference-generated fixture replayed as one lane: same deals, same actions, same reveals"""
    az = _az()
    nat = az._native
    g = np.load(os.path.join(golden_dir, f"traj_n{n}.npz"))
    off = g["offsets"]
    G = len(off) - 1
    groups = {}
    for gi in range(G):
        groups.setdefault((int(g["token_limit"][gi]), int(g["reserve"][gi])), []).append(gi)
    checked = 0
    for (limit, reserve), games in groups.items():
        L = len(games)
        flags = nat.RULE_GIVEBACK | nat.RULE_REFCOMPAT | (nat.RULE_RESERVE if reserve else 0)
        env = az.SplendorEnv(n, L, token_limit=limit, rule_flags=flags, use_tma=use_tma)
        dev = env.device
        nob = np.zeros((L, 5), dtype=np.uint8)
        nob[:, : n + 1] = g["nobles"][games]
        env.reset_explicit(g["deals"][games].astype(np.uint8), nob)
        assert np.array_equal(env.states().cpu().numpy(), g["init_state"][games])
        lens = [off[gi + 1] - off[gi] for gi in games]
        for ply in range(max(lens)):
            live = [k for k in range(L) if ply < lens[k]]
            idx = [off[games[k]] + ply for k in live]
            players = np.zeros(L, dtype=np.uint8); players[live] = g["player"][idx]
            # mask before the move, for the mover
            env.step(None, players=t_u8(players, dev), store_state=False, want_ended=False)
            masks = unpack_words(env.masks.cpu().numpy())[:L]
            want = np.unpackbits(g["mask"][idx], axis=1, bitorder="little")[:, :406].astype(bool)
            assert np.array_equal(masks[live], want), (n, ply)
            # the packed-bits export agrees with the word planes
            assert np.array_equal(env.valids().cpu().numpy().astype(bool), masks)
            actions = np.full(L, -1, dtype=np.int16); actions[live] = g["action"][idx]
            reveals = np.full(L, 255, dtype=np.uint8)
            for k, i in zip(live, idx):
                if not g["det"][i] and g["reveal"][i] >= 0:
                    reveals[k] = g["reveal"][i]
            # a replayed move with no recorded reveal and det=0 means "nothing was drawn" (deck empty / no draw)
            env.step(t_i16(actions, dev), players=t_u8(players, dev), chance="replay", reveals=t_u8(reveals, dev), want_mask=False)
            st = env.states().cpu().numpy()
            assert np.array_equal(st[live], g["state"][idx]), (n, ply)
            assert np.array_equal(env.ended.cpu().numpy()[live], g["ended"][idx])
            assert np.array_equal(env.status.cpu().numpy()[live], (players[live] + 1) % n)
            sc, rd = env.scores()
            assert np.array_equal(sc.cpu().numpy()[live], g["score"][idx])
            assert np.array_equal(rd.cpu().numpy()[live], g["state"][idx][:, 0, 6].view(np.uint8))
            checked += len(live)
    assert checked == len(g["action"])


@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_synthetic(golden_dir, n):
    """perturbed states (token regimes, empty banks, end-of-game ties incl. the n>=3 quirks): mask, end vector, rotations"""
    az = _az()
    g = np.load(os.path.join(golden_dir, f"synth_n{n}.npz"))
    L = len(g["state"])
    env = az.SplendorEnv(n, L)
    dev = env.device
    env.set_states(g["state"])
    assert np.array_equal(env.states().cpu().numpy(), g["state"])       # pack -> unpack round trip
    env.step(None, players=t_u8(g["player"], dev), store_state=False)
    want = np.unpackbits(g["mask"], axis=1, bitorder="little")[:, :406].astype(bool)
    assert np.array_equal(unpack_words(env.masks.cpu().numpy())[:L], want)
    assert np.array_equal(env.ended.cpu().numpy(), g["ended"])
    assert np.array_equal(env.scores()[0].cpu().numpy(), g["score"])
    for k in range(1, n):
        env.set_states(g["state"])
        env.step(None, player=k, rotate=True, want_mask=False)
        assert np.array_equal(env.states().cpu().numpy(), g["rot"][:, k - 1]), k


@pytest.mark.parametrize("n", [2, 3, 4])
def test_symmetries_golden(golden_dir, n):
    az = _az()
    g = np.load(os.path.join(golden_dir, f"sym_n{n}.npz"))
    B = len(g["state"])
    env = az.SplendorEnv(n, B)
    os_, op, ov, cnt = env.symmetries(g["state"], g["pi"], g["valids"].astype(np.uint8))
    os_, op, ov, cnt = os_.cpu().numpy(), op.cpu().numpy(), ov.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, g["count"])
    o = 0
    for b in range(B):
        for v in range(cnt[b]):
            assert np.array_equal(os_[b, v], g["out_state"][o]), (b, v)
            assert np.array_equal(op[b, v], g["out_pi"][o]), (b, v)
            assert np.array_equal(ov[b, v].astype(bool), g["out_valids"][o]), (b, v)
            o += 1
    assert o == len(g["out_state"])


@pytest.mark.parametrize("n,compat,rotate", [(2, True, True), (3, True, True), (4, True, True), (3, False, False), (4, False, True), (2, True, False)])
def test_philox_steps_vs_oracle(n, compat, rotate):
    """lock-step with the oracle under Philox chance: init, mask, pick, move + reveal, rotation, end vector,
    auto reset into the next episode - 96 lanes (ragged: 3 tiles), ~1.5 games each"""
    az = _az()
    nat = az._native
    L, seed, base = 96 - 7, 0xABCDEF0123 + n, 1000
    flags = nat.RULE_RESERVE | nat.RULE_GIVEBACK | (nat.RULE_REFCOMPAT if compat else 0)
    env = az.SplendorEnv(n, L, seed=seed, game_base=base, rule_flags=flags)
    env.reset()
    env.step(None, want_next=True, want_ended=False)
    boards, players, eps = [], [0] * L, [0] * L
    for gidx in range(L):
        b = po.Board(n, ref_compat=compat)
        b.init_philox(seed, base + gidx, 0)
        boards.append(b)
    finished = 0
    for ply in range(int(62 * n * 1.5)):
        st = env.states().cpu().numpy()
        va = unpack_words(env.masks.cpu().numpy())
        acts = env.next_actions.cpu().numpy()
        pl = env.status.cpu().numpy() if (not rotate and ply > 0) else None
        for gidx, b in enumerate(boards):
            assert np.array_equal(st[gidx], b.state), ("state", n, ply, gidx)
            p = players[gidx]
            vo = b.valid_moves(p)
            assert np.array_equal(va[gidx], vo), ("mask", n, ply, gidx)
            a = po.philox_pick(vo, seed, base + gidx, eps[gidx], b.get_round())
            assert a == int(acts[gidx]), ("pick", n, ply, gidx)
            nxt = b.make_move(a, p, -2, seed, base + gidx, eps[gidx])
            assert nxt == (p + 1) % n
            if rotate:
                b.swap_players(nxt)
                players[gidx] = 0
            else:
                players[gidx] = nxt
        pt = None if rotate else t_u8([(p - 1) % n for p in players], env.device)   # the movers of this ply
        env.step(env.next_actions.clone(), players=pt, player=0, chance="philox", rotate=rotate, auto_reset=True,
                 want_next=True, count=True)
        en = env.ended.cpu().numpy()
        stat = env.status.cpu().numpy()
        for gidx, b in enumerate(boards):
            e = b.check_end_game()
            assert np.array_equal(en[gidx], e), ("ended", n, ply, gidx)
            if not rotate:
                assert stat[gidx] == players[gidx]
            if e.any():
                finished += 1
                eps[gidx] += 1
                b.init_philox(seed, base + gidx, eps[gidx])
                players[gidx] = 0
        if not rotate:
            # after an auto reset the player to move is 0; the kernel computed the next mask for that player
            pass
    assert finished >= L
    cnt = env.counters.cpu().numpy()
    assert cnt[0] == finished and cnt[1] == L * int(62 * n * 1.5)
    assert np.array_equal(env.episodes.cpu().numpy(), np.array(eps))


@pytest.mark.parametrize("use_tma", [True, False])
@pytest.mark.parametrize("n,rotate,compat", [(2, False, True), (2, True, True), (3, False, True), (4, False, True), (4, True, False), (3, True, False)])
def test_rollout_kernel_vs_oracle(n, rotate, compat, use_tma):
    """the persistent multi-ply kernel plays whole games; plies and results of every lane's first game equal spo_rollout's
    (which plays in the absolute frame). Games end only when player 0 is to move, so the canonical-rotating kernel gives
    the same vectors - except under ref_compat with n>=3, where the reference's stride-3 noble rotation (F7a) is not a
    true rotation; that combination is covered ply by ply in test_philox_steps_vs_oracle."""
    az = _az()
    nat = az._native
    L, seed, base = 1000, 77 + n, 5
    flags = nat.RULE_RESERVE | nat.RULE_GIVEBACK | (nat.RULE_REFCOMPAT if compat else 0)
    env = az.SplendorEnv(n, L, seed=seed, game_base=base, use_tma=use_tma, rule_flags=flags)
    env.reset()
    fp = torch.zeros(L, dtype=torch.int32, device=env.device)
    fr = torch.zeros((L, n), dtype=torch.float32, device=env.device)
    K = 62 * n + 2
    env.rollout(K, rotate=rotate, first_plies=fp, first_result=fr)
    total, plies, res = po.rollout(n, seed, base, L, ref_compat=compat)
    assert np.array_equal(fp.cpu().numpy(), plies)
    assert np.array_equal(fr.cpu().numpy(), res)
    cnt = env.counters.cpu().numpy()
    assert cnt[1] == L * K and cnt[0] >= L
    # split into two launches: identical final state (the kernel is a pure function of (state, episode, player))
    env2 = az.SplendorEnv(n, L, seed=seed, game_base=base, use_tma=use_tma, rule_flags=flags)
    env2.reset()
    env2.rollout(K // 2, rotate=rotate)
    env2.rollout(K - K // 2, rotate=rotate)
    assert torch.equal(env.planes, env2.planes) and torch.equal(env.episodes, env2.episodes)
    assert np.array_equal(env2.counters.cpu().numpy(), cnt)


def test_rollout_equals_single_steps():
    """K plies in one persistent launch == K single-ply launches fed with their own random picks"""
    az = _az()
    n, L, K = 2, 4096 + 5, 50
    a = az.SplendorEnv(n, L, seed=9); a.reset(); a.rollout(K, rotate=True)
    b = az.SplendorEnv(n, L, seed=9); b.reset()
    b.step(None, want_next=True)
    for _ in range(K):
        b.step(b.next_actions, player=0, chance="philox", rotate=True, auto_reset=True, want_next=True, count=True)
    assert torch.equal(a.planes, b.planes)
    assert torch.equal(a.counters, b.counters)


@pytest.mark.parametrize("L", [1, 31, 32, 33, 1000])
def test_pack_unpack_ragged(L):
    az = _az()
    n = 3
    rng = np.random.default_rng(L)
    x = rng.integers(-128, 128, size=(L, az.rows(n), 7), dtype=np.int8)
    env = az.SplendorEnv(n, L)
    env.set_states(x)
    assert np.array_equal(env.states().cpu().numpy(), x)
    # padding lanes of the last tile are zero
    S = env.S
    tiles = env.planes.view(-1, S, 32).cpu().numpy()
    if L % 32:
        assert not tiles[-1][:, L % 32:].any()


def test_full_size_invariants():
    """BASELINE.json's size (2p, 64k lanes and the 1M-lane sweep point): size-independent properties -
    gem conservation per colour, deck bitmask/count agreement, ply accounting, determinism across launches"""
    az = _az()
    n = 2
    for L in (65536, 1 << 20):
        env = az.SplendorEnv(n, L, seed=31337)
        env.reset()
        env.rollout(64, rotate=True)
        first = env.planes.clone()
        st = env.states()
        bank = st[:, 0, :6].to(torch.int32)
        pg = st[:, 32 + n: 32 + 2 * n, :6].to(torch.int32).sum(1)
        tot = bank + pg
        assert bool((tot[:, :5] == 4).all()) and bool((tot[:, 5] == 5).all())        # gems only move between bank and players
        cnt = st[:, 25:31:2, :5].to(torch.int32)
        bits = st[:, 26:32:2, :5].to(torch.int32) & 0xFF
        pop = torch.zeros_like(bits)
        for k in range(8):
            pop += (bits >> k) & 1
        assert bool((pop == cnt).all())                                              # deck counters == popcount of the bitmask
        c = env.counters.cpu().numpy()
        assert c[1] == L * 64
        env2 = az.SplendorEnv(n, L, seed=31337); env2.reset(); env2.rollout(64, rotate=True)
        assert torch.equal(first, env2.planes)
        del env, env2, st, first


def test_game_api_dropin():
    """the Game protocol mirror driven the way Coach.executeEpisode / Arena.playGame drive the reference"""
    az = _az()
    for n in (2, 3):
        game = az.SplendorGame(n, seed=4242)
        board = game.getInitBoard()
        assert board.shape == game.getBoardSize() and board.dtype == np.int8 and game.getActionSize() == 406
        ob = po.Board(n); ob.init_philox(4242, 0, 0)
        assert np.array_equal(board, ob.state)
        cur, plies = 0, 0
        rng = np.random.default_rng(0)
        while True:
            canonical = game.getCanonicalForm(board, cur)
            if cur == 0:
                assert canonical is board
            oc = ob.copy(); oc.swap_players(cur)
            assert np.array_equal(canonical, oc.state)
            valids = game.getValidMoves(canonical, 0)
            assert valids.dtype == np.bool_ and np.array_equal(valids, oc.valid_moves(0))
            assert np.array_equal(valids, game.getValidMoves(board, cur))
            a = int(rng.choice(np.flatnonzero(valids)))
            prev = board
            board, nxt = game.getNextState(board, cur, a)
            assert np.array_equal(prev, ob.state)             # the input array is not mutated
            assert ob.make_move(a, cur, -2, 4242, 0, 0) == nxt
            assert np.array_equal(board, ob.state)
            assert game.getScore(board, 0) == ob.get_score(0) and game.getRound(board) == ob.get_round()
            assert game.stringRepresentation(board) == ob.state.tobytes()
            r = game.getGameEnded(board, nxt)
            assert r.dtype == np.float32 and np.array_equal(r, ob.check_end_game())
            cur = nxt; plies += 1
            if r.any():
                break
        assert plies > 20
        pi = rng.random(406).astype(np.float32)
        syms = game.getSymmetries(board, pi, valids)
        osyms = ob.symmetries(pi, valids)
        assert len(syms) == len(osyms)
        for (s1, p1, v1), (s2, p2, v2) in zip(syms, osyms):
            assert np.array_equal(s1, s2) and np.array_equal(p1, p2) and np.array_equal(v1, v2)
        # deterministic (in-tree) step and the batched host-buffer form
        b2, _ = game.getNextState(game.getInitBoard(), 0, 30 + 15, deterministic=True)
        boards = np.stack([ob.state] * 5)
        valid0 = np.flatnonzero(ob.valid_moves(cur))
        nb, nv, ne = game.getNextStateBatch(boards, cur, [valid0[0]] * 5, deterministic=True)
        o2 = ob.copy(); nx = o2.make_move(int(valid0[0]), cur, -1); o2.swap_players(nx)
        assert all(np.array_equal(nb[i], o2.state) for i in range(5))
        assert all(np.array_equal(nv[i], o2.valid_moves(0)) for i in range(5))
        assert all(np.array_equal(ne[i], o2.check_end_game()) for i in range(5))


def test_no_gpu_fallback_is_loud():
    az = _az()
    import ctypes as C
    h = C.c_void_p()
    assert az._native.lib().spl_ctx_create(2, 10, 7, 99, C.byref(h)) < 0      # bad device -> error code, nothing thrown
    assert az._native.lib().spl_ctx_create(5, 10, 7, 0, C.byref(h)) < 0
    assert b"bad" in az._native.lib().spl_last_error()


@pytest.mark.parametrize("n,L", [(2, 20000), (3, 1000)])
def test_packed_pipelined_batch_call_equals_the_plain_call(n, L):
    """getNextStateBatch(..., packed_masks=True) - lane chunks on their own streams, 52-byte masks - returns what the one-stream call
    with bool[L,406] masks returns, and what the oracle computes, for a ragged lane count (not a multiple of 32 or of the chunks)"""
    az = _az()
    game = az.SplendorGame(n, seed=77)
    src = az.SplendorEnv(n, L, seed=5)
    src.reset(); src.rollout(30, rotate=True); src.step(None, want_next=True)
    boards, acts = src.states().cpu().numpy(), src.next_actions.cpu().numpy()
    b0, v0, e0 = [x.copy() for x in game.getNextStateBatch(boards, 0, acts, deterministic=True)]
    b1, m1, e1 = [x.copy() for x in game.getNextStateBatch(boards, 0, acts, deterministic=True, packed_masks=True)]
    assert m1.dtype == np.uint32 and m1.shape == (L, 13)
    assert np.array_equal(b0, b1) and np.array_equal(e0, e1) and np.array_equal(v0, game.unpack_masks(m1))
    for g in range(0, L, max(1, L // 40)):
        ob = po.Board(n).set_state(boards[g])
        nxt = ob.make_move(int(acts[g]), 0, -1); ob.swap_players(nxt)
        assert np.array_equal(ob.state, b1[g]) and np.array_equal(ob.valid_moves(0), v0[g]) and np.array_equal(ob.check_end_game(), e1[g])


def test_rollout_batch_round_trip():
    """rolloutBatch: K plies of random legal play per host round trip; the boards that come back are legal game states `plies`
    plies further (or restarted), tokens are conserved, the counters add up"""
    az = _az()
    n, L, K = 2, 5000, 16
    game = az.SplendorGame(n, seed=3)
    src = az.SplendorEnv(n, L, seed=8)
    src.reset(); src.rollout(10, rotate=True)
    boards = src.states().cpu().numpy()
    out, finished, plies = game.rolloutBatch(boards, K)
    out = out.copy()
    assert plies == L * K and 0 <= finished < L
    ply_in, ply_out = boards[:, 0, 6].astype(np.uint8).astype(int), out[:, 0, 6].astype(np.uint8).astype(int)
    assert ((ply_out == ply_in + K) | (ply_out < ply_in + K)).all() and (ply_out == ply_in + K).mean() > 0.9
    gems = out[:, 0, :5].astype(int) + out[:, 34, :5] + out[:, 35, :5]
    assert (gems == 4).all()                                     # 4 gems per colour in a 2-player game, bank + both players
    for g in range(0, L, 250):
        ob = po.Board(n).set_state(out[g])
        assert ob.valid_moves(0).any()
