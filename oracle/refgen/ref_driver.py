"""TEST INFRASTRUCTURE — drives the *patched reference* (numba jitclass Board) and records
everything the parity tests compare against: reveals, masks, states, end values, rotations,
symmetries. Used by gen_golden.py (to freeze fixtures under tests/golden/) and by the live
differential test when /root/reference is present. Never imported by the product.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_patched_ref  # noqa: E402

_ref = {}


def ref():
    """Import the patched reference once; returns a namespace dict."""
    if _ref:
        return _ref
    build_patched_ref.import_ref()
    import warnings
    warnings.filterwarnings("ignore")
    from numba import njit
    from splendor.SplendorLogicNumba import Board
    from splendor import SplendorLogic as L

    @njit
    def seed_numba(k):
        np.random.seed(k)

    _ref.update(Board=Board, L=L, seed_numba=seed_numba)
    return _ref


def card_lookup():
    """(tier, cost+gain bytes) -> (colour, idx)"""
    L = ref()["L"]
    lut = {}
    for t, a in enumerate([L.np_all_cards_1, L.np_all_cards_2, L.np_all_cards_3]):
        for c in range(a.shape[0]):
            for i in range(a.shape[1]):
                lut[(t, a[c, i].tobytes())] = (c, i)
    return lut


def noble_lookup():
    L = ref()["L"]
    return {L.np_all_nobles[i].tobytes(): i for i in range(len(L.np_all_nobles))}


def deck_diff(before, after):
    """returns colour*8+idx of the card drawn between two states, or -1"""
    out = -1
    for t in range(3):
        rb, ra = before[25 + 2 * t + 1, :5].astype(np.uint8), after[25 + 2 * t + 1, :5].astype(np.uint8)
        for c in range(5):
            d = int(rb[c]) & ~int(ra[c])
            if d:
                assert out == -1 and bin(d).count("1") == 1
                out = c * 8 + (7 - d.bit_length() + 1)
    return out


def initial_setup(state, n):
    """recover (deals[12] as colour*8+idx, nobles[n+1]) from a freshly initialised state"""
    lut, nl = card_lookup(), noble_lookup()
    deals = []
    for slot in range(12):
        t = slot // 4
        c, i = lut[(t, state[1 + 2 * slot: 3 + 2 * slot].tobytes())]
        deals.append(c * 8 + i)
    nobles = [nl[state[31 + k].tobytes()] for k in range(n + 1)]
    return np.array(deals, dtype=np.int16), np.array(nobles, dtype=np.int16)


def play_random_game(n, seed, det_prob=0.0, token_limit=10, reserve=True, rng=None, max_plies=400):
    """One uniformly-random game on the reference Board.

    Returns dict of arrays: deals, nobles, and per ply: action, player, reveal (-1 none),
    det (1 if the move was made with deterministic=True), mask (valid moves BEFORE the move,
    packed little-endian bits, 51 bytes), state (AFTER the move), ended (after), next_player.
    """
    R = ref()
    rng = rng or np.random.default_rng(seed)
    R["seed_numba"](seed)
    b = R["Board"](n)
    b.setNumTokenLim(token_limit)
    b.ENABLE_ACTION_RESERVE = reserve
    init_state = b.get_state().copy()
    deals, nobles = initial_setup(init_state, n)
    rec = dict(action=[], player=[], reveal=[], det=[], mask=[], state=[], ended=[], score=[])
    player = 0
    for _ in range(max_plies):
        valids = b.valid_moves(player)
        a = int(rng.choice(np.flatnonzero(valids)))
        det = bool(rng.random() < det_prob)
        before = b.get_state().copy()
        b.copy_state(before, True)
        nxt = b.make_move(a, player, det)
        after = b.get_state().copy()
        rec["action"].append(a); rec["player"].append(player); rec["det"].append(int(det))
        rec["reveal"].append(deck_diff(before, after))
        rec["mask"].append(np.packbits(valids, bitorder="little"))
        rec["state"].append(after)
        e = b.check_end_game()
        rec["ended"].append(e.copy())
        rec["score"].append([int(b.get_score(p)) for p in range(n)])
        player = nxt
        if e.any():
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out["action"] = out["action"].astype(np.int16); out["player"] = out["player"].astype(np.int8)
    out["reveal"] = out["reveal"].astype(np.int16); out["det"] = out["det"].astype(np.int8)
    out["score"] = out["score"].astype(np.int16)
    out.update(deals=deals, nobles=nobles, init_state=init_state, n=np.int32(n), token_limit=np.int32(token_limit),
               reserve=np.int32(reserve))
    return out


def rotations(state, n):
    """swap_players(k) for k = 1..n-1 on a copy of `state` (reference :338-347)"""
    R = ref()
    outs = []
    for k in range(1, n):
        b = R["Board"](n)
        b.copy_state(state, True)
        b.swap_players(k)
        outs.append(b.get_state().copy())
    return np.array(outs)


def symmetries(state, n, pi, valids):
    R = ref()
    b = R["Board"](n)
    b.copy_state(state, True)
    syms = b.get_symmetries(np.asarray(pi, dtype=np.float32), np.asarray(valids, dtype=np.bool_))
    return [(s.copy(), p.copy(), v.copy()) for (s, p, v) in syms]


def eval_state(state, n, player, token_limit=10, reserve=True):
    """(valid mask, ended, scores) of an arbitrary state on the reference"""
    R = ref()
    b = R["Board"](n)
    b.setNumTokenLim(token_limit)
    b.ENABLE_ACTION_RESERVE = reserve
    b.copy_state(state, True)
    return b.valid_moves(player), b.check_end_game(), [int(b.get_score(p)) for p in range(n)]
