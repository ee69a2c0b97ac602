#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — freezes outputs of the (patched) reference into tests/golden/.

Run in the build container (needs /root/reference + numba):  python oracle/refgen/gen_golden.py
The fixtures are what pins the CPU oracle (and through it the CUDA path) to the reference:

  traj_n{2,3,4}.npz   random full games on the reference Board: initial deal, per-ply action /
                      reveal / valid mask / next state / end vector / scores; a share of the
                      moves is made with deterministic=True (the in-tree MCTS step), some games
                      with NUM_TOKEN_LIMIT=8 and with reserve disabled
  synth_n{2,3,4}.npz  perturbed mid-game states (token regimes 8/9/10, empty banks, end-of-game
                      ties incl. the n>=3 quirks F7a/F7b) with the reference's mask / end / score /
                      rotations
  sym_n{2,3,4}.npz    get_symmetries outputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.realpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import ref_driver as rd  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def pack_games(games):
    """list of per-game dicts -> flat arrays with offsets"""
    keys = ["action", "player", "reveal", "det", "mask", "state", "ended", "score"]
    out = {k: np.concatenate([g[k] for g in games]) for k in keys}
    out["offsets"] = np.cumsum([0] + [len(g["action"]) for g in games]).astype(np.int32)
    for k in ["deals", "nobles", "init_state", "token_limit", "reserve"]:
        out[k] = np.array([g[k] for g in games])
    return out


def gen_traj(n):
    specs = []
    ngames = {2: 14, 3: 8, 4: 6}[n]
    for s in range(ngames):
        specs.append(dict(seed=1000 * n + s, det_prob=0.0))
    for s in range(4):
        specs.append(dict(seed=1000 * n + 100 + s, det_prob=0.35))
    specs.append(dict(seed=1000 * n + 200, det_prob=0.0, token_limit=8))
    specs.append(dict(seed=1000 * n + 201, det_prob=0.1, reserve=False))
    games = [rd.play_random_game(n, **sp) for sp in specs]
    return pack_games(games), games


def perturb(state, n, rng):
    """make a plausible-but-stressful variant of a mid-game state"""
    s = state.copy()
    p = int(rng.integers(n))
    g = 32 + n + p
    mode = rng.integers(6)
    if mode <= 2:
        # force the player's token total into the 8/9/10 regimes
        target = [8, 9, 10][mode]
        gems = np.zeros(6, dtype=np.int64)
        for _ in range(target):
            gems[rng.integers(6 if rng.random() < 0.3 else 5)] += 1
        s[g, :6] = gems
    elif mode == 3:
        # thin bank: 1 or 2 colours left
        keep = rng.choice(5, size=int(rng.integers(1, 3)), replace=False)
        bank = np.zeros(5, dtype=np.int64); bank[keep] = rng.integers(1, 5, size=len(keep))
        s[0, :5] = bank
        s[g, :6] = rng.integers(0, 3, size=6)
    elif mode == 4:
        s[0, 5] = 0  # no gold in bank
        tot = int(rng.integers(7, 11)); gems = np.zeros(6, dtype=np.int64)
        for _ in range(tot):
            gems[rng.integers(5)] += 1
        s[g, :6] = gems
    else:
        # rich player: many bonuses (buy logic / nobles)
        s[32 + 3 * n + n * n + p, :5] = rng.integers(0, 6, size=5)
        s[g, :6] = rng.integers(0, 4, size=6)
    return s, p


def endgame_variant(state, n, rng):
    """final-round states with crafted score ties"""
    s = state.copy()
    s[0, 6] = np.int8(np.uint8(n * int(rng.integers(1, 62))))
    if rng.random() < 0.15:
        s[0, 6] = np.int8(np.uint8(62 * n))     # time-out end
    pc = 32 + 3 * n + n * n
    base = int(rng.integers(11, 18))
    for p in range(n):
        s[pc + p, 6] = base + int(rng.integers(-1, 2)) * int(rng.random() < 0.5)
        s[pc + p, :5] = rng.integers(0, 4, size=5)
        if rng.random() < 0.4:
            s[pc + p, :5] = s[pc, :5]            # same number of cards -> 0.01 draws
    # sprinkle noble points in the players_nobles block (n*(n+1) rows) to exercise the stride quirk
    pn = 32 + 2 * n
    s[pn: pn + n * (n + 1), :] = 0
    for r in range(n * (n + 1)):
        if rng.random() < 0.25:
            s[pn + r, 6] = 3
            s[pn + r, :5] = [3, 3, 3, 0, 0]
    return s


def gen_synth(n, games, rng):
    states, players, masks, endeds, scores, rots = [], [], [], [], [], []
    pool = [g["state"][i] for g in games[:8] for i in range(5, len(g["state"]), 3)]
    for k in range(260):
        base = pool[int(rng.integers(len(pool)))]
        if k % 4 == 3:
            s, p = endgame_variant(base, n, rng), int(rng.integers(n))
        else:
            s, p = perturb(base, n, rng)
        m, e, sc = rd.eval_state(s, n, p)
        states.append(s); players.append(p); masks.append(np.packbits(m, bitorder="little"))
        endeds.append(e); scores.append(sc); rots.append(rd.rotations(s, n))
    return dict(state=np.array(states), player=np.array(players, dtype=np.int8), mask=np.array(masks),
                ended=np.array(endeds), score=np.array(scores, dtype=np.int16), rot=np.array(rots))


def gen_sym(n, games, rng):
    states, pis, valids, counts, o_s, o_p, o_v = [], [], [], [], [], [], []
    pool = [g["state"][i] for g in games for i in range(3, len(g["state"]), 11)]
    for k in range(24):
        s = pool[int(rng.integers(len(pool)))]
        pi = rng.random(406).astype(np.float32)
        va = rng.random(406) < 0.3
        syms = rd.symmetries(s, n, pi, va)
        states.append(s); pis.append(pi); valids.append(va); counts.append(len(syms))
        for (a, b, c) in syms:
            o_s.append(a); o_p.append(b); o_v.append(c)
    return dict(state=np.array(states), pi=np.array(pis), valids=np.array(valids), count=np.array(counts, dtype=np.int32),
                out_state=np.array(o_s), out_pi=np.array(o_p), out_valids=np.array(o_v))


def main():
    os.makedirs(GOLD, exist_ok=True)
    for n in (2, 3, 4):
        rng = np.random.default_rng(77 + n)
        packed, games = gen_traj(n)
        np.savez_compressed(os.path.join(GOLD, f"traj_n{n}.npz"), **packed)
        np.savez_compressed(os.path.join(GOLD, f"synth_n{n}.npz"), **gen_synth(n, games, rng))
        np.savez_compressed(os.path.join(GOLD, f"sym_n{n}.npz"), **gen_sym(n, games, rng))
        print(n, "games", len(games), "plies", len(packed["action"]))


if __name__ == "__main__":
    main()
