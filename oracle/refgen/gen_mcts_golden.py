#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - freezes outputs of the reference's own MCTS.py (unmodified, run on the patched reference
Game) into tests/golden/mcts_*.npz. Run in the build container: python oracle/refgen/gen_mcts_golden.py

The network is oracle/fakenn.py (fixed outputs, exact dyadic numbers); Dirichlet noise and the playout-cap coin
are injected through a stand-in for `MCTS.rng` (the reference uses an unseeded np.random.default_rng()).
Each scenario is a sequence of getActionProb calls on successive root states WITHOUT resetting the tree
(cross-move tree reuse, transpositions), recording per call: root state, full-search flag, Dirichlet vector,
returned probs / q, and the root node's raw Nsa / Qsa / Ns / Qs plus the size of the node dictionary.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.realpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import build_patched_ref  # noqa: E402
from oracle import fakenn  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


class dotdict(dict):
    def __getattr__(self, name):
        return self[name]


class FakeRng:
    def __init__(self, coin, seed):
        self.coin, self.seed, self.last_dir = coin, seed, None

    def random(self):
        return self.coin

    def dirichlet(self, alphas):
        self.seed += 1
        self.last_dir = fakenn.dirichlet(self.seed, len(alphas))
        return self.last_dir


SCENARIOS = {
    # name: (n, sims, cpuct, fpu, forced, noise, prob_full, ratio, force_full, moves, start)
    "a_n2_plain":   dict(n=2, sims=300, cpuct=1.0, fpu=0.0, forced=False, noise=False, prob_full=1.0, ratio=5, force=True, moves=6, start="init"),
    "b_n2_forced_noise": dict(n=2, sims=400, cpuct=2.5, fpu=0.3, forced=True, noise=True, prob_full=1.0, ratio=5, force=True, moves=5, start="init"),
    "c_n3_fpu":     dict(n=3, sims=250, cpuct=1.25, fpu=0.1, forced=False, noise=False, prob_full=1.0, ratio=5, force=True, moves=5, start="init"),
    "d_n4_forced":  dict(n=4, sims=200, cpuct=1.0, fpu=0.0, forced=True, noise=True, prob_full=1.0, ratio=5, force=True, moves=4, start="init"),
    "e_n2_cap":     dict(n=2, sims=400, cpuct=1.0, fpu=0.2, forced=True, noise=True, prob_full=0.0, ratio=4, force=False, moves=4, start="init"),
    "f_n2_late":    dict(n=2, sims=300, cpuct=1.5, fpu=0.0, forced=False, noise=False, prob_full=1.0, ratio=5, force=True, moves=5, start="late"),
    "g_n3_late":    dict(n=3, sims=200, cpuct=1.0, fpu=0.25, forced=True, noise=False, prob_full=1.0, ratio=5, force=True, moves=4, start="late"),
    # root softmax with args.temperature[0] != 1 before the noise (MCTS.py:141-143,150-153,244-250; main.py's default 1.25)
    "h_n2_temp":    dict(n=2, sims=400, cpuct=1.0, fpu=0.0, forced=True, noise=True, prob_full=1.0, ratio=5, force=True, moves=5, start="init", temp0=1.25),
    "i_n3_temp":    dict(n=3, sims=200, cpuct=1.25, fpu=0.1, forced=False, noise=True, prob_full=1.0, ratio=5, force=True, moves=4, start="late", temp0=0.8),
}


def late_state(game, n):
    """a canonical state deep in a game: replay of the golden trajectory 0 up to ~85 % of its length"""
    g = np.load(os.path.join(GOLD, f"traj_n{n}.npz"))
    off = g["offsets"]
    i = off[0] + int(0.85 * (off[1] - off[0]))
    st, nxt = g["state"][i].copy(), (int(g["player"][i]) + 1) % n
    return game.getCanonicalForm(st, nxt).copy()


def run(name, sc):
    from MCTS import MCTS
    from splendor.SplendorGame import SplendorGame
    n = sc["n"]
    game = SplendorGame(n)
    args = dotdict(numMCTSSims=sc["sims"], prob_fullMCTS=sc["prob_full"], ratio_fullMCTS=sc["ratio"], forced_playouts=sc["forced"],
                   cpuct=sc["cpuct"], fpu=sc["fpu"], no_mem_optim=True, temperature=[sc.get("temp0", 1.0), 1.0], dirichletAlpha=0.3)
    nnet = fakenn.FakeNNet(n)
    mcts = MCTS(game, nnet, args, dirichlet_noise=sc["noise"])
    mcts.rng = FakeRng(0.5, seed=hash(name) & 0xFFFF if False else sum(map(ord, name)))
    if sc["start"] == "init":
        gg = np.load(os.path.join(GOLD, f"traj_n{n}.npz"))
        board = gg["init_state"][1].copy()          # a reference-dealt start position (player 0 to move: canonical)
    else:
        board = late_state(game, n)
    rec = {k: [] for k in ["root", "full", "dir", "dir_len", "probs", "q", "nsa", "qsa", "ns", "qs", "nodes", "action", "nn_calls", "ps"]}
    for mv in range(sc["moves"]):
        if game.getGameEnded(board, 0).any():
            break
        mcts.rng.last_dir = None
        probs, q, full = mcts.getActionProb(board, temp=1, force_full_search=sc["force"])
        s = game.stringRepresentation(board)
        Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs = mcts.nodes_data[s]
        d = np.zeros(406); dl = 0
        if mcts.rng.last_dir is not None:
            dl = len(mcts.rng.last_dir); d[:dl] = mcts.rng.last_dir
        rec["root"].append(board.copy()); rec["full"].append(bool(full)); rec["dir"].append(d); rec["dir_len"].append(dl)
        rec["probs"].append(np.array(probs, dtype=np.float64)); rec["q"].append(np.array(q, dtype=np.float64))
        rec["nsa"].append(np.array(Nsa, dtype=np.int64)); rec["qsa"].append(np.array(Qsa, dtype=np.float64))
        rec["ns"].append(int(Ns)); rec["qs"].append(np.float32(Qs)); rec["nodes"].append(len(mcts.nodes_data))
        rec["nn_calls"].append(nnet.calls)
        rec["ps"].append(np.array(Ps, dtype=np.float32))        # the root's stored priors (after softmax + noise): compared bit for bit
        a = int(np.argmax(np.array(Nsa)))
        rec["action"].append(a)
        # advance the real game without a reveal (deterministic=True keeps the fixture free of the reference's RNG)
        nb, nxt = game.getNextState(board, 0, a, deterministic=True)
        board = game.getCanonicalForm(nb.copy(), nxt).copy()
    out = {k: np.array(v) for k, v in rec.items()}
    if "temp0" not in sc:
        del out["ps"]                                           # the older fixtures stay byte-identical
    out["cfg"] = np.array([sc["n"], sc["sims"], int(sc["forced"]), int(sc["noise"]), sc["ratio"], int(sc["force"])], dtype=np.int64)
    out["cfgf"] = np.array([sc["cpuct"], sc["fpu"], sc["prob_full"]], dtype=np.float64)
    if "temp0" in sc:
        out["temp0"] = np.float64(sc["temp0"])
    np.savez_compressed(os.path.join(GOLD, f"mcts_{name}.npz"), **out)
    print(name, "moves", len(rec["ns"]), "Ns", rec["ns"], "nodes", rec["nodes"], "nn", rec["nn_calls"])


def main():
    build_patched_ref.import_ref()
    import warnings
    warnings.filterwarnings("ignore")
    only = set(sys.argv[1:])          # optional: names of the scenarios to (re)generate; default all
    for name, sc in SCENARIOS.items():
        if not only or name in only:
            run(name, sc)


if __name__ == "__main__":
    main()
