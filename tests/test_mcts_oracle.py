"""Pins the search oracle (oracle/mcts_oracle.c) to outputs of the reference's own MCTS.py.

Fixtures: tests/golden/mcts_*.npz from oracle/refgen/gen_mcts_golden.py (reference MCTS.py, unmodified, on the patched
reference Game, network = oracle/fakenn.py, injected Dirichlet vectors). Bar: visit counts exact, Qsa within 1e-12,
probs / q within 1e-12, node-dictionary size and network-call count exact.
"""
import glob
import os

import numpy as np
import pytest

from oracle import fakenn
from oracle import pyoracle as po

SCEN = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mcts_*.npz")))


def test_fakenn_c_equals_python():
    rng = np.random.default_rng(1)
    for n in (2, 3, 4):
        b = po.Board(n); b.init_philox(3, n)
        for _ in range(30):
            v = b.valid_moves(0)
            ps, vv = fakenn.predict(b.state, v, n)
            pc, vc = po.fake_predict(b.state, v, n)
            assert np.array_equal(ps, pc) and np.array_equal(vv, vc)
            assert ps.sum(dtype=np.float32) == 1.0 and ps[::-1].sum(dtype=np.float32) == 1.0    # exact in any order
            b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -1); b.swap_players(1)


@pytest.mark.parametrize("name", SCEN)
def test_oracle_reproduces_reference_mcts(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    n, sims, forced, noise, ratio, force = [int(x) for x in g["cfg"]]
    cpuct, fpu, prob_full = [float(x) for x in g["cfgf"]]
    temp0 = float(g["temp0"]) if "temp0" in g else 1.0          # args.temperature[0]: root softmax before the noise
    m = po.MCTSOracle(n, sims, cpuct=cpuct, fpu=fpu, forced_playouts=bool(forced), dirichlet_noise=bool(noise), ratio_full=ratio,
                      temperature0=temp0)
    for i in range(len(g["ns"])):
        d = g["dir"][i] if g["dir_len"][i] > 0 else (np.zeros(406) if noise else None)
        out = m.get_action_prob(g["root"][i], temp=1.0, full_search=bool(g["full"][i]), dir_values=d)
        assert np.array_equal(out["nsa"], g["nsa"][i]), (name, i, "visit counts")
        assert out["ns"] == g["ns"][i] and m.num_nodes == g["nodes"][i] and m.nn_calls == g["nn_calls"][i]
        assert np.allclose(out["qsa"], g["qsa"][i], rtol=0, atol=1e-12), (name, i, np.abs(out["qsa"] - g["qsa"][i]).max())
        assert out["qs"] == g["qs"][i]
        assert np.allclose(out["probs"], g["probs"][i], rtol=0, atol=1e-12)
        assert np.allclose(out["q"], g["q"][i], rtol=0, atol=1e-12)
