#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - freezes the reference's shipped checkpoint genbu.pt (the only trained network in the reference,
SURVEY F4: 2 players, 406 actions) into tests/golden/genbu_n2.npz: its `state_dict` (weights only: the pickled `full_model`
object needs the reference package to unpickle and is not stored) and the outputs of that pickled full_model itself - the
reference's own SplendorNNet instance, torch CPU float32, eval mode, exactly as GenericNNetWrapper.predict returns them
(:160-168) - on 256 mid-game positions with their legal masks. Run in the build container:

    python oracle/refgen/gen_genbu_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.realpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import build_patched_ref  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def main():
    build_patched_ref.import_ref(callers=True)
    from oracle import pyoracle as po
    torch.set_num_threads(1)
    ck = torch.load(os.path.join(build_patched_ref.REF, "genbu.pt"), map_location="cpu", weights_only=False)
    sd = {k: v.detach().cpu().numpy() for k, v in ck["state_dict"].items()}
    model = ck["full_model"].eval()
    rng = np.random.default_rng(5)
    states, valids = [], []
    g = 0
    while len(states) < 256:     # positions spread over whole random games (oracle rules, Philox chance)
        b = po.Board(2); b.init_philox(4711, g); g += 1
        for ply in range(200):
            if b.check_end_game().any():
                break
            v = b.valid_moves(0)
            if ply >= 4 and rng.random() < 0.12 and len(states) < 256:
                states.append(b.state.copy()); valids.append(v.copy())
            nxt = b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 4711, g, 0); b.swap_players(nxt)
    states, valids = np.array(states), np.array(valids)
    with torch.no_grad():
        pi, v, _ = model(torch.from_numpy(states.astype(np.float32)), torch.from_numpy(valids))
    args = {k: ck[k] for k in ("numMCTSSims", "cpuct", "fpu", "dirichletAlpha", "forced_playouts") if k in ck}
    np.savez_compressed(os.path.join(GOLD, "genbu_n2.npz"), state=states, valids=valids, pi=torch.exp(pi).numpy(), v=v.numpy(),
                        sd_keys=np.array(list(sd.keys())), **{"sd/" + k: a for k, a in sd.items()})
    print("genbu: ", len(sd), "tensors,", sum(a.size for a in sd.values()), "parameters; training args in the checkpoint:", args)
    print("pi max", float(torch.exp(pi).max()), "v range", float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
