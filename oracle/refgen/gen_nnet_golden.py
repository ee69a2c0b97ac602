#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - freezes outputs of the reference's own SplendorNNet (SplendorNNet.py, torch CPU float32, eval
mode) into tests/golden/nnet_n{2,3,4}.npz. Run in the build container: python oracle/refgen/gen_nnet_golden.py

Weights: the repo's `nnet.random_state_dict(n, seed)` (reference key names / shapes, reproducible from the seed) loaded
into the reference module with load_state_dict(strict=True) - so the fixture needs no weight file. Inputs: positions of
the golden trajectories with their legal masks. Outputs recorded exactly as GenericNNetWrapper.predict returns them
(:160-168): exp(log_softmax) and tanh(v).
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
SEED = 7


class FakeGame:
    def __init__(self, n):
        self.num_players = n

    def getBoardSize(self):
        return (32 + 10 * self.num_players + self.num_players ** 2, 7)

    def getActionSize(self):
        return 406

    def getMaxScoreDiff(self):
        return 15


def main():
    build_patched_ref.import_ref()
    from splendor.SplendorNNet import SplendorNNet
    import azg_b200
    from azg_b200 import nnet as mynet
    torch.set_num_threads(1)
    for n in (2, 3, 4):
        net = SplendorNNet(FakeGame(n), {"nn_version": 1, "dropout": 0.3}, use_token_exchange=True)
        sd = mynet.random_state_dict(n, SEED)
        net.load_state_dict(sd, strict=True)
        net.eval()
        g = np.load(os.path.join(GOLD, f"traj_n{n}.npz"))
        idx = np.linspace(0, len(g["state"]) - 1, 96).astype(int)
        states = g["state"][idx]
        valids = np.unpackbits(g["mask"][idx], axis=1, bitorder="little")[:, :406].astype(np.bool_)
        # the stored mask belongs to the mover before the move; any legal-looking mask exercises the masked softmax
        with torch.no_grad():
            pi, v, _ = net(torch.from_numpy(states.astype(np.float32)), torch.from_numpy(valids))
        np.savez_compressed(os.path.join(GOLD, f"nnet_n{n}.npz"), state=states, valids=valids, pi=torch.exp(pi).numpy(), v=v.numpy(),
                            seed=np.int64(SEED))
        print(n, "pi max", float(torch.exp(pi).max()), "v range", float(v.min()), float(v.max()))


if __name__ == "__main__":
    main()
