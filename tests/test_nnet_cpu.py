"""The folded inference pass (alphazero-general-ori_b200/nnet.py) vs outputs of the reference's own SplendorNNet
(tests/golden/nnet_n*.npz from oracle/refgen/gen_nnet_golden.py: torch CPU float32, eval mode, same weights by seed).
Pure torch arithmetic, so the math is checked here on the CPU; tests/test_gpu_mcts.py repeats it on the device.
Tolerance (float32 network, BatchNorm folded in float64): 2e-5 absolute on probabilities and values."""
import os

import numpy as np
import pytest
import torch

from azg_b200 import nnet


@pytest.mark.parametrize("n", [2, 3, 4])
def test_folded_forward_matches_reference_network(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"nnet_n{n}.npz"))
    W = nnet.fold(nnet.random_state_dict(n, int(g["seed"])), "cpu", torch.float32)
    pi, v = nnet.forward_folded(W, torch.from_numpy(g["state"]), torch.from_numpy(g["valids"]))
    assert np.abs(pi.numpy() - g["pi"]).max() < 2e-5
    assert np.abs(v.numpy() - g["v"]).max() < 2e-5
    assert (pi.numpy()[~g["valids"]] == 0).all()


def test_state_dict_shapes_are_the_checkpoint_format():
    sh = nnet.state_dict_shapes(2)
    assert sh["dense2d_1.0.weight"] == (128, 56) and sh["output_layers_PI.1.weight"] == (406, 128)
    # (the generator script loads random_state_dict into the reference module with strict=True: names and shapes are the reference's)
    assert set(nnet.random_state_dict(3, 1).keys()) == set(nnet.state_dict_shapes(3).keys())
