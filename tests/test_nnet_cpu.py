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


@pytest.mark.parametrize("n", [2, 3, 4])
def test_blob_packer_folds_batchnorm_like_the_torch_path(n):
    """spl_nnet_pack is pure host code: its fp32 parameter region must agree with nnet.fold (BatchNorm folded in float64)"""
    import azg_b200
    sd = nnet.random_state_dict(n, 5)
    blob = nnet.pack_blob(n, sd).numpy()
    assert blob.size == azg_b200._native.lib().spl_nnet_blob_bytes(n) and blob.size % 16 == 0
    prm = blob[: 1896 * 4].view(np.float32)
    W = nnet.fold(sd, "cpu", torch.float32)
    assert np.allclose(prm[0:128], W["l1_b"].numpy()) and np.allclose(prm[128:135], W["bn1_s"].numpy().ravel(), rtol=1e-6)
    assert np.allclose(prm[136:143], W["bn1_t"].numpy().ravel(), rtol=1e-6, atol=1e-7)
    assert np.allclose(prm[672:792], W["g4_b"].numpy(), rtol=1e-6, atol=1e-7)          # BatchNorm1d(1) folded into the bias
    assert np.allclose(prm[800:928], W["l5a_b"].numpy(), rtol=1e-6, atol=1e-7)
    assert np.allclose(prm[1312:1312 + 406], W["PI1_b"].numpy()) and np.allclose(prm[1888:1888 + n], W["V1_b"].numpy())
    # first weight block: dense2d_1.0.weight [128][R] as bf16 in the canonical UMMA K-major layout (csrc/spl_umma.cuh):
    # element (row, k) at byte (row // 8) * (K1 // 8) * 128 + (k // 8) * 128 + (row % 8) * 16 + (k % 8) * 2, zero padded to K1
    R = 32 + 10 * n + n * n
    K1 = (R + 15) // 16 * 16
    blk = blob[1896 * 4: 1896 * 4 + 128 * K1 * 2].view(np.uint16)
    want = np.zeros((128, K1), dtype=np.uint16)
    want[:, :R] = sd["dense2d_1.0.weight"].to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    rows, ks = np.meshgrid(np.arange(128), np.arange(K1), indexing="ij")
    off = ((rows // 8) * (K1 // 8) * 128 + (ks // 8) * 128 + (rows % 8) * 16 + (ks % 8) * 2) // 2
    assert np.array_equal(blk[off], want)
