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


def _blob_tiles(n, blob):
    """the weight tiles of the transposed evaluator's blob (csrc/spl_nnet2.cu: make_plan / pack_tile) as float32 [128][16 ks] arrays"""
    k1 = ((32 + 10 * n + n * n + 15) // 16 * 16) // 16
    ks = [min(k1, 4)] + ([k1 - 4] if k1 > 4 else []) + [4, 4, 4, 2, 4, 4, 1] + [4] * 11 + [4, 3, 4, 4, 1, 4, 4, 4, 3] + [4, 4, 1] * 4
    off, tiles = 1696 * 4, []
    for k in ks:
        raw = blob[off: off + k * 4096].view(np.uint16)
        off += k * 4096
        rows, cols = np.meshgrid(np.arange(128), np.arange(16 * k), indexing="ij")
        idx = ((rows >> 3) * (2 * k) * 128 + (cols >> 3) * 128 + (rows & 7) * 16 + (cols & 7) * 2) // 2     # canonical K-major core matrices
        tiles.append(torch.from_numpy((raw[idx].astype(np.uint32) << 16).view(np.float32)))
    assert off == blob.size
    return tiles


@pytest.mark.parametrize("n", [2, 3, 4])
def test_blob_packer_folds_batchnorm_like_the_torch_path(n):
    """spl_nnet_pack is pure host code: its fp32 parameter region must agree with nnet.fold (BatchNorm folded in float64) and every
    weight tile must hold the bf16 weights where the kernel's products expect them (pool-dense consumers: [120 dense | 8 zero], then
    the 8 pooled inputs as their own K = 16 step; the two Linear layers of each head folded into one matrix)"""
    import azg_b200
    sd = nnet.random_state_dict(n, 5)
    blob = nnet.pack_blob(n, sd).numpy()
    assert blob.size == azg_b200._native.lib().spl_nnet_blob_bytes(n) and blob.size % 16 == 0
    prm = blob[: 1696 * 4].view(np.float32)
    W = nnet.fold(sd, "cpu", torch.float32)
    assert np.allclose(prm[0:128], W["l1_b"].numpy()) and np.allclose(prm[1664:1671], W["bn1_s"].numpy().ravel(), rtol=1e-6)
    assert np.allclose(prm[1672:1679], W["bn1_t"].numpy().ravel(), rtol=1e-6, atol=1e-7)
    assert np.allclose(prm[640:760], W["g4_b"].numpy(), rtol=1e-6, atol=1e-7)          # BatchNorm1d(1) folded into the bias
    assert np.allclose(prm[768:896], W["l5a_b"].numpy(), rtol=1e-6, atol=1e-7)
    d = {k: v.double() for k, v in sd.items() if torch.is_tensor(v) and v.dtype.is_floating_point}
    heads = {h: (d[f"output_layers_{h}.1.weight"] @ d[f"output_layers_{h}.0.weight"],
                 d[f"output_layers_{h}.1.weight"] @ d[f"output_layers_{h}.0.bias"] + d[f"output_layers_{h}.1.bias"]) for h in ("PI", "V")}
    assert np.allclose(prm[1152:1152 + 406], heads["PI"][1].float().numpy(), rtol=1e-6, atol=1e-7)
    assert np.allclose(prm[1152 + 406:1152 + 406 + n], heads["V"][1].float().numpy(), rtol=1e-6, atol=1e-7)

    def bf(x):
        return x.float().to(torch.bfloat16).float()

    T = _blob_tiles(n, blob)
    R = 32 + 10 * n + n * n
    t = 0
    w1 = torch.zeros(128, (R + 63) // 64 * 64)
    w1[:, :R] = sd["dense2d_1.0.weight"]
    for k0 in range(0, R, 64):                                                          # L1, zero padded to 16 ks columns
        assert torch.equal(T[t], bf(w1[:, k0:k0 + T[t].shape[1]])); t += 1
    for k0 in (0, 64):                                                                  # L2
        assert torch.equal(T[t], bf(sd["dense2d_1.3.weight"][:, k0:k0 + 64])); t += 1
    g1 = torch.zeros(128, 96); g1[:120] = sd["partialgpool_1.dense_part.0.weight"]    # G1: 120 outputs, inputs = features 32..127
    assert torch.equal(T[t], bf(g1[:, :64])) and torch.equal(T[t + 1], bf(g1[:, 64:])); t += 2

    def after_pool(w, t):                                                               # [dense 0..119 | 8 zero], then [8 pooled | 8 zero]
        dense = torch.zeros(w.shape[0], 128); dense[:, :120] = w[:, 8:]
        side = torch.zeros(w.shape[0], 16); side[:, :8] = w[:, :8]
        assert torch.equal(T[t][:w.shape[0]], bf(dense[:, :64])) and torch.equal(T[t + 1][:w.shape[0]], bf(dense[:, 64:]))
        assert torch.equal(T[t + 2][:w.shape[0]], bf(side))
        assert not T[t][w.shape[0]:].any() and not T[t + 2][w.shape[0]:].any()
        return t + 3

    t = after_pool(sd["dense2d_3.0.weight"], t)                                       # L3
    for k0 in range(0, 704, 64):                                                        # L4
        assert torch.equal(T[t], bf(sd["dense1d_4.0.weight"][:, k0:k0 + 64])); t += 1
    g4 = torch.zeros(128, 112); g4[:120] = W["g4_w"].t()                              # G4 (BatchNorm1d(1) folded into the matrix)
    assert torch.allclose(T[t], bf(g4[:, :64]), rtol=8e-3, atol=1e-6) and torch.allclose(T[t + 1], bf(g4[:, 64:]), rtol=8e-3, atol=1e-6); t += 2
    t += 3 + 2 + 2                                                                      # L5a, L5b, G5: same packers
    full = torch.cat([heads["PI"][0], heads["V"][0], torch.zeros(512 - 406 - n, 128, dtype=torch.float64)], 0)
    for m in (0, 2, 1, 3):                                                              # head: 4 M tiles of the folded [406 + n][128] matrix
        t = after_pool(full[128 * m:128 * m + 128], t)
    assert t == len(T)
