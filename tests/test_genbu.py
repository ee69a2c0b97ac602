"""genbu.pt - the one trained checkpoint the reference ships (2 players, 406 actions) - through every evaluator.

tests/golden/genbu_n2.npz (oracle/refgen/gen_genbu_golden.py): the checkpoint's state_dict and the outputs of its pickled
`full_model` (the reference's own SplendorNNet object, torch CPU float32) on 256 mid-game positions.
  * CPU: the folded float32 pass reproduces full_model to 2e-5; the checkpoint reader pulls the weights out of the reference's
    file without unpickling its module
  * GPU: the same for SplendorNNetB200 on the device; the fused bf16 tensor-core kernel against the float32 pass on 4096
    mid-game positions - maximum / mean absolute error, top-1 agreement and KL divergence, with the tolerances in the asserts;
    a search with the trained network, root softmax (temperature 1.25) and the on-device Dirichlet sampler
"""
import os

import numpy as np
import pytest
import torch


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "genbu_n2.npz"))


def test_genbu_folded_float32_matches_the_reference_module(golden_dir):
    from azg_b200 import nnet
    g = _golden(golden_dir)
    sd = nnet.state_dict_from_npz(os.path.join(golden_dir, "genbu_n2.npz"))
    assert len(sd) == 56 and sum(v.numel() for v in sd.values()) == 312584 + 7 * 0 and tuple(sd["output_layers_PI.1.weight"].shape) == (406, 128)
    W = nnet.fold(sd, "cpu", torch.float32)
    pi, v = nnet.forward_folded(W, torch.from_numpy(g["state"]), torch.from_numpy(g["valids"].astype(np.uint8)))
    assert float((pi - torch.from_numpy(g["pi"])).abs().max()) < 2e-5 and float((v - torch.from_numpy(g["v"])).abs().max()) < 2e-5


@pytest.mark.ref
def test_checkpoint_reader_takes_only_the_weights():
    """GenericNNetWrapper.save_checkpoint (:185-198) pickles the whole module next to the state_dict; the reader must not need
    (or execute) it"""
    from azg_b200 import nnet
    path = "/root/reference/genbu.pt"
    if not os.path.isfile(path):
        pytest.skip("reference checkpoint not present")
    ck = nnet.load_checkpoint_file(path)
    assert "state_dict" in ck and ck["numMCTSSims"] == 800 and not isinstance(ck["full_model"], torch.nn.Module)
    here = nnet.state_dict_from_npz(os.path.join(os.path.dirname(__file__), "golden", "genbu_n2.npz"))
    assert all(torch.equal(ck["state_dict"][k], here[k]) for k in here)


@pytest.mark.gpu
def test_genbu_on_the_device_float32(golden_dir):
    import azg_b200 as az
    g = _golden(golden_dir)
    torch.backends.cuda.matmul.allow_tf32 = False
    net = az.SplendorNNetB200(2, state_dict=az.nnet.state_dict_from_npz(os.path.join(golden_dir, "genbu_n2.npz")), dtype=torch.float32)
    pi, v = net(torch.from_numpy(g["state"]).to(net.device), torch.from_numpy(g["valids"].astype(np.uint8)).to(net.device))
    assert float((pi.cpu() - torch.from_numpy(g["pi"])).abs().max()) < 2e-5 and float((v.cpu() - torch.from_numpy(g["v"])).abs().max()) < 2e-5


@pytest.mark.gpu
def test_genbu_fused_bf16_kernel_agreement(golden_dir):
    """the one-launch bf16 evaluator with trained weights vs the float32 pass: 4096 positions from device self-play openings.
    Tolerances (bf16 products, fp32 accumulation): max |dp| < 0.08, mean |dp| < 5e-4 per action, top-1 agreement >= 97 %,
    mean KL(p32 || p16) < 2e-3, max |dv| < 0.06."""
    import azg_b200 as az
    sd = az.nnet.state_dict_from_npz(os.path.join(golden_dir, "genbu_n2.npz"))
    env = az.SplendorEnv(2, 4096, seed=31)
    env.reset()
    env.rollout(6, rotate=True)
    states, valids = [], []
    for k in range(8):        # 8 x 512 lanes, each batch a few plies further into its games
        env.rollout(9, rotate=True)
        env.step(None, player=0, store_state=False, want_ended=False, want_status=False)
        sl = slice(512 * k, 512 * (k + 1))
        states.append(env.states()[sl].clone()); valids.append(env.valids()[sl].clone())
    states, valids = torch.cat(states).contiguous(), torch.cat(valids).contiguous()
    torch.backends.cuda.matmul.allow_tf32 = False
    p32, v32 = az.SplendorNNetB200(2, state_dict=sd, dtype=torch.float32)(states, valids)
    fused = az.FusedSplendorNNet(2, state_dict=sd)
    p16, v16 = fused(states, valids)
    torch.cuda.synchronize()
    p32, v32, p16, v16 = p32.double(), v32.double(), p16.double().clone(), v16.double().clone()
    max_dp, mean_dp, max_dv = float((p32 - p16).abs().max()), float((p32 - p16).abs().mean()), float((v32 - v16).abs().max())
    top1 = float((p32.argmax(1) == p16.argmax(1)).double().mean())
    m = valids.bool() & (p32 > 0)
    kl = float((torch.where(m, p32 * (torch.log(p32.clamp_min(1e-300)) - torch.log(p16.clamp_min(1e-12))), torch.zeros_like(p32))).sum(1).mean())
    print(f"genbu fused vs float32 on {states.shape[0]} positions: max|dp| {max_dp:.4f} mean|dp| {mean_dp:.2e} top-1 {top1:.4f} KL {kl:.2e} max|dv| {max_dv:.4f}")
    assert (p16[~valids.bool()] == 0).all() and np.allclose(p16.sum(1).cpu().numpy(), 1.0, atol=1e-4)
    assert max_dp < 0.08 and mean_dp < 5e-4 and top1 >= 0.97 and kl < 2e-3 and max_dv < 0.06


@pytest.mark.gpu
def test_search_with_genbu_and_config2_settings(golden_dir):
    """SURVEY config 2 in small: genbu weights, cpuct 1.0, dirichletAlpha 0.2, temperature[0] 1.25, forced playouts, captured
    waves; budgets are spent exactly and the root priors are a distribution"""
    import azg_b200 as az
    sd = az.nnet.state_dict_from_npz(os.path.join(golden_dir, "genbu_n2.npz"))
    n, T, sims = 2, 256, 48
    net = az.FusedSplendorNNet(n, state_dict=sd)
    eng = az.SelfPlayEngine(n, T, net, sims, seed=3, cpuct=1.0, fpu=0.0, prob_full=0.25, ratio_full=5, forced_playouts=True,
                            dirichlet_noise=True, dirichlet_alpha=0.2, temperature0=1.25, node_cap=4096)
    eng.env.rollout(20, rotate=True)
    for _ in range(6):
        probs, q, is_full, ended = eng.play_move()
        st = eng.arena.root_stats()
        assert int(st["status"].max()) == 0
        assert torch.equal(st["sims_done"], eng.sims)
        ps = st["ps"].sum(1)
        assert bool(((ps - 1).abs() < 1e-4).all()) and abs(float(probs.sum(1).mean()) - 1.0) < 1e-9
    assert 0.05 < float(is_full.float().mean()) < 0.5
