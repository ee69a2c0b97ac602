"""Distribution tests for the two places where the engine draws its own random numbers instead of the reference's generator
(Numba's MT19937 / np.random.default_rng cannot be reproduced on the device; parity runs replay or inject the draws):

  * deck reveals: `_get_deck_card` (SplendorLogicNumba.py:400-412) draws the colour in proportion to the cards left per colour,
    then uniformly among that colour's remaining cards. Chi-square of the Philox reveals against that two-stage law (and, when the
    reference is importable, a two-sample chi-square against the reference's own sampler), per tier, on a mid-game deck
  * root noise: `applyDirNoise` (MCTS.py:180-186) mixes 25 % of rng.dirichlet([alpha] * k) into the root priors. Mean and variance of
    the on-device gamma / Dirichlet sampler against the closed form (= numpy's), and fresh noise per episode and per move
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pyoracle as po

CHI2_999 = {4: 18.47, 9: 27.88, 14: 36.12, 19: 43.82, 22: 48.27, 24: 51.18, 29: 58.30, 34: 65.25, 39: 72.05}   # 0.999 quantiles


def _chi2_crit(df):
    k = min(CHI2_999, key=lambda d: abs(d - df))
    return CHI2_999[k] * (df / k) if k != df else CHI2_999[k]      # (crude scaling between tabulated df; generous side)


def _midgame(n=2, seed=9, plies=26):
    """a position some plies into a random game in which the player to move may reserve a visible card of every tier"""
    rng = np.random.default_rng(seed)
    b = po.Board(n); b.init_philox(77, seed)
    ply = 0
    while True:
        v = b.valid_moves(0)
        if ply >= plies and all(v[12 + 4 * t: 16 + 4 * t].any() for t in range(3)):
            return b
        nxt = b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, 77, seed, 0); b.swap_players(nxt)
        ply += 1


def _reserve_action(b, tier):
    v = b.valid_moves(0)
    return 12 + 4 * tier + int(np.flatnonzero(v[12 + 4 * tier: 16 + 4 * tier])[0])


def _expected(state, tier):
    """P(colour, idx) of _get_deck_card for the deck rows of `state`"""
    cnt = state[25 + 2 * tier, :5].astype(int)
    bits = state[26 + 2 * tier, :5].astype(np.uint8)
    p = {}
    for c in range(5):
        for i in range(8):
            if bits[c] & (128 >> i):
                p[c * 8 + i] = (cnt[c] / cnt.sum()) * (1.0 / cnt[c])
    return p


def _deck_diff(before, after):
    for t in range(3):
        rb, ra = before[26 + 2 * t, :5].astype(np.uint8), after[26 + 2 * t, :5].astype(np.uint8)
        for c in range(5):
            d = int(rb[c]) & ~int(ra[c])
            if d:
                return t, c * 8 + (8 - d.bit_length())
    return -1, -1


def _chi2(counts, probs, N):
    return sum((counts.get(k, 0) - N * p) ** 2 / (N * p) for k, p in probs.items())


@pytest.mark.parametrize("tier", [0, 1, 2])
def test_philox_reveals_follow_the_two_stage_law(tier):
    b = _midgame()
    N = 20000
    probs = _expected(b.state, tier)
    assert abs(sum(probs.values()) - 1) < 1e-12 and len(probs) >= 5
    uneven = len(set(round(p, 9) for p in probs.values())) > 1      # colour-then-card is NOT uniform over cards when colours differ in size
    act = _reserve_action(b, tier)
    counts = {}
    for g in range(N):
        c = b.copy()
        c.make_move(act, 0, -2, 1234, g, 0)     # reserve a visible card of the tier: its slot is refilled from the deck
        t, code = _deck_diff(b.state, c.state)
        assert t == tier
        counts[code] = counts.get(code, 0) + 1
    assert set(counts) <= set(probs)
    assert _chi2(counts, probs, N) < _chi2_crit(len(probs) - 1)
    if uneven:    # a sampler that drew uniformly over the remaining CARDS must fail this test
        uni = {k: 1.0 / len(probs) for k in probs}
        assert _chi2(counts, uni, N) > 3 * _chi2_crit(len(probs) - 1)


@pytest.mark.ref
@pytest.mark.parametrize("tier", [0, 2])
def test_philox_reveals_vs_the_reference_sampler(tier):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "refgen"))
    import build_patched_ref
    if build_patched_ref.find_ref() is None:
        pytest.skip("reference not available")
    import ref_driver
    R = ref_driver.ref()
    b = _midgame()
    N = 12000
    R["seed_numba"](5)
    rb = R["Board"](2)
    ref_counts, our_counts = {}, {}
    act = _reserve_action(b, tier)
    for g in range(N):
        rb.copy_state(b.state, True)
        rb._get_deck_card(tier)
        t, code = _deck_diff(b.state, rb.get_state())
        ref_counts[code] = ref_counts.get(code, 0) + 1
        c = b.copy()
        c.make_move(act, 0, -2, 99, g, 0)
        t, code = _deck_diff(b.state, c.state)
        our_counts[code] = our_counts.get(code, 0) + 1
    keys = sorted(set(ref_counts) | set(our_counts))
    stat = sum((ref_counts.get(k, 0) - our_counts.get(k, 0)) ** 2 / (ref_counts.get(k, 0) + our_counts.get(k, 0)) for k in keys)   # two-sample chi-square, equal sizes
    assert stat < _chi2_crit(len(keys) - 1), (stat, len(keys))


@pytest.mark.gpu
@pytest.mark.parametrize("tier", [0, 1, 2])
def test_device_reveals_follow_the_two_stage_law(tier):
    """the same law through the step kernel: 65,536 lanes hold the same position and reserve the same card; the game id in the
    Philox key is all that differs"""
    import azg_b200 as az
    b = _midgame()
    L = 65536
    env = az.SplendorEnv(2, L, seed=4321)
    env.set_states(torch.from_numpy(np.repeat(b.state[None], L, 0)))
    env.step(torch.full((L,), _reserve_action(b, tier), dtype=torch.int16, device=env.device), player=0, chance="philox", want_mask=False)
    after = env.states().cpu().numpy()
    before_bits = b.state[26 + 2 * tier, :5].astype(np.uint8)
    diff = before_bits[None, :] & ~after[:, 26 + 2 * tier, :5].astype(np.uint8)
    col = diff.argmax(1)
    d = diff[np.arange(L), col].astype(int)
    assert (np.count_nonzero(diff, axis=1) == 1).all()
    idx = 8 - np.floor(np.log2(d)).astype(int) - 1
    codes = col * 8 + idx
    counts = dict(zip(*[x.tolist() for x in np.unique(codes, return_counts=True)]))
    probs = _expected(b.state, tier)
    assert set(counts) <= set(probs)
    assert _chi2(counts, probs, L) < _chi2_crit(len(probs) - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [0.2, 0.3, 1.5])
def test_device_dirichlet_sampler_moments(alpha):
    """4096 trees hold the same root; the network row is the same for all (fixed network), so the stored priors differ only by the
    noise: Ps = 0.75 * P + 0.25 * Dir with Dir ~ Dirichlet(alpha, ..., alpha) over the k legal moves (MCTS.py:180-186).
    E[Dir_i] = 1/k, Var[Dir_i] = (1/k)(1 - 1/k)/(k alpha + 1) - the moments numpy's rng.dirichlet has."""
    import azg_b200 as az
    b = _midgame(plies=20)
    T = 4096
    k = int(b.valid_moves(0).sum())
    ar = az.MCTSArena(2, T, node_cap=16, dirichlet_alpha=alpha, seed=2024)
    roots = torch.from_numpy(np.repeat(b.state[None], T, 0)).to(ar.device)
    one = torch.ones(T, dtype=torch.int32, device=ar.device)
    ar.begin(roots, one, torch.full((T,), 2, dtype=torch.uint8, device=ar.device))      # SPL_MCTS_MOVE_NOISE
    ar.wave(lambda s, v: ar.fixed_net(s, v))
    ps = ar.root_stats()["ps"].double().cpu().numpy()
    p0, _ = po.fake_predict(b.state, b.valid_moves(0), 2)
    legal = np.flatnonzero(b.valid_moves(0))
    dirs = (ps[:, legal] - 0.75 * p0[legal][None, :].astype(np.float64)) / 0.25
    assert np.abs(dirs.sum(1) - 1).max() < 1e-5 and dirs.min() > -1e-6
    mean, var = 1.0 / k, (1.0 / k) * (1 - 1.0 / k) / (k * alpha + 1)
    se = np.sqrt(var / T)
    assert np.abs(dirs.mean(0) - mean).max() < 5 * se                       # every coordinate's mean (k of them, 5 sigma)
    assert abs(dirs.var(0).mean() / var - 1) < 0.08                         # average variance over the coordinates
    ref = np.random.default_rng(0).dirichlet([alpha] * k, size=T)           # numpy's sampler: the same two statistics
    assert abs(ref.var(0).mean() / var - 1) < 0.08
    # sparsity of small alpha: the share of (near) zero coordinates matches numpy's
    assert abs((dirs < 1e-3).mean() - (ref < 1e-3).mean()) < 0.03


@pytest.mark.gpu
def test_device_dirichlet_noise_is_fresh_per_episode_and_per_move():
    """the sampler is keyed (seed, game, episode, root ply): a lane that restarts draws other noise at the same ply of its next
    game (the reference draws rng.dirichlet per move), different lanes differ, the same key repeats"""
    import azg_b200 as az
    b = _midgame(plies=20)
    T = 64
    roots = torch.from_numpy(np.repeat(b.state[None], T, 0))

    def noise(episodes, seed=7):
        ar = az.MCTSArena(2, T, node_cap=16, dirichlet_alpha=0.3, seed=seed)
        ar.set_episodes(episodes.to(ar.device))
        one = torch.ones(T, dtype=torch.int32, device=ar.device)
        ar.begin(roots.to(ar.device), one, torch.full((T,), 2, dtype=torch.uint8, device=ar.device))
        ar.wave(lambda s, v: ar.fixed_net(s, v))
        return ar.root_stats()["ps"].cpu()
    e0 = torch.zeros(T, dtype=torch.int32)
    a, a2, b1 = noise(e0), noise(e0), noise(e0 + 1)
    assert torch.equal(a, a2)
    assert float((a - b1).abs().max(dim=1).values.min()) > 1e-4            # every lane drew something else in its next episode
    assert float((a[0] - a[1]).abs().max()) > 1e-4                          # lanes differ
    mixed = noise(torch.arange(T, dtype=torch.int32) % 2)
    assert torch.equal(mixed[0::2], a[0::2]) and torch.equal(mixed[1::2], b1[1::2])
