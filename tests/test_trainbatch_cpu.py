"""The training mini-batch assembly (azg_b200.trainbatch) against a NumPy restatement of the reference's own statements
(GenericNNetWrapper.py:68-80 batch build, :333-341 surprise weights). The module is tensor plumbing and runs wherever the
example tensors live; here on CPU tensors."""
import numpy as np
import torch

import azg_b200
from azg_b200 import trainbatch as tb


def _examples(N, n, seed):
    rng = np.random.default_rng(seed)
    R = 32 + 10 * n + n * n
    return dict(board=torch.from_numpy(rng.integers(-3, 9, size=(N, R, 7)).astype(np.int8)),
                pi=torch.from_numpy(rng.random((N, 406)).astype(np.float32)),
                winner=torch.from_numpy(rng.choice([-1.0, 1.0, 0.01], size=(N, n)).astype(np.float32)),
                scdiff=torch.from_numpy(rng.integers(-22, 23, size=(N, n)).astype(np.int32)),
                valids=torch.from_numpy((rng.random((N, 406)) < 0.1).astype(np.uint8)),
                surprise=torch.from_numpy(rng.standard_normal((N, n)).astype(np.float32)))


def _ref_targets(scdiffs, n, max_diff):      # GenericNNetWrapper.py:76-80, statement by statement
    target = np.zeros((len(scdiffs), 2 * max_diff + 1, n), dtype=np.float32)
    for i in range(len(scdiffs)):
        score_diff = (scdiffs[i] + max_diff).clip(0, 2 * max_diff)
        for player in range(n):
            target[i, score_diff[player], player] = 1
    return target


def test_batch_equals_reference_statements():
    for n in (2, 3, 4):
        ex = _examples(500, n, n)
        b = tb.TrainBatcher(ex, 64, max_diff=15, seed=1)
        out = b.batch()
        ids = out["ids"].numpy()
        assert len(set(ids.tolist())) == 64 and b.batches_per_epoch == 500 // 64          # replace=False
        assert np.array_equal(out["boards"].numpy(), ex["board"].numpy()[ids].astype(np.float32)) and out["boards"].dtype == torch.float32
        assert np.array_equal(out["valid_actions"].numpy(), ex["valids"].numpy()[ids].astype(np.bool_))
        assert np.array_equal(out["target_pis"].numpy(), ex["pi"].numpy()[ids]) and np.array_equal(out["target_vs"].numpy(), ex["winner"].numpy()[ids])
        assert np.array_equal(out["target_scdiffs"].numpy(), _ref_targets(ex["scdiff"].numpy()[ids].astype(np.int64), n, 15))
        assert out["target_scdiffs"].sum().item() == 64 * n


def test_surprise_weights_formula_and_sampling():
    ex = _examples(400, 2, 7)
    s = ex["surprise"][:, 0].abs()
    w = tb.surprise_weights(s).numpy()
    sn = s.numpy().astype(np.float64)
    ref = sn / sn.sum() + 1.0 / len(sn)          # :338-339
    ref = ref / ref.sum()
    assert np.allclose(w, ref, rtol=1e-12, atol=0) and abs(w.sum() - 1) < 1e-12
    # sampling follows the weights: inclusion frequency of the heaviest decile vs the lightest, over many batches
    b = tb.TrainBatcher(ex, 40, surprise_weight=True, seed=3)
    hits = np.zeros(400)
    for _ in range(600):
        ids = b.sample_ids().numpy()
        assert len(set(ids.tolist())) == 40
        hits[ids] += 1
    order = np.argsort(w)
    assert hits[order[-40:]].mean() > 1.3 * hits[order[:40]].mean()
    # uniform when the switch is off (the reference's default, main.py)
    u = tb.TrainBatcher(ex, 40, seed=4)
    hu = np.zeros(400)
    for _ in range(600):
        hu[u.sample_ids().numpy()] += 1
    assert abs(hu.mean() - 60) < 1e-9 and hu.std() < 12


def test_feeds_on_engine_examples_format():
    """the dict drain_examples / gather_examples produce (examples.FIELDS) is what the batcher takes"""
    assert set(azg_b200.examples.FIELDS) == {"board", "pi", "winner", "scdiff", "valids", "surprise"}
    ex = _examples(70, 3, 9)
    out = tb.TrainBatcher(ex, 70, seed=0).batch()
    assert sorted(out["ids"].tolist()) == list(range(70))
