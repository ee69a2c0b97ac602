"""Host logic of the example path: finalisation arithmetic vs Coach.py:89-98 restated with np.roll, the Coach tuple format,
and the multi-GPU exchange on two gloo ranks."""
import os
import pickle
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import azg_b200
from azg_b200 import examples as ex


@pytest.mark.parametrize("n", [2, 3, 4])
def test_finalize_matches_coach_arithmetic(n):
    rng = np.random.default_rng(n)
    E = 200
    players = rng.integers(0, n, E)
    c_final = rng.integers(0, n, E)
    r_abs = rng.choice([1.0, -1.0, 0.01], size=(E, n)).astype(np.float32)          # getGameEnded in absolute seat order
    sc_abs = rng.integers(0, 22, size=(E, n)).astype(np.int32)
    # what the engine sees: the final canonical frame (index 0 = seat c_final)
    r_canon = np.stack([np.roll(r_abs[e], -c_final[e]) for e in range(E)])
    sc_canon = np.stack([np.roll(sc_abs[e], -c_final[e]) for e in range(E)])
    w, d = ex.finalize_examples(torch.from_numpy(players), torch.from_numpy(r_canon), torch.from_numpy(sc_canon), torch.from_numpy(c_final))
    for e in range(E):
        p = players[e]
        assert np.array_equal(w[e].numpy(), np.roll(r_abs[e], -p))                                        # Coach.py:94
        assert np.array_equal(d[e].numpy(), np.roll([f - sc_abs[e][p] for f in sc_abs[e]], -p))           # Coach.py:95


def _fake(n, count, seed):
    g = torch.Generator().manual_seed(seed)
    R = 32 + 10 * n + n * n
    return dict(board=torch.randint(-3, 9, (count, R, 7), generator=g).to(torch.int8), pi=torch.rand((count, 406), generator=g),
                winner=torch.rand((count, n), generator=g), scdiff=torch.randint(-15, 15, (count, n), generator=g).to(torch.int32),
                valids=(torch.rand((count, 406), generator=g) < 0.1).to(torch.uint8), surprise=torch.rand((count, n), generator=g))


def test_coach_tuple_format():
    e = _fake(2, 5, 0)
    lst = ex.to_coach_format(e, compress=True)
    assert len(lst) == 5
    b, p, w, d, v, s = pickle.loads(zlib.decompress(lst[3]))          # Coach.py:100 / GenericNNetWrapper.pick_examples :325-331
    assert b.shape == (56, 7) and b.dtype == np.int8 and p.shape == (406,) and v.dtype == np.bool_ and len(s) == 2
    assert np.array_equal(b, e["board"][3].numpy()) and np.array_equal(d, e["scdiff"][3].numpy())


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _fake(3, 3 + 4 * rank, 100 + rank)       # ragged: 3 and 7 examples
    allx = ex.gather_examples(mine)
    ok = True
    off = 0
    for r in range(world):
        want = _fake(3, 3 + 4 * r, 100 + r)
        for k in ex.FIELDS:
            ok &= bool(torch.equal(allx[k][off:off + want[k].shape[0]], want[k]))
        off += 3 + 4 * r
    ok &= allx["board"].shape[0] == off
    none = ex.gather_examples(ex.empty_examples(3, 71, "cpu") if rank == 0 else _fake(3, 2, 7))    # a rank with nothing to give
    ok &= none["pi"].shape[0] == 2
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_examples_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def _worker_mg(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from azg_b200 import multigpu as mg, nnet
    ok = mg.shard(10) == ((0, 5) if rank == 0 else (5, 10)) and mg.shard(7, 1, 3) == (3, 5)
    sd = nnet.random_state_dict(2, seed=40 + rank)                     # every rank starts from its own weights
    got = mg.broadcast_weights(sd, src=1)
    want = nnet.random_state_dict(2, seed=41)
    ok &= all(torch.equal(got[k], want[k]) and got[k].dtype == want[k].dtype for k in want)
    ok &= mg.all_reduce_counts([3 + rank, 10 * rank, 1]) == [7, 10, 2]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_weight_broadcast_and_count_reduction_two_ranks_gloo():
    """the accepted network's weights go to every rank, arena tallies are summed"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_mg, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_game_sharding_is_independent_of_world_size():
    """lane l of rank r is global game r*L + l: the Philox start of a game depends on its id only (CPU check on the rules core)"""
    from tests.hostsim import sim as hs
    a = hs.Sim(2); a.init_philox(99, 5 * 64 + 3)           # rank 5 of 64-lane shards, lane 3
    b = hs.Sim(2); b.init_philox(99, 2 * 128 + 67)         # rank 2 of 128-lane shards, lane 67
    assert np.array_equal(a.state, b.state)
