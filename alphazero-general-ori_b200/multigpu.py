"""What crosses GPUs. Games are independent: rank r owns global game ids [r*L, (r+1)*L) and there is NO collective
on the step or search path. Per training iteration the ranks exchange
  * the finished-game examples          -> examples.gather_examples (all-gather of counts, then one padded payload per field)
  * the accepted network's weights      -> broadcast_weights (Coach.py:144-165: after the arena gate every self-play worker needs
                                           the new `best.pt`; 1.25 MB of float32)
  * arena results                       -> all_reduce_counts (Arena.playGames' oneWon / twoWon / draws, summed over the ranks' games)
torch.distributed does the plumbing (NCCL for device tensors, gloo for host tensors in the CPU tests).
"""
import torch


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def shard(total, rank=None, world=None):
    """games [lo, hi) of `total` for this rank (the first `total % world` ranks take one more)"""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    per, rem = divmod(int(total), world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def broadcast_weights(state_dict, src=0, device=None, group=None):
    """every rank ends up with rank `src`'s state_dict (same keys and shapes everywhere, e.g. a SplendorNNet checkpoint): one flat
    buffer, one broadcast. Returns the dict of tensors (on `device` if given, else where they were)."""
    dist = _dist()
    keys = sorted(state_dict.keys())
    if dist is None or dist.get_world_size(group) == 1:
        return {k: state_dict[k] for k in keys}
    dev = device if device is not None else state_dict[keys[0]].device
    flat = torch.cat([state_dict[k].detach().to(dev, torch.float64).reshape(-1) for k in keys])     # float64 carries int64 counters below 2^53 exactly
    dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    for k in keys:
        t = state_dict[k]
        out[k] = flat[off:off + t.numel()].reshape(t.shape).to(t.dtype)
        off += t.numel()
    return out


def all_reduce_counts(values, device=None, group=None):
    """sum of a few integers over the ranks (arena wins / draws, games played, simulations) -> list of ints"""
    dist = _dist()
    vals = [int(v) for v in values]
    if dist is None or dist.get_world_size(group) == 1:
        return vals
    t = torch.tensor(vals, dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [int(x) for x in t.cpu()]


def play_games_sharded(make_arena, total_games, device=None):
    """Arena.playGames(total_games) over all ranks: every rank plays its shard of the games (global game ids keep the 1-2-2-1 seat
    order and the Philox keys independent of the GPU count) and the results are summed. make_arena(game_base) -> BatchedArena."""
    lo, hi = shard(total_games)
    lo4 = lo - lo % 4                      # shards start on a multiple of four so that game i keeps seat order i % 4 (Arena.py:199)
    pit = make_arena(lo4)
    one, two, draws, d = pit.play_games(hi - lo4)
    skip = lo - lo4
    if skip:                               # games [lo4, lo) belong to the previous rank: drop them from this rank's tally
        r0, ovt = d["result_seat0"][:skip], d["one_vs_two"][:skip]
        o = int((((r0 == 1.0) & ovt) | ((r0 == -1.0) & ~ovt)).sum()); t = int((((r0 == -1.0) & ovt) | ((r0 == 1.0) & ~ovt)).sum())
        one, two, draws = one - o, two - t, draws - (skip - o - t)
    return tuple(all_reduce_counts([one, two, draws], device=device)) + (d,)
