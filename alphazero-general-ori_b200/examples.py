"""Training examples of batched self-play: assembly on the device, exchange between GPUs, hand-over to the reference.

What it replaces / feeds (reference):
  * Coach.executeEpisode's bookkeeping (Coach.py:72-100): one example per full-search move
    `[canonicalBoard, curPlayer, pi, valids, surprise]`, completed at the end of the game with
    `winner = np.roll(r, -curPlayer)` and `score difference = np.roll([f - final[curPlayer] for f in final], -curPlayer)`;
    the symmetric variants of Coach.py:77-80 (getSymmetries) are produced by the device kernel (SplendorEnv.symmetries)
  * the example tuple `(board, pi, winner, scdiff, valids, surprise)` and its zlib+pickle form (Coach.py:91-100), which
    GenericNNetWrapper.train / pick_examples consume unchanged (GenericNNetWrapper.py:43-139, 325-331)
  * multi-GPU: games shard over ranks with no collective on the step / search path; once per training
    iteration every rank contributes its examples to an all-gather (counts first, then the padded payload).

Everything here is tensor plumbing (torch); the heavy lifting stays in the CUDA kernels.
"""
import pickle
import zlib

import numpy as np
import torch

FIELDS = ("board", "pi", "winner", "scdiff", "valids", "surprise")


def finalize_examples(players, r_canon, score_canon, c_final):
    """Coach.py:89-98 for a batch of examples, as index arithmetic.
    players int64[E]: absolute seat that was to move when the example was recorded; r_canon float32[E,n], score_canon
    int32[E,n]: getGameEnded / getScore of the finished game in its final canonical frame (index 0 = absolute seat
    c_final[E]). Returns (winner float32[E,n], scdiff int32[E,n]) in each example's own frame (index 0 = its mover)."""
    n = r_canon.shape[1]
    i = torch.arange(n, device=r_canon.device).view(1, n)
    idx = (i + players.view(-1, 1) - c_final.view(-1, 1)) % n       # abs seat (i + p) sits at canonical index (i + p - c_final)
    winner = torch.gather(r_canon, 1, idx)
    sc = torch.gather(score_canon, 1, idx)
    return winner, sc - sc[:, :1]


class ExampleBuffer:
    """Per-lane staging of the examples of the game in progress + a flat list of finished-game examples (device tensors)."""

    def __init__(self, n_players, n_lanes, rows, device, max_per_game=None):
        self.n, self.T, self.R, self.device = n_players, n_lanes, rows, device
        self.M = max_per_game or 62 * n_players + 2
        M, T, A = self.M, n_lanes, 406
        self.board = torch.zeros((T, M, rows, 7), dtype=torch.int8, device=device)
        self.pi = torch.zeros((T, M, A), dtype=torch.float32, device=device)
        self.valids = torch.zeros((T, M, A), dtype=torch.uint8, device=device)
        self.surprise = torch.zeros((T, M, n_players), dtype=torch.float32, device=device)
        self.player = torch.zeros((T, M), dtype=torch.int64, device=device)
        self.count = torch.zeros(T, dtype=torch.int64, device=device)
        self.cur_player = torch.zeros(T, dtype=torch.int64, device=device)     # absolute seat to move (Coach.curPlayer)
        self._lane = torch.arange(T, device=device)
        self.finished = []   # list of dicts of tensors

    def record(self, boards, pi, valids, q, is_full):
        """one move of every lane: boards int8[T,R,7] canonical, pi float[T,406], valids uint8[T,406], q float[T,n],
        is_full bool[T] (only full searches are recorded, Coach.py:76)"""
        sel = is_full & (self.count < self.M)
        lanes = self._lane[sel]
        slot = self.count[sel]
        self.board[lanes, slot] = boards[sel]
        self.pi[lanes, slot] = pi[sel].to(torch.float32)
        self.valids[lanes, slot] = valids[sel]
        self.surprise[lanes, slot] = q[sel].to(torch.float32)
        self.player[lanes, slot] = self.cur_player[sel]
        self.count += sel.to(torch.int64)

    def advance(self, ended, scores, moved=None, absolute=False):
        """after the move: ended float32[T,n] (getGameEnded in the new canonical frame), scores int32[T,n] (getScore of
        the stored state, same frame). Finished lanes hand their examples over and start a new game at seat 0.
        moved bool[T] (None = all): the lanes that actually made a move in this call.
        absolute: ended / scores are in the absolute seat order (the boards are kept like Coach keeps them, Coach.py:86-98)."""
        n = self.n
        if moved is None:
            self.cur_player = (self.cur_player + 1) % n
            done = (ended != 0).any(dim=1)
        else:
            self.cur_player = torch.where(moved, (self.cur_player + 1) % n, self.cur_player)
            done = moved & (ended != 0).any(dim=1)
        return self.finish(done, ended, scores, absolute)

    def finish(self, done, ended, scores, absolute=False):
        """the lanes in `done` hand the examples of their game over (winner / score difference per Coach.py:89-98) and start anew"""
        if bool(done.any()):
            lanes = self._lane[done]
            cnt = self.count[lanes]
            m = torch.arange(self.M, device=self.device).view(1, -1) < cnt.view(-1, 1)          # [D, M]
            li = lanes.view(-1, 1).expand(-1, self.M)[m]
            si = torch.arange(self.M, device=self.device).view(1, -1).expand(lanes.numel(), -1)[m]
            c_final = torch.zeros_like(self.cur_player[li]) if absolute else self.cur_player[li]
            winner, scdiff = finalize_examples(self.player[li, si], ended[li], scores[li].to(torch.int32), c_final)
            self.finished.append(dict(board=self.board[li, si].clone(), pi=self.pi[li, si].clone(), winner=winner, scdiff=scdiff,
                                      valids=self.valids[li, si].clone(), surprise=self.surprise[li, si].clone()))
            self.count[lanes] = 0
            self.cur_player[lanes] = 0
        return done

    def drain(self):
        """-> dict of device tensors with all finished-game examples so far (and forgets them)"""
        if not self.finished:
            return empty_examples(self.n, self.R, self.device)
        out = {k: torch.cat([f[k] for f in self.finished]) for k in FIELDS}
        self.finished = []
        return out


def empty_examples(n, rows, device):
    return dict(board=torch.zeros((0, rows, 7), dtype=torch.int8, device=device), pi=torch.zeros((0, 406), dtype=torch.float32, device=device),
                winner=torch.zeros((0, n), dtype=torch.float32, device=device), scdiff=torch.zeros((0, n), dtype=torch.int32, device=device),
                valids=torch.zeros((0, 406), dtype=torch.uint8, device=device), surprise=torch.zeros((0, n), dtype=torch.float32, device=device))


def expand_symmetries(env, ex):
    """Coach.py:77-80: every example in all its symmetric variants (identity first), through the device kernel
    (SplendorLogicNumba.py:349-395). winner / scdiff / surprise are copied to every variant."""
    E = ex["board"].shape[0]
    if E == 0:
        return ex
    st, pi, va, cnt = env.symmetries(ex["board"], ex["pi"], ex["valids"])
    V = st.shape[1]
    keep = torch.arange(V, device=st.device).view(1, -1) < cnt.view(-1, 1).to(torch.int64)
    rep = lambda x: x.unsqueeze(1).expand(-1, V, *x.shape[1:])[keep]
    return dict(board=st[keep], pi=pi[keep], winner=rep(ex["winner"]), scdiff=rep(ex["scdiff"]), valids=va[keep], surprise=rep(ex["surprise"]))


def gather_examples(ex, group=None):
    """all ranks -> every rank holds the concatenation (rank order) of everybody's examples. Two collectives per field
    set: the counts, then one padded payload per field (NCCL for device tensors, gloo for host tensors)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ex
    world = dist.get_world_size(group)
    dev = ex["board"].device
    cnt = torch.tensor([ex["board"].shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    out = {}
    for k in FIELDS:
        x = ex[k]
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
        pad[: x.shape[0]] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)])
    return out


def to_coach_format(ex, compress=True):
    """-> the list Coach.executeEpisode returns (Coach.py:91-100): tuples (board int8[R,7], pi float32[406],
    winner float32[n], scdiff int[n], valids bool[406], surprise list[float]) - zlib+pickle'd unless compress=False"""
    host = {k: v.cpu().numpy() for k, v in ex.items()}
    out = []
    for i in range(host["board"].shape[0]):
        t = (host["board"][i], host["pi"][i], host["winner"][i], host["scdiff"][i].astype(np.int64), host["valids"][i].astype(np.bool_),
             [float(x) for x in host["surprise"][i]])
        out.append(zlib.compress(pickle.dumps(t), level=1) if compress else t)
    return out


def save_train_examples(history, folder, filename="checkpoint.examples"):
    """Coach.saveTrainExamples (Coach.py:167-173): `history` = trainExamplesHistory, a list (one entry per iteration) of
    lists / deques of examples in Coach's form (tuples, or their zlib+pickle bytes) -> pickle file the reference loads"""
    import os
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, filename)
    with open(path, "wb") as f:
        pickle.dump(history, f)
    return path


def load_train_examples(path, no_compression=False, num_iters_history=None, maxlen_of_queue=None):
    """Coach.loadTrainExamples (Coach.py:175-208): reads the pickled history, harmonises the compression of its items with
    `no_compression` (tuples <-> zlib+pickle bytes) and applies the two trims (latest iterations, items per iteration)"""
    with open(path, "rb") as f:
        hist = pickle.load(f)
    if hist and len(hist[0]) > 0:
        is_tuple = type(hist[0][0]) is tuple
        if is_tuple and not no_compression:
            hist = [type(h)(zlib.compress(pickle.dumps(x), level=1) for x in h) for h in hist]
        elif not is_tuple and no_compression:
            hist = [type(h)(pickle.loads(zlib.decompress(x)) for x in h) for h in hist]
    if num_iters_history is not None and len(hist) > num_iters_history:
        hist = hist[-num_iters_history:]
    if maxlen_of_queue is not None:
        for h in hist:
            while len(h) > maxlen_of_queue:
                h.pop()
    return hist


def from_coach_format(items, device="cpu"):
    """the inverse of to_coach_format: Coach's example tuples (or their zlib+pickle bytes) -> dict of tensors (FIELDS)"""
    rows = [pickle.loads(zlib.decompress(x)) if isinstance(x, (bytes, bytearray)) else x for x in items]
    if not rows:
        raise ValueError("no examples")
    f = lambda i, dt: torch.from_numpy(np.stack([np.asarray(r[i]) for r in rows]).astype(dt)).to(device)
    return dict(board=f(0, np.int8), pi=f(1, np.float32), winner=f(2, np.float32), scdiff=f(3, np.int32), valids=f(4, np.uint8), surprise=f(5, np.float32))
