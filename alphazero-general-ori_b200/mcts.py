"""MCTS on the GPU tree arena (C ABI: include/splendor_b200.h, "MCTS tree arena").

Two surfaces:

* `MCTSArena` - the batched engine: T trees (one per game lane) searched in lock-step waves
  (`begin` -> repeat `select` -> network -> `expand` -> `policy`). Every tree runs exactly the reference's sequential
  algorithm (MCTS.py:99-177); the batch is the number of trees. Leaves are handed to the network as device tensors
  (`leaf_states` int8[T,R,7], `leaf_valids` uint8[T,406]) with no copy - they are torch tensors, i.e. DLPack-exportable.
* `MCTS` - drop-in mirror of the reference class (MCTS.py:16-192): same constructor, `getActionProb(canonicalBoard, temp,
  force_full_search, bias) -> (list, list, bool)`, `search`, `reset_all_search_trees`, attributes `step`, `last_cleaning`,
  `rng`, so Coach.py:46,75,122 / Arena.py:171 / pit.py:54-91 drive it unchanged. It is an arena with one tree whose
  network is whatever `nnet.predict(board, valids)` the caller supplies (GenericNNetWrapper.py:141).

No CPU path: both raise without a CUDA device.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _native as nat
from .engine import rows


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class MCTSArena:
    def __init__(self, n_players, n_trees, node_cap, edge_cap=None, device=0, cpuct=1.0, fpu=0.0, temperature0=1.0,
                 dirichlet_alpha=0.3, seed=0, game_base=0, edge_reserve=32, gc_reachable=False, rounds=1, max_levels=0, token_limit=10,
                 rule_flags=nat.RULES_DEFAULT, pool_nodes=None, pool_bytes=None, leaves_per_tree=1):
        """node_cap: the most nodes ONE tree may hold (its hash table is sized for it). The node records themselves come from a
        page pool all trees share: `pool_nodes` records per tree on average (default node_cap, i.e. every tree can reach its
        limit at once; production sizes it for the average tree - a tree is retired whenever a real move reveals a card) with
        `edge_cap` edges per tree on average (default 24 per pooled node); or give `pool_bytes` directly.
        leaves_per_tree (1..4): simulations one tree may have in flight per wave. 1 = the reference's sequential search (the parity
        mode). More = virtual-loss leaf batching - NOT the reference's algorithm: a simulation in flight counts as a lost visit on
        the edges it walked, so one wave evaluates several different leaves of a tree; the leaf buffers then hold
        n_trees * leaves_per_tree rows (row = tree * leaves_per_tree + slot)."""
        if not torch.cuda.is_available():
            raise RuntimeError("MCTSArena needs a CUDA device (sm_100a); there is no CPU fallback")
        self.n, self.T = int(n_players), int(n_trees)
        self.K = int(leaves_per_tree)
        self.rows = self.T * self.K
        self.R, self.S = rows(self.n), 7 * rows(self.n)
        self.device = torch.device("cuda", device)
        self.node_cap = int(node_cap)
        self.pool_nodes = int(pool_nodes) if pool_nodes else self.node_cap
        self.edge_cap = int(edge_cap) if edge_cap else self.pool_nodes * 24
        self._lib = nat.lib()
        h = C.c_void_p()
        nat.check(self._lib.spl_ctx_create(self.n, token_limit, rule_flags, device, C.byref(h)))
        self._ctx = h
        if pool_bytes is None:
            pool_bytes = self.T * (self.pool_nodes * self._lib.spl_mcts_record_bytes(self.n, 0) + self.edge_cap * 24) + (2 * self.T + 64) * 32768
        self.pool_bytes = int(pool_bytes)
        nbytes = self._lib.spl_mcts_arena_bytes(self.n, self.T, self.node_cap, self.pool_bytes, self.K)
        if nbytes == 0:
            raise ValueError("MCTSArena: bad sizes (leaves_per_tree must be 1..4)")
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)      # spl_mcts_reset initialises what needs it
            off = (-self.arena.data_ptr()) % 256
            self._arena_ptr = self.arena.data_ptr() + off
            self.leaf_states = torch.zeros((self.rows, self.R, 7), dtype=torch.int8, device=self.device)
            self.leaf_valids = torch.zeros((self.rows, nat.NUM_ACTIONS), dtype=torch.uint8, device=self.device)
            self.leaf_flags = torch.zeros(self.rows, dtype=torch.uint8, device=self.device)
            self.counters = torch.zeros(2, dtype=torch.int32, device=self.device)
        m = C.c_void_p()
        nat.check(self._lib.spl_mcts_create(self._ctx, self.T, self.node_cap, self.pool_bytes, self.K, C.c_void_p(self._arena_ptr), nbytes, C.byref(m)))
        self._m = m
        self.arena_bytes = nbytes
        self.params = dict(cpuct=cpuct, fpu=fpu, temperature0=temperature0, dirichlet_alpha=dirichlet_alpha, seed=seed,
                           game_base=game_base, edge_reserve=edge_reserve, gc_reachable=int(bool(gc_reachable)), rounds=int(rounds), max_levels=int(max_levels))
        self.set_params()
        self.launches = 0
        self._nn_pending, self._nn_dir = None, None
        self.reset()

    def __del__(self):
        try:
            if getattr(self, "_m", None):
                self._lib.spl_mcts_destroy(self._m); self._m = None
            if getattr(self, "_ctx", None):
                self._lib.spl_ctx_destroy(self._ctx); self._ctx = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def wave_nnet_launches(self):
        """kernels of one spl_mcts_wave_nnet call: descent (+ the rules step inside it up to 6144 trees), attach, network; every further
        round (`rounds` > 1) is an (attach, descend, rules) pass more"""
        r = int(self.params.get("rounds", 1))
        if r <= 1:
            return 3 if (self.T <= 6144 and self.K == 1) else 4
        return 4 + 3 * (r - 1)

    def set_params(self, **kw):
        self.params.update(kw)
        p = nat.MctsParams(**{k: self.params[k] for k, _ in nat.MctsParams._fields_})
        nat.check(self._lib.spl_mcts_set_params(self._m, C.byref(p)))

    def set_episodes(self, episodes):
        """the lanes' episode counters (int32 / uint32 [T] device tensor, read at every begin): part of the key of the on-device
        Dirichlet sampler, so that every episode of a lane draws fresh root noise (the reference draws rng.dirichlet per move)"""
        assert episodes is None or (episodes.is_cuda and episodes.numel() == self.T and episodes.element_size() == 4)
        self._episodes = episodes
        nat.check(self._lib.spl_mcts_set_episodes(self._m, _ptr(episodes)))

    def pool_stats(self):
        """-> dict(pages, free, min_free, page_bytes) of the shared page pool (host sync)"""
        out = (C.c_int32 * 4)()
        nat.check(self._lib.spl_mcts_pool_stats(self._m, out, self._stream()))
        return dict(pages=int(out[0]), free=int(out[1]), min_free=int(out[2]), page_bytes=int(out[3]))

    def set_rules(self, token_limit=10, rule_flags=nat.RULES_DEFAULT):
        nat.check(self._lib.spl_ctx_set_rules(self._ctx, token_limit, rule_flags))

    # ------------------------------------------------------------------ the wave protocol
    def reset(self, tree_select=None):
        """reset_all_search_trees (MCTS.py:188-192)"""
        nat.check(self._lib.spl_mcts_reset(self._m, _ptr(tree_select), self._stream()))
        self.launches += 1

    def clean(self, fill_percent=50):
        """between waves: trees filled beyond fill_percent compact their pools (all in one launch)"""
        nat.check(self._lib.spl_mcts_clean(self._m, int(fill_percent), self._stream()))
        self.launches += 1

    def begin(self, roots, sims, move_flags=None, tree_select=None, dir_values=None):
        """roots int8[T,R,7] canonical boards (device), sims int32[T], move_flags uint8[T] (MCTS_MOVE_FORCED | MCTS_MOVE_NOISE),
        dir_values float64[T,406] (the Dirichlet vector per root, parity runs) or None (on-device Philox sampler)"""
        assert roots.dtype == torch.int8 and roots.is_cuda and roots.is_contiguous() and roots.numel() == self.T * self.S
        assert sims.dtype == torch.int32 and sims.numel() == self.T
        self._hold = (roots, sims, move_flags, tree_select, dir_values)
        nat.check(self._lib.spl_mcts_begin(self._m, _ptr(roots), _ptr(sims), _ptr(move_flags), _ptr(tree_select), _ptr(dir_values),
                                           self._stream()))
        self.launches += 1

    def select(self, count=False):
        """one selection wave (`rounds` x (descend, rules, attach) kernels)"""
        self.drain_nnet()
        nat.check(self._lib.spl_mcts_select(self._m, _ptr(self.leaf_states), _ptr(self.leaf_valids), _ptr(self.leaf_flags),
                                            _ptr(self.counters) if count else None, self._stream()))
        self.launches += 3 * self.params["rounds"]

    def expand(self, pi, v, dir_values=None):
        assert pi.dtype == torch.float32 and pi.is_contiguous() and pi.numel() == self.rows * nat.NUM_ACTIONS
        assert v.dtype == torch.float32 and v.is_contiguous() and v.numel() == self.rows * self.n
        nat.check(self._lib.spl_mcts_expand(self._m, _ptr(pi), _ptr(v), _ptr(dir_values), self._stream()))
        self.launches += 1

    def expand_select(self, pi, v, dir_values=None, count=False):
        """expand the previous wave's leaves and select the next ones (expansion + descent fused in one launch)"""
        nat.check(self._lib.spl_mcts_expand_select(self._m, _ptr(pi), _ptr(v), _ptr(dir_values), _ptr(self.leaf_states), _ptr(self.leaf_valids),
                                                   _ptr(self.leaf_flags), _ptr(self.counters) if count else None, self._stream()))
        self.launches += 3 * self.params["rounds"]

    def wave_steady(self, evaluator, dir_values=None):
        """steady-state wave for leaves that are already selected: network -> expand + next selection"""
        pi, v = evaluator(self.leaf_states, self.leaf_valids)
        self.expand_select(pi, v, dir_values)

    def wave_nnet(self, net, dir_values=None):
        """steady-state wave with the fused evaluator inside (spl_mcts_wave_nnet): expansion of the previous leaves + next
        descent -> rules -> {attach || network}. `net` = FusedSplendorNNet; its static output rows carry the network results
        from one wave to the next. Same results as `wave_steady`."""
        pi, v = net.out_buffers(self.rows)
        if self._nn_pending is not net:      # leaves were selected the classic way: their rows need the network first
            net(self.leaf_states, self.leaf_valids)
            self._nn_pending = net
        self._nn_dir = dir_values
        nat.check(self._lib.spl_mcts_wave_nnet(self._m, C.c_void_p(net.blob_ptr), _ptr(pi), _ptr(v), _ptr(dir_values), _ptr(self.leaf_states),
                                               _ptr(self.leaf_valids), _ptr(self.leaf_flags), None, self._stream()))
        self.launches += self.wave_nnet_launches

    def drain_nnet(self, dir_values=None):
        """after `wave_nnet`: back the pending network results up, so that the classic calls (select / expand / finish) can follow"""
        if self._nn_pending is not None:
            pi, v = self._nn_pending.out_buffers(self.rows)
            self._nn_pending = None
            self.expand(pi, v, dir_values if dir_values is not None else self._nn_dir)

    def remaining(self):
        """one (leaf-less) selection wave that counts the trees whose budget is not spent yet -> int (host sync)"""
        self.counters.zero_()
        self.select(count=True)
        return int(self.counters[1].item())

    def fixed_net(self, states=None, valids=None, pi=None, v=None):
        """the deterministic stand-in network (exact dyadic outputs) on device rows"""
        states = self.leaf_states if states is None else states
        valids = self.leaf_valids if valids is None else valids
        B = states.numel() // self.S
        if pi is None:
            pi = torch.empty((B, nat.NUM_ACTIONS), dtype=torch.float32, device=self.device)
        if v is None:
            v = torch.empty((B, self.n), dtype=torch.float32, device=self.device)
        nat.check(self._lib.spl_mcts_fixed_net(self._ctx, _ptr(states), _ptr(valids), B, _ptr(pi), _ptr(v), self._stream()))
        self.launches += 1
        return pi, v

    def search(self, roots, sims, evaluator, move_flags=None, dir_values=None, tree_select=None, waves=None):
        """one getActionProb for every tree: `evaluator(leaf_states, leaf_valids) -> (pi float32[T,406], v float32[T,n])`
        on the device. Runs max(sims) waves without host synchronisation (every wave finishes one simulation of nearly
        every unfinished tree), then keeps going until no tree has budget left (descents that crossed several
        transpositions or terminal nodes need an extra wave)."""
        self.begin(roots, sims, move_flags, tree_select, dir_values)
        if waves is None:
            waves = int(sims.max().item())
        waves = -(-waves // self.K)        # a wave finishes up to leaves_per_tree simulations of a tree
        if self._is_fused(evaluator):      # network inside the wave, next to the attach kernel (same results)
            self.select()
            for _ in range(waves):
                self.wave_nnet(evaluator, dir_values)
        else:
            for _ in range(waves):
                self.wave(evaluator, dir_values)
        self.finish(evaluator, dir_values)

    @staticmethod
    def _is_fused(evaluator):
        from .nnet import FusedSplendorNNet
        return isinstance(evaluator, FusedSplendorNNet)

    def wave(self, evaluator, dir_values=None):
        self.select()
        pi, v = evaluator(self.leaf_states, self.leaf_valids)
        self.expand(pi, v, dir_values)

    def finish(self, evaluator, dir_values=None, chunk=None, max_extra=100000):  # noqa: C901
        """waves until every tree has spent its budget. One host synchronisation per `chunk()` call (default: 8 waves);
        returns the number of extra waves."""
        extra = 0
        fused = chunk is None and self._is_fused(evaluator)
        while True:
            self.counters.zero_()
            self.select(count=True)
            left = int(self.counters[1].item())
            if left == 0:
                return extra
            if fused:
                for _ in range(8):
                    self.wave_nnet(evaluator, dir_values)
                extra += 8
                continue
            pi, v = evaluator(self.leaf_states, self.leaf_valids)
            self.expand(pi, v, dir_values)
            extra += 1
            if chunk is not None:
                extra += chunk()
            else:
                for _ in range(7):
                    self.wave(evaluator, dir_values)
                extra += 7
            if extra > max_extra:
                raise nat.NativeError("MCTS arena: search does not terminate")

    def get_action_prob_batch(self, boards, sims, evaluator, temp=1.0, move_flags=None, out=None):
        """MCTS.getActionProb for T games with HOST buffers: boards int8[T,R,7] (numpy / pinned tensor) in,
        (probs float64[T,406], q float64[T,n]) pinned host tensors out. `sims`: int or int32[T] device tensor."""
        if not hasattr(self, "_h_roots"):
            self._h_roots = torch.empty((self.T, self.R, 7), dtype=torch.int8).pin_memory()
            self._d_roots = torch.empty((self.T, self.R, 7), dtype=torch.int8, device=self.device)
            self._d_sims = torch.empty(self.T, dtype=torch.int32, device=self.device)
            self._h_probs = torch.empty((self.T, nat.NUM_ACTIONS), dtype=torch.float64).pin_memory()
            self._h_q = torch.empty((self.T, self.n), dtype=torch.float64).pin_memory()
        b = boards if torch.is_tensor(boards) else torch.from_numpy(np.ascontiguousarray(boards, dtype=np.int8))
        if b.data_ptr() != self._h_roots.data_ptr():
            self._h_roots.copy_(b.view(self.T, self.R, 7))
        self._d_roots.copy_(self._h_roots, non_blocking=True)
        if torch.is_tensor(sims):
            self._d_sims.copy_(sims)
            waves = None
        else:
            self._d_sims.fill_(int(sims))
            waves = int(sims)
        self.search(self._d_roots, self._d_sims, evaluator, move_flags, None, None, waves)
        probs, q = self.policy(temp)
        self._h_probs.copy_(probs, non_blocking=True)
        self._h_q.copy_(q, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_probs, self._h_q

    def policy(self, temp=1.0):
        """getActionProb's tail (MCTS.py:61-97) -> (probs float64[T,406], q float64[T,n])"""
        probs = torch.empty((self.T, nat.NUM_ACTIONS), dtype=torch.float64, device=self.device)
        q = torch.empty((self.T, self.n), dtype=torch.float64, device=self.device)
        nat.check(self._lib.spl_mcts_policy(self._m, float(temp), _ptr(probs), _ptr(q), self._stream()))
        self.launches += 1
        return probs, q

    def sample_moves(self, temp=1.0, episodes=None, actions=None, finished=None, counters=None):
        """for every tree whose budget is spent: one action drawn from its visit distribution (getActionProb's tail + the caller's
        np.random.choice, Coach.py:75-86) without materialising the probabilities -> (actions int16[T], -1 = still searching;
        finished uint8[T]); counters int64[2] (optional) += (simulations of the finished trees, finished trees); episodes: the lanes'
        episode counters (int32 / uint32 [T], part of the Philox key of the draw)"""
        if actions is None:
            actions = torch.empty(self.T, dtype=torch.int16, device=self.device)
        if finished is None:
            finished = torch.empty(self.T, dtype=torch.uint8, device=self.device)
        nat.check(self._lib.spl_mcts_sample_moves(self._m, float(temp), _ptr(episodes), _ptr(actions), _ptr(finished), _ptr(counters),
                                                  self._stream()))
        self.launches += 1
        return actions, finished

    def root_stats(self, want_arrays=True):
        """-> dict of device tensors: nsa int32[T,406], qsa float64[T,406], ps float32[T,406], and per-tree scalars"""
        T, A = self.T, nat.NUM_ACTIONS
        nsa = torch.empty((T, A), dtype=torch.int32, device=self.device) if want_arrays else None
        qsa = torch.empty((T, A), dtype=torch.float64, device=self.device) if want_arrays else None
        ps = torch.empty((T, A), dtype=torch.float32, device=self.device) if want_arrays else None
        info = torch.empty((T, nat.MCTS_INFO_WORDS), dtype=torch.int32, device=self.device)
        nat.check(self._lib.spl_mcts_root_stats(self._m, _ptr(nsa), _ptr(qsa), _ptr(ps), _ptr(info), self._stream()))
        self.launches += 1
        return dict(nsa=nsa, qsa=qsa, ps=ps, nodes=info[:, 0], edges=info[:, 1], ns=info[:, 2], sims_done=info[:, 3],
                    nn_calls=info[:, 4], status=info[:, 5] & 0xFF, truncated=info[:, 5] >> 8, resets=info[:, 6] >> 16, cleanings=info[:, 6] & 0xFFFF,
                    qs=info[:, 7].contiguous().view(torch.float32), last_v=info[:, 8:8 + self.n].contiguous().view(torch.float32), depth_sum=info[:, 12], spec_hits=info[:, 13],
                    dropped=info[:, 14], pages=info[:, 15])

    def check_status(self):
        st = self.root_stats(want_arrays=False)["status"]
        bad = int(st.max().item())
        if bad:
            raise nat.NativeError(f"MCTS arena: tree status bits {bad} (1 tree at its node limit, 2 shared page pool dry, 4 protocol): "
                                  f"raise node_cap / the pool (now {self.node_cap} nodes per tree, {self.pool_bytes} pool bytes)")


class MCTS:
    """Mirror of the reference's `MCTS` (MCTS.py:16-192) on a one-tree arena."""

    _instances = weakref.WeakSet()

    def __init__(self, game, nnet, args, dirichlet_noise=False, batch_info=None, node_cap=None, device=None):
        self.game = game
        self.nnet = nnet
        self.args = args
        self.dirichlet_noise = dirichlet_noise
        self.batch_info = batch_info
        self.rng = np.random.default_rng()
        self.step = 0
        self.last_cleaning = 0
        n = game.num_players
        sims = int(args.numMCTSSims)
        dev = device if device is not None else getattr(game, "_device", 0)
        board = getattr(game, "board", None)
        limit = getattr(board, "NUM_TOKEN_LIMIT", 10)
        flags = getattr(board, "_flags", nat.RULES_DEFAULT)
        temperature0 = 1.0
        if dirichlet_noise:
            temperature0 = float(args.temperature[0])
        try:      # pit.py's args (utils.dotdict: a missing key raises KeyError, not AttributeError) carry no dirichletAlpha
            alpha = float(args.dirichletAlpha or 0.3)
        except (AttributeError, KeyError):
            alpha = 0.3
        self._arena = MCTSArena(n, 1, node_cap or max(4096, 10 * sims), device=dev, cpuct=float(args.cpuct), fpu=float(args.fpu),
                                temperature0=temperature0, dirichlet_alpha=alpha,
                                token_limit=limit, rule_flags=flags)
        self._dev = self._arena.device
        self._root = torch.zeros((1, self._arena.R, 7), dtype=torch.int8, device=self._dev)
        self._sims = torch.zeros(1, dtype=torch.int32, device=self._dev)
        self._flags = torch.zeros(1, dtype=torch.uint8, device=self._dev)
        self._dir = torch.zeros((1, nat.NUM_ACTIONS), dtype=torch.float64, device=self._dev)
        self._pi = torch.zeros((1, nat.NUM_ACTIONS), dtype=torch.float32, device=self._dev)
        self._v = torch.zeros((1, n), dtype=torch.float32, device=self._dev)
        MCTS._instances.add(self)

    # ------------------------------------------------------------------
    @property
    def nodes_data(self):
        """the reference exposes its dictionary; here only its size is meaningful"""
        return range(int(self._arena.root_stats(want_arrays=False)["nodes"][0].item()))

    @nodes_data.setter
    def nodes_data(self, value):
        if not value:
            self._arena.reset()

    def _sync_rules(self):
        board = getattr(self.game, "board", None)
        if board is not None and hasattr(board, "_flags"):
            self._arena.set_rules(board.NUM_TOKEN_LIMIT, board._flags)

    def _run(self, canonicalBoard, nb, forced, noise):
        ar = self._arena
        self._sync_rules()
        self._root.copy_(torch.from_numpy(np.ascontiguousarray(canonicalBoard, dtype=np.int8)).view(1, ar.R, 7))
        self._sims.fill_(int(nb))
        self._flags.fill_((nat.MCTS_MOVE_FORCED if forced else 0) | (nat.MCTS_MOVE_NOISE if noise else 0))
        dirv = None
        if noise:   # applyDirNoise (:180-186): one value per legal action, drawn from the host generator like the reference
            k = int(np.count_nonzero(self.game.getValidMoves(canonicalBoard, 0)))
            d = np.zeros(nat.NUM_ACTIONS, dtype=np.float64)
            d[:k] = self.rng.dirichlet([self.args.dirichletAlpha] * k)
            self._dir.copy_(torch.from_numpy(d).view(1, -1))
            dirv = self._dir
        ar.begin(self._root, self._sims, self._flags, None, dirv)
        while True:
            ar.counters.zero_()
            ar.select(count=True)
            has_leaf, left = [int(x) for x in ar.counters.cpu()]
            if left == 0:
                break
            if not has_leaf:
                continue
            board = ar.leaf_states[0].cpu().numpy()
            valids = ar.leaf_valids[0].cpu().numpy().astype(np.bool_)
            if self.batch_info is None:
                ps, v = self.nnet.predict(board, valids)
            else:
                ps, v = self.nnet.predict_client(board, valids, self.batch_info)
            self._pi.copy_(torch.from_numpy(np.ascontiguousarray(ps, dtype=np.float32)).view(1, -1))
            self._v.copy_(torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).view(1, -1))
            ar.expand(self._pi, self._v, dirv)
        ar.check_status()

    def getActionProb(self, canonicalBoard, temp=1, force_full_search=False, bias=None):
        """MCTS.py:45-97 -> (probs list[float] len 406, q list[float] len n, is_full_search bool)"""
        is_full_search = bool(force_full_search or (self.rng.random() < self.args.prob_fullMCTS))
        nb = self.args.numMCTSSims if is_full_search else self.args.numMCTSSims // self.args.ratio_fullMCTS
        forced = bool(is_full_search and self.args.forced_playouts)
        noise = bool(is_full_search and self.dirichlet_noise)
        self._run(canonicalBoard, nb, forced, noise)
        self.step = max(0, int(nb) - 1)
        if temp == 0:   # :87-92 - a random one among the most visited
            st = self._arena.root_stats()
            counts = st["nsa"][0].cpu().numpy().astype(np.int64)
            if forced:
                probs1, _ = self._arena.policy(1.0)
                counts = probs1[0].cpu().numpy()
            best = np.flatnonzero(counts == counts.max())
            a = int(np.random.choice(best))
            probs = [0] * nat.NUM_ACTIONS
            probs[a] = 1
            _, q = self._arena.policy(1.0)
            return probs, [float(x) for x in q[0].cpu().numpy()], is_full_search
        probs, q = self._arena.policy(float(temp))
        return [float(x) for x in probs[0].cpu().numpy()], [float(x) for x in q[0].cpu().numpy()], is_full_search

    def search(self, canonicalBoard, dirichlet_noise=False, forced_playouts=False):
        """one simulation (MCTS.py:99-177; called directly by SplendorPlayers.py:178) -> float32[n]"""
        self._run(canonicalBoard, 1, bool(forced_playouts), bool(dirichlet_noise))
        return self._arena.root_stats(want_arrays=False)["last_v"][0].cpu().numpy().astype(np.float32)

    def root_stats(self):
        st = self._arena.root_stats()
        return {k: (v[0].cpu().numpy() if v is not None else None) for k, v in st.items()}

    @staticmethod
    def reset_all_search_trees():
        """MCTS.py:188-192"""
        for obj in list(MCTS._instances):
            obj._arena.reset()
            obj.last_cleaning = 0
