"""Batched self-play on one GPU: T game lanes, one MCTS tree per lane, everything resident on the device.

This is the loop `Coach.executeEpisode` runs one game at a time (Coach.py:67-100):

    canonical = getCanonicalForm(board, cur)                    -> the env keeps every lane canonical (rotate=True)
    pi, q, is_full = mcts.getActionProb(canonical, temp=1)      -> MCTSArena.search + policy for all lanes at once
    action = pick from pi                                       -> torch.multinomial on the device
    board, cur = getNextState(board, cur, action)               -> SplendorEnv.step (Philox deck reveals)
    r = getGameEnded(board, cur)                                -> same launch; finished lanes restart (auto reset)

Playout-cap randomisation (MCTS.py:54-56): each lane flips its own coin per move (full search with probability
prob_fullMCTS, else numMCTSSims // ratio_fullMCTS simulations without noise or forced playouts).
Games shard over GPUs by global game id (`game_base`); there is no collective on this path.
"""
import torch

from . import _native as nat
from .engine import SplendorEnv
from .examples import ExampleBuffer, expand_symmetries
from .mcts import MCTSArena


class SelfPlayEngine:
    def __init__(self, n_players, n_games, evaluator, num_sims, device=0, seed=0, game_base=0, cpuct=1.0, fpu=0.0,
                 prob_full=1.0, ratio_full=5, forced_playouts=False, dirichlet_noise=False, dirichlet_alpha=0.3,
                 temperature0=1.0, node_cap=None, edge_cap=None, gc_reachable=False, edge_reserve=24, graph_waves=0, rounds=1,
                 max_levels=0, record_examples=False, clean_every=0, clean_percent=50, overlap_nnet=None, tick_graph=None, pool_nodes=None,
                 absolute=None, leaves_per_tree=1):
        self.n, self.T, self.num_sims = n_players, n_games, int(num_sims)
        self.prob_full, self.ratio_full = float(prob_full), int(ratio_full)
        self.forced, self.noise = bool(forced_playouts), bool(dirichlet_noise)
        self._evaluator = None
        self._overlap_req = overlap_nnet      # None: whenever the evaluator is the fused kernel
        self.evaluator = evaluator
        self.env = SplendorEnv(n_players, n_games, device=device, seed=seed, game_base=game_base)
        self.device = self.env.device
        # How the boards are kept. Two players: always canonical (rotated after every move, one fused launch per move). More
        # players: in the absolute seat order with a player-to-move per lane, like Coach / Arena keep them, and canonical copies
        # are made for the search - the reference's rotation of the noble rows is not equivariant for n >= 3, so
        # only this reproduces the boards, end-of-game checks and examples of the reference's callers.
        self.absolute = (n_players > 2) if absolute is None else bool(absolute)
        # node_cap: the most nodes one tree may hold (a line of ~node_cap / num_sims moves without a revealed card); the shared page
        # pool is sized for the average tree, pool_nodes records per lane (default: node_cap, every tree at its limit at once)
        node_cap = node_cap or (3 if gc_reachable else 8) * self.num_sims
        self.arena = MCTSArena(n_players, n_games, node_cap, edge_cap, device=device, cpuct=cpuct, fpu=fpu, temperature0=temperature0,
                               dirichlet_alpha=dirichlet_alpha, seed=seed, game_base=game_base, edge_reserve=edge_reserve,
                               gc_reachable=gc_reachable, rounds=rounds, max_levels=max_levels, pool_nodes=pool_nodes, leaves_per_tree=leaves_per_tree)
        self.arena.set_episodes(self.env.episodes)      # the on-device Dirichlet sampler is keyed (seed, game, episode, ply)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed) * 1000003 + int(game_base))
        self.sims = torch.empty(n_games, dtype=torch.int32, device=self.device)
        self.flags = torch.empty(n_games, dtype=torch.uint8, device=self.device)
        self.roots = torch.empty((n_games, self.env.R, 7), dtype=torch.int8, device=self.device)
        self.actions = torch.empty(n_games, dtype=torch.int16, device=self.device)
        self.graph_waves = int(graph_waves)
        self.clean_every, self.clean_percent, self._ticks = int(clean_every), int(clean_percent), 0
        self._graph = None
        # asynchronous mode: replay the per-tick move logic as a CUDA graph (default: whenever the waves are captured)
        self.tick_graph = (self.graph_waves > 0) if tick_graph is None else bool(tick_graph)
        self._tick_graph, self._tick_state, self._tick_temp = None, 0, None
        self._fin8 = torch.zeros(n_games, dtype=torch.uint8, device=self.device)
        self.env.reset()
        self.examples = ExampleBuffer(n_players, n_games, self.env.R, self.device) if record_examples else None
        self.moves = 0
        self.extra_waves = 0
        self.games_finished = torch.zeros((), dtype=torch.int64, device=self.device)
        self.sims_total = torch.zeros((), dtype=torch.int64, device=self.device)   # simulations requested so far

    @property
    def evaluator(self):
        return self._evaluator

    @evaluator.setter
    def evaluator(self, ev):
        from .nnet import FusedSplendorNNet
        self._evaluator = ev
        fused = isinstance(ev, FusedSplendorNNet)
        self.overlap_nnet = fused if self._overlap_req is None else (bool(self._overlap_req) and fused)

    # ------------------------------------------------------------------
    def _roots(self):
        """canonical boards of all lanes -> self.roots (Coach.py:73)"""
        if not self.absolute:
            return self.env.states(out=self.roots)
        if not hasattr(self, "_ended_canon"):
            self._ended_canon = torch.zeros((self.T, self.n), dtype=torch.float32, device=self.device)
        self.env.canonical(self.roots, ended_out=self._ended_canon)
        # With the reference's n >= 3 quirks the end-of-game check of a rotated board can disagree with the check of the board in
        # seat order (get_score reads the noble rows with the wrong stride after a rotation). The reference's Coach
        # crashes there (the search of a finished position returns no visits); here such a game ends with the rotated board's result.
        stuck = (self._ended_canon != 0).any(dim=1)
        if self.examples is not None:
            i = torch.arange(self.n, device=self.device).view(1, -1)
            back = (i - self.env.players.view(-1, 1).to(torch.int64)) % self.n       # seat i sits at canonical index (i - p)
            scores, _ = self.env.scores()
            self.examples.finish(stuck, torch.gather(self._ended_canon, 1, back), scores, absolute=True)
        self._restart(stuck)
        self.env.canonical(self.roots)

    def _valids(self):
        """getValidMoves(canonical, 0) of all lanes (Coach.py:77)"""
        if not self.absolute:
            self.env.step(None, player=0, store_state=False, want_ended=False, want_status=False)
        return self.env.valids()          # (absolute: left there by env.canonical)

    def _move(self, chance="philox", reveals=None):
        """the real move of every lane (self.actions, negative = none): getNextState + getGameEnded (Coach.py:86-88)"""
        if self.absolute:
            pl = self.env.players
            self.env.step(self.actions, players=pl, chance=chance, reveals=reveals, rotate=False, auto_reset=False, want_mask=False,
                          want_status=True, count=True)
            st = self.env.status
            pl.copy_(torch.where((self.actions >= 0) & (st >= 0), st, pl.to(torch.int32)).to(torch.uint8))
        else:
            self.env.step(self.actions, player=0, chance=chance, reveals=reveals, rotate=True, auto_reset=False, want_mask=False,
                          want_status=False, count=True)

    def _restart(self, done):
        """finished lanes start their next game and forget their trees (Coach.py:67,122)"""
        done8 = done.to(torch.uint8)
        self.env.episodes += done.to(torch.int32)      # Philox key: game id, episode
        self.env.reset(done8)
        if self.absolute:
            self.env.players.masked_fill_(done, 0)
        self.arena.reset(done8)
        self.games_finished += done.sum()

    def _wave(self):
        if getattr(self, "_dir_values", None) is not None:      # parity runs: the Dirichlet vectors the caller supplies
            return self.arena.wave(self.evaluator, self._dir_values)
        if getattr(self, "_async", False):
            if self.overlap_nnet:
                self.arena.wave_nnet(self.evaluator)   # network next to the attach kernel (spl_mcts_wave_nnet)
            else:
                self.arena.wave_steady(self.evaluator)     # leaves are always selected one wave ahead
        else:
            self.arena.wave(self.evaluator)

    def _run_waves(self, waves):
        if self.graph_waves <= 0:
            for _ in range(waves):
                self._wave()
            return
        if self._graph is None:   # capture `graph_waves` waves once; replays reuse the arena's static leaf buffers
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._wave()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                for _ in range(self.graph_waves):
                    self._wave()
        for _ in range((waves + self.graph_waves - 1) // self.graph_waves):
            self._graph.replay()

    def search(self, temp=1.0, is_full=None, dir_values=None):
        """one getActionProb for every lane -> (probs float64[T,406], q float64[T,n], is_full bool[T]).
        is_full bool[T] / dir_values float64[T,406] (optional): the playout-cap coins and the root Dirichlet vectors of a recorded
        run instead of the engine's own draws (parity tests against Coach.executeEpisode)"""
        T = self.T
        self._dir_values = dir_values
        if is_full is not None:
            is_full = is_full.to(self.device, dtype=torch.bool)
            max_sims = self.num_sims
        elif self.prob_full >= 1.0:
            is_full = torch.ones(T, dtype=torch.bool, device=self.device)
            max_sims = self.num_sims
        else:
            is_full = torch.rand(T, device=self.device, generator=self.gen) < self.prob_full
            max_sims = self.num_sims
        self.sims.copy_(torch.where(is_full, self.num_sims, max(1, self.num_sims // self.ratio_full)).to(torch.int32))
        fl = (nat.MCTS_MOVE_FORCED if self.forced else 0) | (nat.MCTS_MOVE_NOISE if self.noise else 0)
        self.flags.copy_(torch.where(is_full, fl, 0).to(torch.uint8))
        self.sims_total += self.sims.sum()
        self._roots()
        if self.graph_waves > 0 and self._graph is None:
            # the warm-up waves of the capture are real waves: run them on a scratch search and throw its trees away, so that
            # the first move of every lane is exactly the reference's sequential search (no extra simulations, no second noise)
            self.arena.begin(self.roots, self.sims, self.flags)
            self._run_waves(0)
            self.arena.reset()
        self.arena.begin(self.roots, self.sims, self.flags, None, dir_values)
        self._run_waves(-(-max_sims // self.arena.K))
        chunk = None
        if self.graph_waves > 0:
            def chunk():
                self._graph.replay()
                return self.graph_waves
        self.extra_waves += self.arena.finish(self.evaluator, dir_values, chunk=chunk)   # stragglers (descents that crossed transpositions / terminal nodes)
        self._dir_values = None
        probs, q = self.arena.policy(temp)
        return self._usable_policy(probs), q, is_full

    def _usable_policy(self, probs):
        """getActionProb divides by the sum of the (pruned) visit counts (MCTS.py:69-76,94-97). With forced playouts, few simulations
        and many legal moves every move may have been visited once, the pruning (`c if c > 1 else 0`) then leaves nothing and the
        reference returns NaNs that crash its caller's np.random.choice. The engine keeps such a lane going on the raw visit counts."""
        bad = ~torch.isfinite(probs).all(dim=1) | ~(probs.sum(dim=1) > 0)
        nsa = self.arena.root_stats()["nsa"].to(probs.dtype)
        raw = nsa / nsa.sum(dim=1, keepdim=True).clamp_min(1.0)
        return torch.where(bad.view(-1, 1), raw, probs)

    def play_move(self, temp=1.0, is_full=None, dir_values=None, forced_actions=None, reveals=None):
        """search, sample an action per lane from the visit distribution, advance every game (finished lanes restart and
        their trees are cleared). Returns (probs, q, is_full, ended float32[T,n]).
        forced_actions int16[T] / reveals uint8[T] (colour*8+idx, 255 = nothing drawn): replay of a recorded game - the action
        Coach picked and the card the reference revealed - instead of the engine's own sample and Philox reveal"""
        probs, q, is_full = self.search(temp, is_full, dir_values)
        if forced_actions is not None:
            self.actions.copy_(forced_actions.to(self.device, dtype=torch.int16))
        else:
            a = torch.multinomial(probs.to(torch.float32), 1, generator=self.gen).view(-1)
            self.actions.copy_(a.to(torch.int16))
        if self.examples is not None:   # Coach.py:76-80: the position, its policy target and legal mask, before the move
            self.examples.record(self.roots, probs, self._valids(), q, is_full)
        if reveals is not None:
            self._reveals = reveals.to(self.device, dtype=torch.uint8).contiguous()
            self._move("replay", self._reveals)
        else:
            self._move()
        ended = self.env.ended
        done = (ended != 0).any(dim=1)
        if self.examples is not None:
            scores, _ = self.env.scores()
            self.examples.advance(ended, scores, absolute=self.absolute)
        self._restart(done)
        self.moves += 1
        return probs, q, is_full, ended

    def drain_examples(self, symmetries=True):
        """finished-game examples so far as device tensors (examples.FIELDS); with their symmetric variants (Coach.py:77-80)"""
        ex = self.examples.drain()
        return expand_symmetries(self.env, ex) if symmetries else ex

    # ------------------------------------------------------------------ asynchronous moves
    # In lock-step every move waits for the slowest tree (deep end-game lines need several waves per simulation and a
    # few hundred extra waves per move). Here every lane advances on its own: after each batch of waves the trees that
    # have spent their budget sample their action, make the real move and start the next search; the others keep going.
    # Each tree still runs exactly the reference's sequential search on its own positions.
    def start_async(self):
        T = self.T
        self._assign_budgets(torch.ones(T, dtype=torch.bool, device=self.device))
        self._roots()
        self.arena.begin(self.roots, self.sims, self.flags)
        self.arena.select()
        self._move_counters = torch.zeros(2, dtype=torch.int64, device=self.device)    # simulations, moves of the completed searches
        self._async = True

    def _assign_budgets(self, lanes):
        T = self.T
        if self.prob_full >= 1.0:
            is_full = torch.ones(T, dtype=torch.bool, device=self.device)
        else:
            is_full = torch.rand(T, device=self.device, generator=self.gen) < self.prob_full
        new_sims = torch.where(is_full, self.num_sims, max(1, self.num_sims // self.ratio_full)).to(torch.int32)
        fl = (nat.MCTS_MOVE_FORCED if self.forced else 0) | (nat.MCTS_MOVE_NOISE if self.noise else 0)
        new_flags = torch.where(is_full, fl, 0).to(torch.uint8)
        self.sims.copy_(torch.where(lanes, new_sims, self.sims))
        self.flags.copy_(torch.where(lanes, new_flags, self.flags))
        if not hasattr(self, "_is_full"):
            self._is_full = is_full.clone()
        else:
            self._is_full.copy_(torch.where(lanes, is_full, self._is_full))

    def tick(self, waves=None, temp=1.0):
        """`waves` selection waves (default: one graph replay) for every tree, then the lanes whose search is complete move on"""
        if self.graph_waves > 0:
            self._run_waves(waves or self.graph_waves)
        else:
            for _ in range(waves or 16):
                self._wave()
        self._ticks += 1
        if self.clean_every > 0 and self._ticks % self.clean_every == 0:
            self.arena.clean(self.clean_percent)      # every tree that needs it, in one launch, off the path of begin
        # the moves themselves: ~50 small launches (statistics, policy, sampling, env step, resets, begin). With captured waves
        # they are captured too (once, after one eager pass) and replayed as one graph: the launch gaps between them were ~10 %
        # of the run at 64 waves per tick. Example recording keeps the eager path (its buffers grow).
        if self.examples is None and self.tick_graph:
            if self._tick_state == 0 or self._tick_temp != temp:
                self._tick_tail(temp)
                self._tick_state, self._tick_temp, self._tick_graph = 1, temp, None
            else:
                if self._tick_graph is None:
                    torch.cuda.current_stream(self.device).synchronize()
                    self._tick_graph = torch.cuda.CUDAGraph()
                    self._tick_graph.register_generator_state(self.gen)
                    with torch.cuda.graph(self._tick_graph):
                        self._tick_tail(temp)
                self._tick_graph.replay()
        else:
            self._tick_tail(temp)

    def _tick_tail(self, temp):
        if self.examples is None:
            return self._tick_tail_fused(temp)
        st = self.arena.root_stats(want_arrays=False)
        fin = (st["sims_done"] >= self.sims) | (st["status"] != 0)
        probs, q = self.arena.policy(temp)
        probs = self._usable_policy(probs)
        p = probs.to(torch.float32)
        p = torch.where(fin.view(-1, 1) & (p.sum(dim=1, keepdim=True) > 0), p, torch.ones_like(p))
        a = torch.multinomial(p, 1, generator=self.gen).view(-1).to(torch.int16)
        self.actions.copy_(torch.where(fin, a, torch.full_like(a, -1)))
        if self.examples is not None:
            self.examples.record(self.roots, probs, self._valids(), q, self._is_full & fin)
        self._move()
        ended = self.env.ended
        done = fin & (ended != 0).any(dim=1)
        if self.examples is not None:
            scores, _ = self.env.scores()
            self.examples.advance(ended, scores, moved=fin, absolute=self.absolute)
        self._restart(done)
        self._move_counters[0] += torch.where(fin, st["sims_done"], torch.zeros_like(st["sims_done"])).sum()
        self._move_counters[1] += fin.sum()
        self._assign_budgets(fin)
        self._roots()
        self._fin8.copy_(fin)
        self.arena.begin(self.roots, self.sims, self.flags, self._fin8)

    def _tick_tail_fused(self, temp):
        """the same moves without materialising the policies (no examples are being recorded): spl_mcts_sample_moves draws every
        finished lane's action from its visit counts on the device (Philox keyed by game / episode / ply) and counts the simulations"""
        self.arena.sample_moves(temp, self.env.episodes, self.actions, self._fin8, self._move_counters)
        fin = self._fin8.to(torch.bool)
        self._move()
        done = fin & (self.env.ended != 0).any(dim=1)
        self._restart(done)
        self._assign_budgets(fin)
        self._roots()
        self.arena.begin(self.roots, self.sims, self.flags, self._fin8)

    @property
    def sims_completed(self):
        return self._move_counters[0]

    @property
    def moves_completed(self):
        return self._move_counters[1]

    def sims_in_flight(self):
        return self.arena.root_stats(want_arrays=False)["sims_done"].sum()
