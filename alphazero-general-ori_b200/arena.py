"""Batched pit: `Arena.playGames` (Arena.py:175-227) for all games at once on one GPU.

The reference plays `num` games one after the other; game i seats player1 first when `i % 4 in (0, 3)` and player2 first
otherwise (the 1-2-2-1 order, Arena.py:199); with `other_way` the seats are `[player2] + [player1] * (n - 1)`, otherwise
`[player1, player2]` or `[player1, player2, player3]` (Arena.py:87-92). Every player owns its own search tree and answers
`argmax(getActionProb(canonical, temp=0, force_full_search=True))` (pit.py:91). A game's result is `getGameEnded(board)[0]`,
the outcome of SEAT 0, which counts for player1 or player2 depending on who sat there (Arena.py:201-207).

Here every game is a lane: all lanes make their k-th move together (seat k mod n), the lanes where agent A is to move search
in A's tree arena with A's network, the others in B's; the real move is the env kernel with its Philox reveal. Lanes whose
game is over idle until the last one finishes. Playout-cap randomisation (config 4 of BASELINE.json) is a per-lane coin.
"""
import torch

from . import _native as nat
from .engine import SplendorEnv
from .mcts import MCTSArena


class BatchedArena:
    def __init__(self, n_players, evaluators, num_sims, device=0, seed=0, game_base=0, cpuct=1.0, fpu=0.0, prob_full=1.0,
                 ratio_full=5, forced_playouts=False, node_cap=None, gc_reachable=False, third_is_second=True, pool_nodes=None):
        """evaluators: [player1's network, player2's network] (callables on device leaf rows). A third seat, if any, is
        played by player2 (`-p A -p B -p B`, pit.py:105-112). num_sims / cpuct / fpu / forced_playouts: one value for both
        players or a pair (each player of pit.py has its own MCTS args, pit.py:54-61)."""
        assert len(evaluators) == 2
        pair = lambda x: list(x) if isinstance(x, (list, tuple)) else [x, x]
        self.n, self.evaluators = n_players, evaluators
        self.num_sims_ab = [int(x) for x in pair(num_sims)]
        self.num_sims = max(self.num_sims_ab)
        self.device = torch.device("cuda", device)
        self.dev_index, self.seed, self.game_base = device, seed, game_base
        self.kw = [dict(cpuct=float(c), fpu=float(f), gc_reachable=gc_reachable, pool_nodes=pool_nodes) for c, f in zip(pair(cpuct), pair(fpu))]
        self.prob_full, self.ratio_full = float(prob_full), int(ratio_full)
        self.forced_ab = [bool(x) for x in pair(forced_playouts)]
        self.node_cap = node_cap or 8 * self.num_sims
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed) * 7919 + int(game_base))
        n = n_players
        self.seats_fwd = [0, 1] + [1] * (n - 2)     # [player1, player2(, player2 ...)]
        self.seats_rev = [1] + [0] * (n - 1)        # other_way: [player2] + [player1] * (n - 1)

    def play_games(self, num, max_moves=None, init_boards=None, reveals=None, record_actions=False):
        """-> (oneWon, twoWon, draws, details) like Arena.playGames (:175-227); details holds the per-game tensors.
        init_boards int8[num,R,7] / reveals int[num, moves] (colour*8+idx, negative = nothing drawn): replay of recorded games
        (their deals and revealed cards) instead of Philox chance - parity tests against the reference's Arena"""
        n, T, dev = self.n, int(num), self.device
        env = SplendorEnv(n, T, device=self.dev_index, seed=self.seed, game_base=self.game_base)
        env.reset()
        if init_boards is not None:
            env.set_states(torch.as_tensor(init_boards).to(dev, dtype=torch.int8).contiguous())
        if reveals is not None:
            rv = torch.as_tensor(reveals).to(dev, dtype=torch.int64)
            rv = torch.where(rv < 0, torch.full_like(rv, 255), rv).to(torch.uint8)
        absolute = n > 2      # boards in the absolute seat order like Arena.playGame keeps them (see SelfPlayEngine.absolute)
        ended_canon = torch.zeros((T, n), dtype=torch.float32, device=dev)
        log = []
        arenas = [MCTSArena(n, T, self.node_cap, device=self.dev_index, seed=self.seed + a, game_base=self.game_base, **self.kw[a]) for a in range(2)]
        i = torch.arange(T, device=dev)
        one_vs_two = ((i % 4) == 0) | ((i % 4) == 3)                          # Arena.py:199
        active = torch.ones(T, dtype=torch.bool, device=dev)
        result0 = torch.zeros(T, dtype=torch.float32, device=dev)             # getGameEnded(final board)[0]: seat 0's outcome
        moves = torch.zeros(T, dtype=torch.int32, device=dev)
        roots = torch.empty((T, env.R, 7), dtype=torch.int8, device=dev)
        sims = torch.empty(T, dtype=torch.int32, device=dev)
        flags = torch.empty(T, dtype=torch.uint8, device=dev)
        actions = torch.empty(T, dtype=torch.int16, device=dev)
        total_sims = 0
        k = 0
        limit = max_moves or 62 * n + n
        while k < limit and bool(active.any()):
            seat = k % n
            agent = torch.where(one_vs_two, self.seats_fwd[seat], self.seats_rev[seat])
            if absolute:
                env.canonical(roots, ended_out=ended_canon)
                # a rotated board whose end-of-game check fires although the board in seat order goes on (the reference's
                # Arena would crash in the player's search there): the game ends with the rotated board's result
                stuck = active & (ended_canon != 0).any(dim=1)
                r0s = torch.gather(ended_canon, 1, ((0 - env.players.to(torch.int64)) % n).view(-1, 1)).view(-1)
                result0 = torch.where(stuck, r0s, result0)
                active = active & ~stuck
            else:
                env.states(out=roots)
            is_full = torch.ones(T, dtype=torch.bool, device=dev) if self.prob_full >= 1.0 else \
                (torch.rand(T, device=dev, generator=self.gen) < self.prob_full)
            actions.fill_(-1)
            for a in range(2):
                sel = active & (agent == a)
                if not bool(sel.any()):
                    continue
                ar = arenas[a]
                sims.copy_(torch.where(is_full, self.num_sims_ab[a], max(1, self.num_sims_ab[a] // self.ratio_full)).to(torch.int32))
                flags.copy_(torch.where(is_full, nat.MCTS_MOVE_FORCED if self.forced_ab[a] else 0, 0).to(torch.uint8))
                ar.search(roots, sims, self.evaluators[a], flags, None, sel.to(torch.uint8), waves=self.num_sims_ab[a])
                probs, _ = ar.policy(0.0)                                   # temp = 0: one-hot of the most visited action
                act = probs.argmax(dim=1).to(torch.int16)
                actions.copy_(torch.where(sel, act, actions))
                total_sims += int(sims[sel].sum().item())
            if record_actions:
                log.append(actions.clone())
            col, chance = None, "philox"
            if reveals is not None:
                col = rv[:, k].contiguous() if k < rv.shape[1] else torch.full((T,), 255, dtype=torch.uint8, device=dev)
                chance = "replay"
            if absolute:
                env.step(actions, players=env.players, chance=chance, reveals=col, rotate=False, auto_reset=False, want_mask=False, want_status=True)
                env.players.copy_(torch.where((actions >= 0) & (env.status >= 0), env.status, env.players.to(torch.int32)).to(torch.uint8))
            else:
                env.step(actions, player=0, chance=chance, reveals=col, rotate=True, auto_reset=False, want_mask=False, want_status=False)
            moves += active.to(torch.int32)
            ended = env.ended
            c = (k + 1) % n                                                  # seat at canonical index 0 after the move
            r0 = ended[:, 0] if absolute else ended[:, (-c) % n]             # r_abs[0] = r_canon[(0 - c) mod n]
            done = active & (ended != 0).any(dim=1)
            result0 = torch.where(done, r0, result0)
            active = active & ~done
            k += 1
        for ar in arenas:
            ar.reset()                                                       # MCTS.reset_all_search_trees (Arena.py:171)
        one = ((result0 == 1.0) & one_vs_two) | ((result0 == -1.0) & ~one_vs_two)        # Arena.py:201-207
        two = ((result0 == -1.0) & one_vs_two) | ((result0 == 1.0) & ~one_vs_two)
        finished = ~active
        one_won, two_won = int((one & finished).sum()), int((two & finished).sum())
        draws = T - one_won - two_won
        return one_won, two_won, draws, dict(result_seat0=result0, one_vs_two=one_vs_two, moves=moves, unfinished=int(active.sum()),
                                             plies=k, total_sims=total_sims, actions=torch.stack(log, dim=1) if log else None)
