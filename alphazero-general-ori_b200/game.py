"""Drop-in mirror of the reference's `SplendorGame` (SplendorGame.py:11-86; protocol Game.py:14-155).

Same method names, argument meaning and return types (numpy int8[R,7] boards, bool[406] masks,
float32[n] end vectors) so Coach / Arena / MCTS-style callers drive it unchanged; every call runs
the sm_100a kernels through the C ABI (one game lane for the per-game methods, L lanes for the
`*_batch` methods). No CPU path: constructing it without a CUDA device raises.

Differences that are part of the contract:
  * action space 406, pass (405) = no-op + ply counter
  * chance comes from Philox4x32-10 keyed (seed, game id, episode, ply), not Numba's MT19937;
    `getNextState(..., reveal=code)` replays a given draw (colour*8+idx) for bit-exact comparison
  * the n>=3 quirks of the reference are reproduced while `ref_compat=True` (default)
"""
import numpy as np
import torch

from . import _native as nat
from .engine import SplendorEnv, rows


def observation_size(num_players):
    """SplendorLogicNumba.py:25-27"""
    return (rows(num_players), 7)


def action_size():
    """SplendorLogicNumba.py:29-36 with patch P2"""
    return nat.NUM_ACTIONS


class Board:
    """Shim of the jitclass `Board` surface that callers reach through `game.board`
    (Arena.py:116,173; MCTS.py:161; SplendorPlayers.py). Holds one state as numpy int8[R,7]."""

    def __init__(self, num_players, seed=0, ref_compat=True, device=0):
        self.num_players = num_players
        self.n = num_players
        self._flags = nat.RULE_RESERVE | nat.RULE_GIVEBACK | (nat.RULE_REFCOMPAT if ref_compat else 0)
        self._limit = 10
        self._env = SplendorEnv(num_players, 1, device=device, seed=seed, token_limit=self._limit, rule_flags=self._flags)
        self._game_id = 0
        self.max_moves = np.uint8(62 * num_players)
        self.score_win = 15
        self.state = np.zeros(observation_size(num_players), dtype=np.int8)

    # ---- rule switches (Board fields :96-98, setter :214)
    @property
    def ENABLE_ACTION_RESERVE(self):
        return bool(self._flags & nat.RULE_RESERVE)

    @ENABLE_ACTION_RESERVE.setter
    def ENABLE_ACTION_RESERVE(self, v):
        self._flags = (self._flags | nat.RULE_RESERVE) if v else (self._flags & ~nat.RULE_RESERVE)
        self._env.set_rules(self._limit, self._flags)

    @property
    def ENABLE_ACTION_GIVEBACK(self):
        return bool(self._flags & nat.RULE_GIVEBACK)

    @ENABLE_ACTION_GIVEBACK.setter
    def ENABLE_ACTION_GIVEBACK(self, v):
        self._flags = (self._flags | nat.RULE_GIVEBACK) if v else (self._flags & ~nat.RULE_GIVEBACK)
        self._env.set_rules(self._limit, self._flags)

    @property
    def NUM_TOKEN_LIMIT(self):
        return self._limit

    def setNumTokenLim(self, n):
        self._limit = int(n)
        self._env.set_rules(self._limit, self._flags)

    # ---- state binding (:291-303)
    def get_state(self):
        return self.state

    def copy_state(self, state, copy_or_not):
        if self.state is state and not copy_or_not:
            return
        self.state = state.copy() if copy_or_not else state

    def _upload(self):
        self._env.set_states(torch.from_numpy(np.ascontiguousarray(self.state, dtype=np.int8)))

    def _download(self):
        self.state = self._env.states().cpu().numpy()[0]

    # ---- game start (:222-246)
    def init_game(self, game_id=None):
        if game_id is not None:
            self._game_id = int(game_id)
        self._env.game_base = self._game_id
        self._env.episodes.zero_()
        self._env.reset()
        self._download()
        self._game_id += 1

    # ---- rules
    def valid_moves(self, player):
        self._upload()
        self._env.step(None, player=int(player), store_state=False, want_ended=False, want_status=False)
        return self._env.valids().cpu().numpy()[0].astype(np.bool_)

    def make_move(self, move, player, deterministic, reveal=None):
        self._upload()
        act = torch.tensor([int(move)], dtype=torch.int16, device=self._env.device)
        rv = None
        if deterministic:
            chance = "det"
        elif reveal is not None:
            chance, rv = "replay", torch.tensor([int(reveal)], dtype=torch.uint8, device=self._env.device)
        else:
            chance = "philox"
        self._env.step(act, player=int(player), chance=chance, reveals=rv, want_mask=False, want_ended=False)
        status = int(self._env.status.cpu()[0])
        if status < 0:
            raise ValueError(f"move {move} is undefined in this state (status {status})")
        self._download()
        return status

    def check_end_game(self):
        self._upload()
        self._env.step(None, store_state=False, want_mask=False, want_status=False)
        return self._env.ended.cpu().numpy()[0].copy()

    def get_score(self, player):
        self._upload()
        return int(self._env.scores()[0].cpu()[0, int(player)])

    def get_round(self):
        return np.uint8(self.state[0, 6])   # :397-398 reads the int8 ply counter as uint8

    def swap_players(self, nb_swaps):
        self._upload()
        self._env.step(None, player=int(nb_swaps) % self.n, rotate=True, want_mask=False, want_ended=False, want_status=False)
        self._download()

    def get_symmetries(self, policy, valid_actions):
        os_, op, ov, cnt = self._env.symmetries(torch.from_numpy(np.ascontiguousarray(self.state)),
                                                torch.from_numpy(np.ascontiguousarray(policy, dtype=np.float32)),
                                                torch.from_numpy(np.ascontiguousarray(valid_actions).astype(np.uint8)))
        k = int(cnt.cpu()[0])
        os_, op, ov = os_.cpu().numpy()[0], op.cpu().numpy()[0], ov.cpu().numpy()[0]
        return [(os_[i].copy(), op[i].copy(), ov[i].astype(np.bool_)) for i in range(k)]


class SplendorGame:
    """Game protocol (Game.py:14-155) as implemented by the reference's SplendorGame (SplendorGame.py:11-86)."""

    def __init__(self, N, is_fill=True, seed=0, ref_compat=True, device=0):
        self.NUMBER_PLAYERS = N
        self.num_players = N
        self.board = Board(N, seed=seed, ref_compat=ref_compat, device=device)
        self._batch_envs = {}
        self._seed, self._ref_compat, self._device = seed, ref_compat, device

    def getInitBoard(self):
        self.board.init_game()
        return self.board.get_state()

    def getBoardSize(self):
        return observation_size(self.num_players)

    def getActionSize(self):
        return action_size()

    def getMaxScoreDiff(self):
        return 15

    def getNextState(self, board, player, action, deterministic=False, reveal=None):
        self.board.copy_state(board, True)
        next_player = self.board.make_move(action, player, deterministic, reveal=reveal)
        return (self.board.get_state(), next_player)

    def getValidMoves(self, board, player):
        self.board.copy_state(board, False)
        return self.board.valid_moves(player)

    def getGameEnded(self, board, next_player):
        self.board.copy_state(board, False)
        return self.board.check_end_game()

    def getScore(self, board, player):
        self.board.copy_state(board, False)
        return self.board.get_score(player)

    def getRound(self, board):
        self.board.copy_state(board, False)
        return self.board.get_round()

    def getCanonicalForm(self, board, player):
        if player == 0:
            return board   # the reference returns the same object (SplendorGame.py:52-53)
        self.board.copy_state(board, True)
        self.board.swap_players(player)
        return self.board.get_state()

    def getSymmetries(self, board, pi, valid_actions):
        self.board.copy_state(board, True)
        return self.board.get_symmetries(np.array(pi, dtype=np.float32), valid_actions)

    def stringRepresentation(self, board):
        return board.tobytes()

    def getNumberOfPlayers(self):
        return self.NUMBER_PLAYERS

    def moveToString(self, move, current_player):
        return move_to_str(move)

    def printBoard(self, numpy_board):
        """SplendorGame.py:72-75 -> print_board (SplendorLogic.py:600-607): round and scores, nobles, the three tiers with their deck
        sizes, the bank, then gems / cards / reserved cards of every player - the same content in plain text (no console colours)"""
        print(board_to_text(np.asarray(numpy_board, dtype=np.int8), self.num_players, self))

    def getNobleGemIDs(self, board):
        """SplendorGame.py:77-80 prints the nobles rows of `board.nobles` and returns None"""
        print("nobles: ", np.asarray(board.state if hasattr(board, "state") else board)[31:32 + self.num_players])
        return None

    def disableReserve(self):
        self.board.ENABLE_ACTION_RESERVE = False

    def enableReserve(self):
        self.board.ENABLE_ACTION_RESERVE = True

    # ------------------------------------------------------------------ batched forms (host buffers in, host buffers out)
    def _env_for(self, n_lanes):
        env = self._batch_envs.get(n_lanes)
        if env is None:
            env = SplendorEnv(self.num_players, n_lanes, device=self._device, seed=self._seed,
                              token_limit=self.board.NUM_TOKEN_LIMIT, rule_flags=self.board._flags)
            env._h_in = torch.empty((n_lanes, env.R, 7), dtype=torch.int8).pin_memory()
            env._h_act = torch.empty(n_lanes, dtype=torch.int16).pin_memory()
            env._h_out = torch.empty((n_lanes, env.R, 7), dtype=torch.int8).pin_memory()
            env._h_valid = torch.empty((n_lanes, nat.NUM_ACTIONS), dtype=torch.uint8).pin_memory()
            env._h_ended = torch.empty((n_lanes, self.num_players), dtype=torch.float32).pin_memory()
            env._d_in = torch.empty((n_lanes, env.R, 7), dtype=torch.int8, device=env.device)
            env._d_act = torch.empty(n_lanes, dtype=torch.int16, device=env.device)
            env._d_out = torch.empty((n_lanes, env.R, 7), dtype=torch.int8, device=env.device)
            env._d_valid = torch.empty((n_lanes, nat.NUM_ACTIONS), dtype=torch.uint8, device=env.device)
            self._batch_envs[n_lanes] = env
        return env

    def getNextStateBatch(self, boards, player, actions, deterministic=False, canonical=True, packed_masks=False):
        """L games at once, host arrays in and out: boards int8[L,R,7], actions int[L], all moved by `player`.
        Returns (next boards int8[L,R,7] (rotated to the next player's canonical form when `canonical`),
        valid masks bool[L,406] for the player to move, end vectors float32[L,n]) =
        getNextState + getCanonicalForm + getValidMoves + getGameEnded of the reference, per lane.
        packed_masks: the masks come back as their 406 bits - uint32[L,13], bit a of word a // 32 = action a (52 bytes per game
        instead of 406; `unpack_masks` expands them) - and the call runs as a pipeline of lane chunks on several streams, so that
        the copies in both directions and the kernels overlap."""
        boards = np.ascontiguousarray(boards, dtype=np.int8)
        L = boards.shape[0]
        if packed_masks:
            pipe = self._pipe_for(L)
            pipe.h_in.numpy()[...] = boards.reshape(L, -1)
            pipe.h_act.numpy()[...] = np.asarray(actions, dtype=np.int16)
            return pipe.step(int(player), deterministic, canonical)
        env = self._env_for(L)
        env._h_in.numpy()[...] = boards
        env._h_act.numpy()[...] = np.asarray(actions, dtype=np.int16)
        return self._step_pinned(env, player, deterministic, canonical)

    def rolloutBatch(self, boards, plies):
        """The loop of Arena.playGame with random players (Arena.py:99-160; SplendorPlayers.RandomPlayer) for L games and `plies` plies
        per call: canonical boards int8[L,R,7] in; every game plays uniformly random legal moves with Philox reveals, finished games
        restart; canonical boards out, plus (games finished, plies played). One host round trip per `plies` plies."""
        boards = np.ascontiguousarray(boards, dtype=np.int8)
        L = boards.shape[0]
        pipe = self._pipe_for(L)
        pipe.h_in.numpy()[...] = boards.reshape(L, -1)
        return pipe.rollout(int(plies))

    @staticmethod
    def unpack_masks(words):
        """uint32[L,13] mask words -> bool[L,406] as getValidMoves returns them"""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        return np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, :nat.NUM_ACTIONS].astype(np.bool_)

    def _pipe_for(self, n_lanes):
        key = ("pipe", n_lanes)
        pipe = self._batch_envs.get(key)
        if pipe is None:
            pipe = self._batch_envs[key] = _BatchPipe(self, n_lanes)
        return pipe

    def _step_pinned(self, env, player, deterministic, canonical):
        env._d_in.copy_(env._h_in, non_blocking=True)
        env._d_act.copy_(env._h_act, non_blocking=True)
        env.set_states(env._d_in)
        env.step(env._d_act, player=int(player), chance="det" if deterministic else "philox", rotate=canonical, want_status=False)
        env.states(out=env._d_out)
        env.valids(out=env._d_valid)
        env._h_out.copy_(env._d_out, non_blocking=True)
        env._h_valid.copy_(env._d_valid, non_blocking=True)
        env._h_ended.copy_(env.ended, non_blocking=True)
        torch.cuda.current_stream(env.device).synchronize()
        return env._h_out.numpy(), env._h_valid.numpy().view(np.bool_), env._h_ended.numpy()


class _BatchPipe:
    """Host-buffer calls for L games as a pipeline: the lanes are cut into chunks (multiples of the 32-lane tile), every chunk has
    its own stream and environment object; chunk c + 1 uploads while chunk c computes and chunk c - 1 downloads (PCIe is full
    duplex). Pinned host buffers for the whole batch, reused from call to call."""

    def __init__(self, game, n_lanes, chunks=None):
        n = game.num_players
        self.L, self.n = n_lanes, n
        dev = torch.device("cuda", game._device)
        C = chunks or max(1, min(8, n_lanes // 8192))
        per = -(-n_lanes // C)
        per = -(-per // 32) * 32
        self.bounds = [(lo, min(n_lanes, lo + per)) for lo in range(0, n_lanes, per)]
        S = 7 * rows(n)
        self.h_in = torch.empty((n_lanes, S), dtype=torch.int8).pin_memory()
        self.h_act = torch.empty(n_lanes, dtype=torch.int16).pin_memory()
        self.h_out = torch.empty((n_lanes, rows(n), 7), dtype=torch.int8).pin_memory()
        self.h_mask = torch.empty((n_lanes, nat.MASK_WORDS), dtype=torch.int32).pin_memory()
        self.h_ended = torch.empty((n_lanes, n), dtype=torch.float32).pin_memory()
        self.h_cnt = torch.zeros((len(self.bounds), 2), dtype=torch.int64).pin_memory()
        self.parts = []
        with torch.cuda.device(dev):
            for lo, hi in self.bounds:
                env = SplendorEnv(n, hi - lo, device=game._device, seed=game._seed, game_base=lo, token_limit=game.board.NUM_TOKEN_LIMIT,
                                  rule_flags=game.board._flags)
                d = dict(env=env, stream=torch.cuda.Stream(dev), lo=lo, hi=hi,
                         d_in=torch.empty((hi - lo, S), dtype=torch.int8, device=dev), d_act=torch.empty(hi - lo, dtype=torch.int16, device=dev),
                         d_out=torch.empty((hi - lo, rows(n), 7), dtype=torch.int8, device=dev),
                         d_mask=torch.empty((hi - lo, nat.MASK_WORDS), dtype=torch.int32, device=dev))
                self.parts.append(d)
        self.device = dev

    def step(self, player, deterministic, canonical):
        cur = torch.cuda.current_stream(self.device)
        for p in self.parts:
            st, env, lo, hi = p["stream"], p["env"], p["lo"], p["hi"]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                p["d_in"].copy_(self.h_in[lo:hi], non_blocking=True)
                p["d_act"].copy_(self.h_act[lo:hi], non_blocking=True)
                env.set_states(p["d_in"])
                env.step(p["d_act"], player=player, chance="det" if deterministic else "philox", rotate=canonical, want_status=False)
                env.states(out=p["d_out"])
                p["d_mask"].copy_(env.masks[:, : hi - lo].t())          # mask planes [13][lanes] -> words per game [lanes][13]
                self.h_out[lo:hi].copy_(p["d_out"], non_blocking=True)
                self.h_mask[lo:hi].copy_(p["d_mask"], non_blocking=True)
                self.h_ended[lo:hi].copy_(env.ended, non_blocking=True)
        for p in self.parts:
            p["stream"].synchronize()
        return self.h_out.numpy(), self.h_mask.numpy().view(np.uint32), self.h_ended.numpy()

    def rollout(self, plies):
        cur = torch.cuda.current_stream(self.device)
        for i, p in enumerate(self.parts):
            st, env, lo, hi = p["stream"], p["env"], p["lo"], p["hi"]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                p["d_in"].copy_(self.h_in[lo:hi], non_blocking=True)
                env.set_states(p["d_in"])
                env.counters.zero_()
                env.rollout(plies, rotate=True)
                env.states(out=p["d_out"])
                self.h_out[lo:hi].copy_(p["d_out"], non_blocking=True)
                self.h_cnt[i].copy_(env.counters, non_blocking=True)
        for p in self.parts:
            p["stream"].synchronize()
        c = self.h_cnt.numpy().sum(0)
        return self.h_out.numpy(), int(c[0]), int(c[1])

    def bytes_per_call(self, rollout=False):
        S = 7 * rows(self.n)
        if rollout:
            return self.L * S, self.L * S + 16 * len(self.parts)
        return self.L * (S + 2), self.L * (S + 4 * nat.MASK_WORDS + 4 * self.n)


_COLORS = ["white", "blue", "green", "red", "black"]
_SHORT = ["W", "U", "G", "R", "K", "gold"]


def _card_text(cost, gain):
    if int(gain[:5].sum()) == 0:
        return "-"
    col = _SHORT[int(np.flatnonzero(gain[:5] != 0)[0])]
    price = " ".join(f"{int(cost[c])}{_SHORT[c]}" for c in range(5) if cost[c] != 0)
    return f"[{col} {int(gain[6])}pt: {price}]"


def board_to_text(state, n, game=None):
    """the content of print_board (SplendorLogic.py:476-607) for a state in the reference's int8[R,7] layout"""
    R0 = 32 + n
    gems, nobles_p, cards, res = R0, R0 + n, R0 + 2 * n + n * n, R0 + 3 * n + n * n
    lines = []
    scores = [game.getScore(state, p) if game is not None else int(state[cards + p, 6]) for p in range(n)]
    lines.append(f"Round {int(np.uint8(state[0, 6])) // n} (ply {int(np.uint8(state[0, 6]))})  scores: " + "  ".join(f"P{p}={scores[p]}" for p in range(n)))
    nb = []
    for i in range(n + 1):
        row = state[31 + i]
        nb.append("< empty >" if row[6] == 0 else "< %d points %s >" % (int(row[6]), " ".join(f"{int(row[c])}{_SHORT[c]}" for c in range(5) if row[c] != 0)))
    lines.append("Nobles:  " + " ".join(nb))
    for tier in (2, 1, 0):
        deck = int(state[25 + 2 * tier, :5].sum())
        cs = [_card_text(state[1 + 8 * tier + 2 * i], state[2 + 8 * tier + 2 * i]) for i in range(4)]
        lines.append(f"Tier {tier} ({deck:2d} in deck): " + "  ".join(cs))
    lines.append("Bank:    " + " ".join(f"{int(state[0, c])}{_SHORT[c]}" for c in range(6)))
    for p in range(n):
        g = state[gems + p]
        own = [int(r[6]) for r in state[nobles_p + 3 * p: nobles_p + 3 * p + 3] if r[6] > 0]       # the reference reads a stride of 3 (:562)
        rs = [_card_text(state[res + 6 * p + 2 * r], state[res + 6 * p + 2 * r + 1]) for r in range(3) if int(state[res + 6 * p + 2 * r].sum()) != 0]
        lines.append(f"Player {p}: gems " + " ".join(f"{int(g[c])}{_SHORT[c]}" for c in range(6)) + f" (sum {int(g[:6].sum())})  cards " +
                     " ".join(f"{int(state[cards + p, c])}{_SHORT[c]}" for c in range(5)) + (f"  nobles {own}" if own else "") +
                     ("  reserved " + " ".join(rs) if rs else ""))
    return "\n".join(lines)


def move_to_str(move):
    """plain-text action names (the reference's coloured console rendering, SplendorLogic.py:60-223, is out of scope)"""
    if move < 12:
        return f"buy tier {move // 4} index {move % 4}"
    if move < 24:
        return f"reserve tier {(move - 12) // 4} index {(move - 12) % 4}"
    if move < 27:
        return f"reserve from deck of tier {move - 24}"
    if move < 30:
        return f"buy reserved card {move - 27}"
    if move < 60:
        return f"take gems (combination {move - 30})"
    if move < 290:
        return f"take and give back gems (exchange {move - 60})"
    if move < 365:
        return f"reserve {(move - 290) // 5} and give back 1 {_COLORS[(move - 290) % 5]}"
    if move < 405:
        return f"take 3 and give back 3 (exchange {move - 365})"
    return "do nothing"
