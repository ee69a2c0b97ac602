"""Batched Splendor environment on one B200: thin host object over the C ABI (include/splendor_b200.h).

All game state lives in HBM as lane tiles (int8[T][7R][32]); torch only owns the memory and the stream.
Methods mirror the reference's per-game calls for a whole batch of game lanes:

    reset        <- Board.init_game                  (SplendorLogicNumba.py:222-246)
    step         <- make_move + swap_players + check_end_game + valid_moves
                    = SplendorGame.getNextState / getCanonicalForm / getGameEnded / getValidMoves
                    (SplendorGame.py:30-57) fused in one launch
    rollout      <- the Arena.playGame loop with random players (Arena.py:99-160), many plies per launch
    states / set_states  <- Board.copy_state / Board.state   (:291-303)
"""
import ctypes as C

import torch

from . import _native as nat


def rows(n):
    """observation_size()[0] (SplendorLogicNumba.py:25-27)"""
    return 32 + 10 * n + n * n


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class SplendorEnv:
    def __init__(self, n_players, n_lanes, device=0, seed=0, game_base=0, token_limit=10,
                 rule_flags=nat.RULES_DEFAULT, use_tma=True):
        if not torch.cuda.is_available():
            raise RuntimeError("SplendorEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        self.n, self.L = int(n_players), int(n_lanes)
        self.device = torch.device("cuda", device)
        self.seed, self.game_base = int(seed), int(game_base)
        self.R, self.S = rows(self.n), 7 * rows(self.n)
        self._lib = nat.lib()
        h = C.c_void_p()
        nat.check(self._lib.spl_ctx_create(self.n, token_limit, rule_flags, device, C.byref(h)))
        self._ctx = h
        if not use_tma:
            nat.check(self._lib.spl_ctx_set_tma(self._ctx, 0))
        self.Lpad = self._lib.spl_lanes_padded(self.L)
        with torch.cuda.device(self.device):
            self.planes = torch.zeros(self.Lpad * self.S, dtype=torch.int8, device=self.device)
            self.masks = torch.zeros((nat.MASK_WORDS, self.Lpad), dtype=torch.int32, device=self.device)
            self.ended = torch.zeros((self.L, self.n), dtype=torch.float32, device=self.device)
            self.next_actions = torch.zeros(self.L, dtype=torch.int16, device=self.device)
            self.status = torch.zeros(self.L, dtype=torch.int32, device=self.device)
            self.episodes = torch.zeros(self.L, dtype=torch.int32, device=self.device)
            self.players = torch.zeros(self.L, dtype=torch.uint8, device=self.device)
            self.counters = torch.zeros(2, dtype=torch.int64, device=self.device)   # finished games, plies
        self.launches = 0   # kernels launched through this object (bench's gpu_launches)

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                self._lib.spl_ctx_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_rules(self, token_limit=10, rule_flags=nat.RULES_DEFAULT):
        """setNumTokenLim (:214) / disableReserve / enableReserve (SplendorGame.py:82-86)"""
        nat.check(self._lib.spl_ctx_set_rules(self._ctx, token_limit, rule_flags))

    def set_tma(self, enabled):
        nat.check(self._lib.spl_ctx_set_tma(self._ctx, int(bool(enabled))))

    # ------------------------------------------------------------------ state in / out
    def set_states(self, aos):
        """aos: int8[L, R, 7] (device or host) in the reference's layout"""
        aos = torch.as_tensor(aos).to(self.device, dtype=torch.int8, non_blocking=True).contiguous()
        assert aos.numel() == self.L * self.S
        nat.check(self._lib.spl_pack(self._ctx, _ptr(aos), _ptr(self.planes), self.L, self._stream()))
        self.launches += 1
        self._keep = aos

    def states(self, out=None, planes=None):
        """-> int8[L, R, 7] device tensor in the reference's layout"""
        if out is None:
            out = torch.empty((self.L, self.R, 7), dtype=torch.int8, device=self.device)
        nat.check(self._lib.spl_unpack(self._ctx, _ptr(self.planes if planes is None else planes), _ptr(out), self.L, self._stream()))
        self.launches += 1
        return out

    def valids(self, out=None):
        """mask planes of the last step -> uint8[L, 406] (what valid_moves returns, :251-265)"""
        if out is None:
            out = torch.empty((self.L, nat.NUM_ACTIONS), dtype=torch.uint8, device=self.device)
        nat.check(self._lib.spl_mask_unpack(self._ctx, _ptr(self.masks), _ptr(out), self.L, self._stream()))
        self.launches += 1
        return out

    def scores(self):
        """-> (int32[L, n] get_score, int32[L] get_round)"""
        sc = torch.empty((self.L, self.n), dtype=torch.int32, device=self.device)
        rd = torch.empty(self.L, dtype=torch.int32, device=self.device)
        nat.check(self._lib.spl_scores(self._ctx, _ptr(self.planes), self.L, _ptr(sc), _ptr(rd), self._stream()))
        self.launches += 1
        return sc, rd

    # ------------------------------------------------------------------ game start
    def reset(self, lane_select=None):
        """Philox game start for all lanes (or those with lane_select != 0), keyed (seed, game_base+lane, episode)"""
        sel = None if lane_select is None else lane_select.to(self.device, dtype=torch.uint8).contiguous()
        nat.check(self._lib.spl_reset_philox(self._ctx, _ptr(self.planes), self.L, self.seed, self.game_base,
                                             _ptr(self.episodes), _ptr(sel), self._stream()))
        self.launches += 1
        if lane_select is None:
            self.players.zero_()

    def reset_explicit(self, deals, nobles):
        """deals uint8[L,12] (colour*8+idx per visible slot), nobles uint8[L,5]: replay of a reference deal"""
        d = torch.as_tensor(deals).to(self.device, dtype=torch.uint8).contiguous()
        nb = torch.as_tensor(nobles).to(self.device, dtype=torch.uint8).contiguous()
        assert d.shape == (self.L, 12) and nb.shape == (self.L, 5)
        nat.check(self._lib.spl_reset_explicit(self._ctx, _ptr(self.planes), self.L, _ptr(d), _ptr(nb), self._stream()))
        self.launches += 1
        self.players.zero_()

    # ------------------------------------------------------------------ one ply for every lane
    def canonical(self, out=None, players=None, ended_out=None):
        """getCanonicalForm (SplendorGame.py:51-57) of every lane for its player to move (`players` uint8[L], default
        self.players) WITHOUT touching the stored states: a scratch copy of the lane tiles is rotated and unpacked.
        Also leaves getValidMoves(canonical, 0) in self.masks. For callers that keep the boards in the absolute seat order
        like the reference's Coach / Arena do (with more than two players the reference's rotation is not equivariant -
        SURVEY F7a -, so a board that is rotated after every move is not the board the reference would have)."""
        if getattr(self, "_scratch", None) is None:
            self._scratch = torch.empty_like(self.planes)
        self._scratch.copy_(self.planes)
        self.step(None, players=self.players if players is None else players, rotate=True, want_ended=ended_out is not None, want_status=False,
                  planes=self._scratch, ended_out=ended_out)
        return self.states(out, planes=self._scratch)

    def step(self, actions=None, players=None, player=0, chance="philox", reveals=None, rotate=False,
             auto_reset=False, store_state=True, want_mask=True, want_ended=True, want_next=False,
             want_status=True, count=False, planes=None, ended_out=None):
        """actions int16[L] (None / negative: query only). players uint8[L] or a single `player`.
        Returns nothing; results are in self.masks / self.ended / self.next_actions / self.status."""
        mode = {"det": nat.CHANCE_DETERMINISTIC, "deterministic": nat.CHANCE_DETERMINISTIC,
                "replay": nat.CHANCE_REPLAY, "philox": nat.CHANCE_PHILOX}[chance]
        a = nat.StepArgs()
        a.planes = (self.planes if planes is None else planes).data_ptr(); a.n_lanes = self.L
        self._hold = (actions, players, reveals)
        a.actions = None if actions is None else actions.data_ptr()
        a.players = None if players is None else players.data_ptr()
        a.player = int(player); a.chance_mode = mode
        a.reveals = None if reveals is None else reveals.data_ptr()
        a.seed = self.seed; a.game_base = self.game_base
        a.episodes = self.episodes.data_ptr()
        a.rotate = int(rotate); a.auto_reset = int(auto_reset); a.store_state = int(store_state)
        a.mask_out = self.masks.data_ptr() if want_mask else None
        a.ended_out = (self.ended if ended_out is None else ended_out).data_ptr() if want_ended else None
        a.next_actions = self.next_actions.data_ptr() if want_next else None
        a.status_out = self.status.data_ptr() if want_status else None
        a.counters = self.counters.data_ptr() if count else None
        nat.check(self._lib.spl_step(self._ctx, C.byref(a), self._stream()))
        self.launches += 1

    # ------------------------------------------------------------------ many plies per launch
    def rollout(self, plies, rotate=True, first_plies=None, first_result=None, count=True):
        a = nat.RolloutArgs()
        a.planes = self.planes.data_ptr(); a.n_lanes = self.L; a.plies = int(plies)
        a.seed = self.seed; a.game_base = self.game_base
        a.episodes = self.episodes.data_ptr()
        a.players = self.players.data_ptr()
        a.rotate = int(rotate)
        a.first_plies = None if first_plies is None else first_plies.data_ptr()
        a.first_result = None if first_result is None else first_result.data_ptr()
        a.counters = self.counters.data_ptr() if count else None
        nat.check(self._lib.spl_rollout(self._ctx, C.byref(a), self._stream()))
        self.launches += 1

    # ------------------------------------------------------------------ symmetries on AoS input
    def symmetries(self, aos, pi, valids):
        """get_symmetries (:349-395) for B states: returns (states int8[B,18,R,7], pi f32[B,18,406],
        valids uint8[B,18,406], count int32[B]); only the first count[b] variants of row b are defined"""
        aos = torch.as_tensor(aos).to(self.device, dtype=torch.int8).contiguous()
        B = aos.numel() // self.S
        pi = torch.as_tensor(pi).to(self.device, dtype=torch.float32).contiguous()
        va = torch.as_tensor(valids).to(self.device, dtype=torch.uint8).contiguous()
        M = nat.MAX_SYMMETRIES
        os_ = torch.zeros((B, M, self.R, 7), dtype=torch.int8, device=self.device)
        op = torch.zeros((B, M, nat.NUM_ACTIONS), dtype=torch.float32, device=self.device)
        ov = torch.zeros((B, M, nat.NUM_ACTIONS), dtype=torch.uint8, device=self.device)
        cnt = torch.zeros(B, dtype=torch.int32, device=self.device)
        nat.check(self._lib.spl_symmetries(self._ctx, _ptr(aos), _ptr(pi), _ptr(va), B, _ptr(os_), _ptr(op), _ptr(ov),
                                           _ptr(cnt), self._stream()))
        self.launches += 1
        return os_, op, ov, cnt
