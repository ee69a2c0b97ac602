"""TEST INFRASTRUCTURE - ctypes binding of oracle/liboracle.so (see splendor_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
ACTIONS = 406


class Rules(C.Structure):
    _fields_ = [("n_players", C.c_int), ("token_limit", C.c_int), ("enable_reserve", C.c_int),
                ("enable_giveback", C.c_int), ("ref_compat", C.c_int)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h", ".inc"))]
    if force or not os.path.isfile(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        i8p, u8p, f32p, i32p = (C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32))
        RP = C.POINTER(Rules)
        L.spo_rows.restype = C.c_int
        L.spo_default_rules.argtypes = [RP, C.c_int]
        L.spo_init_empty.argtypes = [i8p, RP]
        L.spo_deal_to_slot.argtypes = [i8p, RP, C.c_int, C.c_int, C.c_int]
        L.spo_set_noble.argtypes = [i8p, RP, C.c_int, C.c_int]
        L.spo_valid_moves.argtypes = [i8p, RP, C.c_int, u8p]
        L.spo_make_move.argtypes = [i8p, RP, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.spo_check_end_game.argtypes = [i8p, RP, f32p]
        L.spo_get_score.argtypes = [i8p, RP, C.c_int]
        L.spo_get_round.argtypes = [i8p]
        L.spo_swap_players.argtypes = [i8p, RP, C.c_int]
        L.spo_symmetries.argtypes = [i8p, RP, f32p, u8p, i8p, f32p, u8p]
        L.spo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.spo_philox_draw.argtypes = [i8p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.spo_init_philox.argtypes = [i8p, RP, C.c_uint64, C.c_uint32, C.c_uint32]
        L.spo_philox_pick.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.spo_rollout.argtypes = [RP, C.c_uint64, C.c_uint32, C.c_int, i32p, f32p]
        L.spo_rollout.restype = C.c_long
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def rows(n):
    return 32 + 10 * n + n * n


class Board:
    """One game on the CPU oracle. `state` is the reference's int8[R,7] array."""

    def __init__(self, n, ref_compat=True, token_limit=10, enable_reserve=True, enable_giveback=True):
        self.n = n
        self.rules = Rules(n, token_limit, int(enable_reserve), int(enable_giveback), int(ref_compat))
        self.state = np.zeros((rows(n), 7), dtype=np.int8)
        lib().spo_init_empty(_p(self.state, C.c_int8), self.rules)

    def copy(self):
        b = Board.__new__(Board)
        b.n, b.rules, b.state = self.n, Rules.from_buffer_copy(self.rules), self.state.copy()
        return b

    def set_state(self, st):
        self.state = np.ascontiguousarray(st, dtype=np.int8).copy()
        return self

    def init_empty(self):
        lib().spo_init_empty(_p(self.state, C.c_int8), self.rules)

    def init_explicit(self, deals, nobles):
        """deals: 12 (color, idx) pairs in slot order; nobles: n+1 noble ids"""
        self.init_empty()
        for slot, (c, i) in enumerate(deals):
            if lib().spo_deal_to_slot(_p(self.state, C.c_int8), self.rules, slot, int(c), int(i)):
                raise ValueError("card not in deck")
        for s, nid in enumerate(nobles):
            lib().spo_set_noble(_p(self.state, C.c_int8), self.rules, s, int(nid))

    def init_philox(self, seed, game, episode=0):
        lib().spo_init_philox(_p(self.state, C.c_int8), self.rules, seed, game, episode)

    def valid_moves(self, player):
        out = np.zeros(ACTIONS, dtype=np.uint8)
        lib().spo_valid_moves(_p(self.state, C.c_int8), self.rules, player, _p(out, C.c_uint8))
        return out.astype(np.bool_)

    def make_move(self, move, player, reveal=-1, seed=0, game=0, episode=0):
        """reveal: -1 deterministic, -2 philox, else color*8+idx"""
        return lib().spo_make_move(_p(self.state, C.c_int8), self.rules, int(move), int(player), int(reveal), seed, game, episode)

    def check_end_game(self):
        out = np.zeros(self.n, dtype=np.float32)
        lib().spo_check_end_game(_p(self.state, C.c_int8), self.rules, _p(out, C.c_float))
        return out

    def get_score(self, player):
        return lib().spo_get_score(_p(self.state, C.c_int8), self.rules, player)

    def get_round(self):
        return lib().spo_get_round(_p(self.state, C.c_int8))

    def swap_players(self, k):
        lib().spo_swap_players(_p(self.state, C.c_int8), self.rules, int(k))

    def symmetries(self, pi, valids):
        pi = np.ascontiguousarray(pi, dtype=np.float32)
        va = np.ascontiguousarray(valids, dtype=np.uint8)
        S = self.state.size
        os_ = np.zeros((20, S), dtype=np.int8); op = np.zeros((20, ACTIONS), dtype=np.float32); ov = np.zeros((20, ACTIONS), dtype=np.uint8)
        k = lib().spo_symmetries(_p(self.state, C.c_int8), self.rules, _p(pi, C.c_float), _p(va, C.c_uint8),
                                 _p(os_, C.c_int8), _p(op, C.c_float), _p(ov, C.c_uint8))
        return [(os_[i].reshape(self.state.shape).copy(), op[i].copy(), ov[i].astype(np.bool_)) for i in range(k)]

    def philox_draw(self, tier, seed, game, episode, ply, stream):
        return lib().spo_philox_draw(_p(self.state, C.c_int8), tier, seed, game, episode, ply, stream)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().spo_philox4x32_10(c, k, o)
    return list(o)


def philox_pick(valids, seed, game, episode, ply):
    va = np.ascontiguousarray(valids, dtype=np.uint8)
    return lib().spo_philox_pick(_p(va, C.c_uint8), seed, game, episode, ply)


def rollout(n, seed, game0, games, ref_compat=True):
    r = Rules(n, 10, 1, 1, int(ref_compat))
    plies = np.zeros(games, dtype=np.int32); res = np.zeros((games, n), dtype=np.float32)
    total = lib().spo_rollout(r, seed, game0, games, _p(plies, C.c_int32), _p(res, C.c_float))
    return total, plies, res


# ---------------------------------------------------------------------------------------------- search oracle (mcts_oracle.h)
class MoArgs(C.Structure):
    _fields_ = [("num_sims", C.c_int), ("ratio_full", C.c_int), ("forced_playouts", C.c_int), ("dirichlet_noise", C.c_int),
                ("cpuct", C.c_double), ("fpu", C.c_double), ("temperature0", C.c_double)]


PREDICT_FN = C.CFUNCTYPE(None, C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def _mo():
    L = lib()
    if not getattr(L, "_mo_ready", False):
        L.mo_create.restype = C.c_void_p
        L.mo_create.argtypes = [C.POINTER(Rules), C.POINTER(MoArgs), C.c_void_p, C.c_void_p]
        L.mo_destroy.argtypes = [C.c_void_p]
        L.mo_reset.argtypes = [C.c_void_p]
        L.mo_num_nodes.argtypes = [C.c_void_p]; L.mo_num_nodes.restype = C.c_long
        L.mo_nn_calls.argtypes = [C.c_void_p]; L.mo_nn_calls.restype = C.c_long
        L.mo_get_action_prob.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_double, C.c_int, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_float)]
        L.mo_search.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
        L.mo_fake_predict.argtypes = [C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L._mo_ready = True
    return L


class MCTSOracle:
    """the reference's MCTS (MCTS.py) restated in C; `predict(board int8[R,7], valids bool[406]) -> (Ps, v)` or None = fakenn"""

    def __init__(self, n, num_sims, cpuct=1.0, fpu=0.0, forced_playouts=False, dirichlet_noise=False, ratio_full=5,
                 temperature0=1.0, ref_compat=True, predict=None):
        self.n = n
        self.rules = Rules(n, 10, 1, 1, int(ref_compat))
        self.args = MoArgs(num_sims, ratio_full, int(forced_playouts), int(dirichlet_noise), cpuct, fpu, temperature0)
        self._cb = None
        if predict is not None:
            R = rows(n)

            def cb(sp, vp, pp, vvp, _u):
                st = np.ctypeslib.as_array(sp, shape=(R, 7)); va = np.ctypeslib.as_array(vp, shape=(ACTIONS,)).astype(np.bool_)
                ps, v = predict(st, va)
                np.ctypeslib.as_array(pp, shape=(ACTIONS,))[:] = ps
                np.ctypeslib.as_array(vvp, shape=(n,))[:] = v
            self._cb = PREDICT_FN(cb)
        self._h = _mo().mo_create(self.rules, self.args, C.cast(self._cb, C.c_void_p) if self._cb else None, None)

    def __del__(self):
        if getattr(self, "_h", None):
            _mo().mo_destroy(self._h); self._h = None

    def reset(self):
        _mo().mo_reset(self._h)

    @property
    def num_nodes(self):
        return _mo().mo_num_nodes(self._h)

    @property
    def nn_calls(self):
        return _mo().mo_nn_calls(self._h)

    def get_action_prob(self, canonical, temp=1.0, full_search=True, dir_values=None):
        st = np.ascontiguousarray(canonical, dtype=np.int8)
        probs = np.zeros(ACTIONS); q = np.zeros(self.n); nsa = np.zeros(ACTIONS, dtype=np.int64); qsa = np.zeros(ACTIONS)
        ns = C.c_long(); qs = C.c_float()
        d = None if dir_values is None else np.ascontiguousarray(dir_values, dtype=np.float64)
        rc = _mo().mo_get_action_prob(self._h, _p(st, C.c_int8), float(temp), int(full_search), None if d is None else _p(d, C.c_double),
                                      _p(probs, C.c_double), _p(q, C.c_double), _p(nsa, C.c_int64), _p(qsa, C.c_double),
                                      C.byref(ns), C.byref(qs))
        if rc:
            raise RuntimeError(f"mo_get_action_prob: {rc}")
        return dict(probs=probs, q=q, nsa=nsa, qsa=qsa, ns=ns.value, qs=np.float32(qs.value))


def fake_predict(state, valids, n):
    st = np.ascontiguousarray(state, dtype=np.int8); va = np.ascontiguousarray(valids, dtype=np.uint8)
    ps = np.zeros(ACTIONS, dtype=np.float32); v = np.zeros(n, dtype=np.float32)
    _mo().mo_fake_predict(_p(st, C.c_int8), st.size, _p(va, C.c_uint8), n, _p(ps, C.c_float), _p(v, C.c_float))
    return ps, v
