"""TEST-ONLY ctypes binding of the g++-compiled device rules core (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhostsim.so")
CSRC = os.path.realpath(os.path.join(HERE, "..", "..", "alphazero-general-ori_b200", "csrc"))

F_RESERVE, F_GIVEBACK, F_REFCOMPAT = 1, 2, 4
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [os.path.join(HERE, "hostsim.cpp"), os.path.join(CSRC, "spl_rules.cuh"), os.path.join(CSRC, "spl_tables.cuh")]
        if not os.path.isfile(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                                   "-o", LIB, os.path.join(HERE, "hostsim.cpp")])
        L = C.CDLL(LIB)
        i8p, u8p, u32p, f32p = C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
        L.hs_valid_mask.argtypes = [C.c_int, i8p, C.c_int, C.c_int, C.c_uint32, u32p]
        L.hs_apply_move.argtypes = [C.c_int, i8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.hs_game_ended.argtypes = [C.c_int, i8p, C.c_int, C.c_uint32, f32p]
        L.hs_rotate.argtypes = [C.c_int, i8p, C.c_int, C.c_uint32]
        L.hs_init_explicit.argtypes = [C.c_int, i8p, u8p, u8p]
        L.hs_init_philox.argtypes = [C.c_int, i8p, C.c_uint64, C.c_uint32, C.c_uint32]
        L.hs_pick_random.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.hs_score.argtypes = [C.c_int, i8p, C.c_int, C.c_uint32]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def unpack_mask(m13):
    return np.unpackbits(m13.view(np.uint8), bitorder="little")[:406].astype(np.bool_)


class Sim:
    def __init__(self, n, limit=10, flags=F_RESERVE | F_GIVEBACK | F_REFCOMPAT):
        self.n, self.limit, self.flags = n, limit, flags
        self.state = np.zeros((32 + 10 * n + n * n, 7), dtype=np.int8)

    def set_state(self, st):
        self.state = np.ascontiguousarray(st, dtype=np.int8).copy()
        return self

    def valid_words(self, player):
        m = np.zeros(13, dtype=np.uint32)
        lib().hs_valid_mask(self.n, _p(self.state, C.c_int8), player, self.limit, self.flags, _p(m, C.c_uint32))
        return m

    def valid_moves(self, player):
        return unpack_mask(self.valid_words(player))

    def make_move(self, a, player, reveal=-1, seed=0, game=0, episode=0):
        mode, code = (0, 0) if reveal == -1 else ((2, 0) if reveal == -2 else (1, reveal))
        return lib().hs_apply_move(self.n, _p(self.state, C.c_int8), int(a), int(player), mode, code, seed, game, episode)

    def check_end_game(self):
        out = np.zeros(self.n, dtype=np.float32)
        lib().hs_game_ended(self.n, _p(self.state, C.c_int8), self.limit, self.flags, _p(out, C.c_float))
        return out

    def swap_players(self, k):
        lib().hs_rotate(self.n, _p(self.state, C.c_int8), k, self.flags)

    def init_explicit(self, deals, nobles):
        d = np.ascontiguousarray(deals, dtype=np.uint8); nb = np.ascontiguousarray(nobles, dtype=np.uint8)
        lib().hs_init_explicit(self.n, _p(self.state, C.c_int8), _p(d, C.c_uint8), _p(nb, C.c_uint8))

    def init_philox(self, seed, game, episode=0):
        lib().hs_init_philox(self.n, _p(self.state, C.c_int8), seed, game, episode)

    def get_score(self, p):
        return lib().hs_score(self.n, _p(self.state, C.c_int8), p, self.flags)


def pick_random(words, seed, game, episode, ply):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    return lib().hs_pick_random(_p(w, C.c_uint32), seed, game, episode, ply)


# ---------------------------------------------------------------------------------------------- tree arena (csrc/spl_mcts.cuh)
LIB_MCTS = os.path.join(HERE, "libhostsim_mcts.so")
_lib_mcts = None
MCTS_F_FORCED, MCTS_F_NOISE = 1, 2


def lib_mcts():
    global _lib_mcts
    if _lib_mcts is None:
        deps = [os.path.join(HERE, "hostsim_mcts.cpp")] + [os.path.join(CSRC, f) for f in ("spl_mcts.cuh", "spl_rules.cuh", "spl_tables.cuh")]
        if not os.path.isfile(LIB_MCTS) or any(os.path.getmtime(d) > os.path.getmtime(LIB_MCTS) for d in deps):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-ffp-contract=off",
                                   "-o", LIB_MCTS, os.path.join(HERE, "hostsim_mcts.cpp")])
        L = C.CDLL(LIB_MCTS)
        L.hm_create.restype = C.c_void_p
        L.hm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.hm_destroy.argtypes = [C.c_void_p]
        L.hm_reset.argtypes = [C.c_void_p]
        L.hm_search.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_int, C.c_uint32, C.POINTER(C.c_double)]
        L.hm_policy.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.hm_root_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.hm_set_clean.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hm_set_episode.argtypes = [C.c_void_p, C.c_uint32]
        L.hm_codec_roundtrip.argtypes = [C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int8), C.POINTER(C.c_uint8)]
        L.hm_free_pages.argtypes = [C.c_void_p]
        L.hm_total_pages.argtypes = [C.c_void_p]
        L.hm_fixed_net.argtypes = [C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib_mcts = L
    return _lib_mcts


class TreeSim:
    """one tree of the arena driven on the host: getActionProb = begin + (select, fixed network, expand)*"""

    def __init__(self, n, num_sims, cpuct=1.0, fpu=0.0, forced_playouts=False, dirichlet_noise=False, ratio_full=5,
                 temperature0=1.0, cap=4096, ecap=None, edge_reserve=32, gc_reachable=False, limit=10, flags=F_RESERVE | F_GIVEBACK | F_REFCOMPAT):
        self.n, self.num_sims, self.forced, self.noise, self.ratio = n, num_sims, forced_playouts, dirichlet_noise, ratio_full
        # cap: node limit of the tree; the page pool holds cap records with ecap edges in all (default 48 per node)
        self._h = lib_mcts().hm_create(n, cap, ecap or cap * 48, limit, flags, cpuct, fpu, temperature0, edge_reserve, int(gc_reachable))   # gc_reachable: 0 exact rule, 1 reachable

    def __del__(self):
        if getattr(self, "_h", None):
            lib_mcts().hm_destroy(self._h)
            self._h = None

    def reset(self):
        lib_mcts().hm_reset(self._h)

    def free_pages(self):
        """(free, total) pages of the shared pool"""
        return lib_mcts().hm_free_pages(self._h), lib_mcts().hm_total_pages(self._h)

    def set_episode(self, e):
        lib_mcts().hm_set_episode(self._h, int(e))

    def set_clean(self, every, gc_reachable=False):
        """test hook: run the between-waves cleaning every `every` driver steps, whatever the search is doing"""
        lib_mcts().hm_set_clean(self._h, int(every), int(gc_reachable))

    def get_action_prob(self, canonical, temp=1.0, full_search=True, dir_values=None):
        st = np.ascontiguousarray(canonical, dtype=np.int8)
        sims = self.num_sims if full_search else self.num_sims // self.ratio
        fl = (MCTS_F_FORCED if (full_search and self.forced) else 0) | (MCTS_F_NOISE if (full_search and self.noise) else 0)
        d = None if dir_values is None else np.ascontiguousarray(dir_values, dtype=np.float64)
        status = lib_mcts().hm_search(self._h, _p(st, C.c_int8), sims, fl, None if d is None else _p(d, C.c_double))
        probs = np.zeros(406); q = np.zeros(self.n)
        lib_mcts().hm_policy(self._h, float(temp), _p(probs, C.c_double), _p(q, C.c_double))
        nsa = np.zeros(406, dtype=np.int32); qsa = np.zeros(406); ps = np.zeros(406, dtype=np.float32); info = np.zeros(16, dtype=np.int32)
        lib_mcts().hm_root_stats(self._h, _p(nsa, C.c_int32), _p(qsa, C.c_double), _p(ps, C.c_float), _p(info, C.c_int32))
        return dict(probs=probs, q=q, nsa=nsa.astype(np.int64), qsa=qsa, ps=ps, ns=int(info[2]), qs=info[7:8].view(np.float32)[0],
                    nodes=int(info[0]), edges=int(info[1]), dropped=int(info[14]), pages=int(info[15]), nn_calls=int(info[4]), status=status, truncated=int(info[5]) >> 8, resets=int(info[6]) >> 16,
                    compactions=int(info[6]) & 0xFFFF)


def codec_roundtrip(state, n):
    """compact node-state codec of the tree arena: -> (accepted, decoded state, compact bytes)"""
    st = np.ascontiguousarray(state, dtype=np.int8)
    back = np.zeros_like(st); cst = np.zeros(128, dtype=np.uint8)
    ok = lib_mcts().hm_codec_roundtrip(n, _p(st, C.c_int8), _p(back, C.c_int8), _p(cst, C.c_uint8))
    return bool(ok), back, cst


def fixed_net(state, valids, n):
    st = np.ascontiguousarray(state, dtype=np.int8); va = np.ascontiguousarray(valids, dtype=np.uint8)
    pi = np.zeros(406, dtype=np.float32); v = np.zeros(4, dtype=np.float32)
    lib_mcts().hm_fixed_net(n, _p(st, C.c_int8), _p(va, C.c_uint8), _p(pi, C.c_float), _p(v, C.c_float))
    return pi, v[:n]
