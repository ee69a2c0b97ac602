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
