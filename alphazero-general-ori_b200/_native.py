"""ctypes binding of libsplendor_b200.so (C ABI: include/splendor_b200.h).

The library is built in-tree by `build()` (nvcc, sm_100a only). There is no CPU fallback: if the
library is missing, or no CUDA device is present when a context is created, this raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.realpath(os.path.join(HERE, "..", "include"))
LIB = os.environ.get("SPL_B200_LIB") or os.path.join(HERE, "libsplendor_b200.so")   # SPL_B200_LIB: a variant build (experiments)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

NUM_ACTIONS = 406
MASK_WORDS = 13
LANE_TILE = 32
MAX_SYMMETRIES = 18
RULE_RESERVE, RULE_GIVEBACK, RULE_REFCOMPAT = 1, 2, 4
RULES_DEFAULT = 7
CHANCE_DETERMINISTIC, CHANCE_REPLAY, CHANCE_PHILOX = 0, 1, 2

EXPORTS = [
    "spl_abi_version", "spl_last_error", "spl_ctx_create", "spl_ctx_set_rules", "spl_ctx_set_tma", "spl_ctx_destroy",
    "spl_state_rows", "spl_state_bytes", "spl_lanes_padded", "spl_planes_bytes", "spl_mask_planes_bytes",
    "spl_pack", "spl_unpack", "spl_mask_unpack", "spl_reset_philox", "spl_reset_explicit", "spl_step", "spl_rollout",
    "spl_scores", "spl_symmetries",
    "spl_mcts_record_bytes", "spl_mcts_arena_bytes", "spl_mcts_create", "spl_mcts_set_episodes", "spl_mcts_pool_stats", "spl_mcts_destroy", "spl_mcts_set_params", "spl_mcts_reset", "spl_mcts_clean", "spl_mcts_begin",
    "spl_mcts_select", "spl_mcts_expand", "spl_mcts_expand_select", "spl_mcts_wave_nnet", "spl_mcts_debug_profile", "spl_mcts_policy", "spl_mcts_sample_moves", "spl_mcts_root_stats", "spl_mcts_fixed_net",
    "spl_nnet_blob_bytes", "spl_nnet_pack", "spl_nnet_forward", "spl_nnet_debug_stamps", "spl_nnet_debug_tile_stamps", "spl_nnet_debug_cta_times", "spl_umma_selftest", "spl_umma_selftest_mn", "spl_umma_mma_cycles", "spl_umma_stream_cycles",
]
MCTS_MOVE_FORCED, MCTS_MOVE_NOISE = 1, 2
MCTS_INFO_WORDS = 16


class StepArgs(C.Structure):
    _fields_ = [
        ("planes", C.c_void_p), ("n_lanes", C.c_int), ("actions", C.c_void_p), ("players", C.c_void_p),
        ("player", C.c_int), ("chance_mode", C.c_int), ("reveals", C.c_void_p), ("seed", C.c_uint64),
        ("game_base", C.c_uint32), ("episodes", C.c_void_p), ("rotate", C.c_int), ("auto_reset", C.c_int),
        ("store_state", C.c_int), ("mask_out", C.c_void_p), ("ended_out", C.c_void_p), ("next_actions", C.c_void_p),
        ("status_out", C.c_void_p), ("counters", C.c_void_p),
    ]


class RolloutArgs(C.Structure):
    _fields_ = [
        ("planes", C.c_void_p), ("n_lanes", C.c_int), ("plies", C.c_int), ("seed", C.c_uint64),
        ("game_base", C.c_uint32), ("episodes", C.c_void_p), ("players", C.c_void_p), ("rotate", C.c_int),
        ("first_plies", C.c_void_p), ("first_result", C.c_void_p), ("counters", C.c_void_p),
    ]


class MctsParams(C.Structure):
    _fields_ = [
        ("cpuct", C.c_double), ("fpu", C.c_double), ("temperature0", C.c_double), ("dirichlet_alpha", C.c_double),
        ("seed", C.c_uint64), ("game_base", C.c_uint32), ("edge_reserve", C.c_int), ("gc_reachable", C.c_int), ("rounds", C.c_int), ("max_levels", C.c_int),
    ]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h"))


def is_stale():
    return (not os.path.isfile(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in sources())


def build(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU"""
    if force or is_stale():
        units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
        objdir = os.path.join(HERE, "build")
        os.makedirs(objdir, exist_ok=True)
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
        procs = []
        for u in units:   # one nvcc per translation unit, in parallel
            obj = os.path.join(objdir, u[:-3] + ".o")
            procs.append((u, obj, subprocess.Popen(["nvcc"] + flags + ["-c", "-o", obj, os.path.join(CSRC, u)])))
        for u, _, p in procs:
            if p.wait() != 0:
                raise RuntimeError(f"nvcc failed on {u}")
        subprocess.check_call(["nvcc"] + NVCC_FLAGS + ["-o", LIB] + [obj for _, obj, _ in procs])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            raise RuntimeError(f"{LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is the product; there is no CPU fallback)")
        L = C.CDLL(LIB)
        vp, ci, u32, u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
        L.spl_abi_version.restype = ci
        L.spl_last_error.restype = C.c_char_p
        L.spl_ctx_create.argtypes = [ci, ci, u32, ci, C.POINTER(vp)]
        L.spl_ctx_set_rules.argtypes = [vp, ci, u32]
        L.spl_ctx_set_tma.argtypes = [vp, ci]
        L.spl_ctx_destroy.argtypes = [vp]
        L.spl_ctx_destroy.restype = None
        for f in ("spl_state_rows", "spl_state_bytes", "spl_lanes_padded"):
            getattr(L, f).argtypes = [ci]
            getattr(L, f).restype = ci
        L.spl_planes_bytes.argtypes = [ci, ci]
        L.spl_planes_bytes.restype = C.c_size_t
        L.spl_mask_planes_bytes.argtypes = [ci]
        L.spl_mask_planes_bytes.restype = C.c_size_t
        L.spl_pack.argtypes = [vp, vp, vp, ci, vp]
        L.spl_unpack.argtypes = [vp, vp, vp, ci, vp]
        L.spl_mask_unpack.argtypes = [vp, vp, vp, ci, vp]
        L.spl_reset_philox.argtypes = [vp, vp, ci, u64, u32, vp, vp, vp]
        L.spl_reset_explicit.argtypes = [vp, vp, ci, vp, vp, vp]
        L.spl_step.argtypes = [vp, C.POINTER(StepArgs), vp]
        L.spl_rollout.argtypes = [vp, C.POINTER(RolloutArgs), vp]
        L.spl_scores.argtypes = [vp, vp, ci, vp, vp, vp]
        L.spl_symmetries.argtypes = [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp]
        L.spl_mcts_record_bytes.argtypes = [ci, ci]
        L.spl_mcts_record_bytes.restype = C.c_size_t
        L.spl_mcts_arena_bytes.argtypes = [ci, ci, ci, C.c_size_t, ci]
        L.spl_mcts_arena_bytes.restype = C.c_size_t
        L.spl_mcts_create.argtypes = [vp, ci, ci, C.c_size_t, ci, vp, C.c_size_t, C.POINTER(vp)]
        L.spl_mcts_set_episodes.argtypes = [vp, vp]
        L.spl_mcts_pool_stats.argtypes = [vp, vp, vp]
        L.spl_mcts_destroy.argtypes = [vp]
        L.spl_mcts_destroy.restype = None
        L.spl_mcts_set_params.argtypes = [vp, C.POINTER(MctsParams)]
        L.spl_mcts_reset.argtypes = [vp, vp, vp]
        L.spl_mcts_clean.argtypes = [vp, ci, vp]
        L.spl_mcts_begin.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.spl_mcts_select.argtypes = [vp, vp, vp, vp, vp, vp]
        L.spl_mcts_expand.argtypes = [vp, vp, vp, vp, vp]
        L.spl_mcts_expand_select.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.spl_mcts_wave_nnet.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.spl_mcts_debug_profile.argtypes = [vp, vp]
        L.spl_mcts_policy.argtypes = [vp, C.c_double, vp, vp, vp]
        L.spl_mcts_sample_moves.argtypes = [vp, C.c_double, vp, vp, vp, vp, vp]
        L.spl_mcts_root_stats.argtypes = [vp, vp, vp, vp, vp, vp]
        L.spl_mcts_fixed_net.argtypes = [vp, vp, vp, ci, vp, vp, vp]
        L.spl_nnet_blob_bytes.argtypes = [ci]
        L.spl_nnet_blob_bytes.restype = C.c_size_t
        L.spl_nnet_pack.argtypes = [ci, C.POINTER(C.POINTER(C.c_float)), vp, C.c_size_t]
        L.spl_umma_selftest.argtypes = [vp, vp, vp, vp, ci, ci, vp, vp]
        L.spl_umma_selftest_mn.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp]
        L.spl_umma_stream_cycles.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp]
        L.spl_umma_mma_cycles.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp]
        L.spl_nnet_forward.argtypes = [vp, vp, vp, vp, ci, vp, vp, vp]
        _lib = L
    return _lib


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NativeError(f"libsplendor_b200: error {rc}: {lib().spl_last_error().decode()}")
