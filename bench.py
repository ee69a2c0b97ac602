#!/usr/bin/env python3
"""bench.py - the BASELINE.json metric "Splendor env steps/s & MCTS sims/s" on N B200s, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # headline: MCTS sims/s on configs[1] + env steps/s nested
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                            # the CPU arm (oracle port of MCTS.py + torch-CPU network)
    python bench.py --workload env ...                              # env-step throughput only (configs[4])

Headline workload = BASELINE.json configs[1]: 2-player self-play, 1600 MCTS simulations per move with SplendorNNet
(random-init weights), 4096 parallel games per GPU. A "step" = one move of every game: getActionProb for all lanes
(tree-arena begin + 1600 waves of select -> network -> expand/backup + policy), an action sampled from the visit
distribution and the real move with its Philox deck reveal. value = simulations per second over all ranks.
The env-step half of the metric (configs[4], one step = one launch of the persistent ply kernel over 1 Mi lanes) is
measured in the same run and reported under "env_steps", with its own roofline, e2e and CPU baseline.
Games shard over GPUs with no collective on the path (weak scaling: games per GPU fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC_MCTS, UNIT_MCTS = "splendor_mcts_sims_per_s", "sims/s"
METRIC_ENV, UNIT_ENV = "splendor_env_steps_per_s", "steps/s"
B_STEP = {2: 846, 3: 1060, 4: 1302}     # algorithmic bytes per env step: 2S + 52 + 2 + 4n
B_SIM = {2: 4500, 3: 5300, 4: 6300}     # algorithmic bytes per simulation (d=4.83, m=20)
NN_FLOPS_PER_LEAF = {2: 1.24e6, 3: 1.30e6, 4: 1.37e6}   # SplendorNNet forward, 2 x MACs (0.62 M MAC at n = 2; the first layer grows with R)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "mcts", "env"])
    ap.add_argument("--players", type=int, default=2)
    # MCTS (configs[1])
    ap.add_argument("--trees", type=int, default=18944, help="parallel games (trees) per GPU (18,944 = 148 SMs x 128: the evaluator's CTAs of 64 leaves fill two whole rounds)")
    ap.add_argument("--sims", type=int, default=1600, help="simulations per move")
    ap.add_argument("--nn-dtype", default="fused", choices=["fp32", "bf16", "fused"],
                    help="fp32 / bf16: torch evaluator; fused: the one-launch bf16 tensor-core kernel (csrc/spl_nnet.cu)")
    ap.add_argument("--graph-waves", type=int, default=128, help="waves per CUDA-graph replay (0: plain launches)")
    ap.add_argument("--tick-waves", type=int, default=0, help="waves between two rounds of moves (0: --graph-waves, or 16 with plain launches)")
    ap.add_argument("--gc", default="exact", choices=["exact", "reachable"],
                    help="tree cleaning: exact = drops only nodes no later search can look up again (ply + deck rule; the parity-tested mode), "
                         "reachable = a tree at its node limit keeps only what the root reaches")
    ap.add_argument("--pool-nodes", type=float, default=0, help="node records per tree the shared page pool is sized for (0: 4.8 x sims; measured mean 4.6 x sims at the sampling points)")
    ap.add_argument("--clean-moves", type=float, default=4.0, help="asynchronous mode: clean over-full trees every this many moves' worth of waves")
    ap.add_argument("--rounds", type=int, default=1, help="(descend, rules, attach) passes per selection wave")
    ap.add_argument("--async-moves", type=int, default=1, help="1: every lane moves on as soon as its own search is complete (no lock-step per move)")
    ap.add_argument("--max-levels", type=int, default=16, help="edges a descend call walks before it yields to the next wave (0: no limit)")
    ap.add_argument("--overlap", type=int, default=1, help="1: the fused network runs next to the attach kernel (spl_mcts_wave_nnet); 0: one stream")
    ap.add_argument("--e2e-max-levels", type=int, default=0, help="max_levels of the lock-step e2e leg (0: a descent never yields)")
    ap.add_argument("--e2e-rounds", type=int, default=1, help="(attach, descend, rules) passes per wave of the lock-step e2e leg: a tree whose descent crossed a "
                                                              "transposition or a terminal node still ends the wave with a leaf (measured at 16,384 trees: 2 passes cost more per wave than they save in waves)")
    ap.add_argument("--node-cap", type=int, default=0)
    ap.add_argument("--fixed-net", action="store_true", help="use the deterministic stand-in network instead of SplendorNNet")
    ap.add_argument("--opening-plies", type=int, default=24, help="random plies before the first search (mid-game positions)")
    ap.add_argument("--wide-trees", type=int, default=65536, help="second MCTS point: this many games per GPU (0: skip)")
    ap.add_argument("--wide-tick-waves", type=int, default=16, help="waves between two rounds of moves of the second MCTS point")
    ap.add_argument("--wide-sims", type=int, default=64, help="simulations per move of the second MCTS point")
    # env (configs[4])
    ap.add_argument("--lanes", type=int, default=1 << 20, help="game lanes per GPU")
    ap.add_argument("--plies", type=int, default=16, help="plies per lane per launch (rollout mode)")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "step"])
    ap.add_argument("--env-steps", type=int, default=20)
    ap.add_argument("--e2e-lanes", type=int, default=65536)
    ap.add_argument("--burn-in", type=int, default=300, help="untimed plies per lane before warm-up (de-synchronises the games)")
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    # CPU legs
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--cpu-mcts-worker", type=str, default="", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-ref-worker", type=str, default="", help=argparse.SUPPRESS)
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (config 2 with genbu.pt, configs[2], configs[3], float32 evaluator)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference (oracle/), all host cores
# ----------------------------------------------------------------------------------------------
def cpu_rollouts(n_players, seed, seconds, threads=None):
    """plays whole random games (same Philox policy as the GPU path) on `threads` host threads for about
    `seconds`; returns (steps/s, threads, games, plies, seconds)"""
    from oracle import pyoracle as po
    po.lib()
    threads = threads or len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    po.rollout(n_players, seed, 0, 200)
    per_game = (time.perf_counter() - t0) / 200
    games = max(200, int(seconds / per_game))
    results = [0] * threads

    def work(i):
        results[i], _, _ = po.rollout(n_players, seed, 1000 + i * games, games)   # ctypes releases the GIL

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    plies = int(sum(results))
    return plies / dt, threads, games * threads, plies, dt


def cpu_mcts_worker(spec):
    """one process = one core: the C port of MCTS.py (oracle/mcts_oracle.c) driving the network on the CPU through a
    per-leaf `predict` (torch float32, batch 1, one thread - GenericNNetWrapper.py:7,20,141-168), like the reference"""
    n, sims, moves, seed, idx, fixed = [int(x) for x in spec.split(",")]
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import pyoracle as po
    import azg_b200
    from azg_b200 import nnet
    predict = None
    if not fixed:
        W = nnet.fold(nnet.random_state_dict(n, seed), "cpu", torch.float32)

        def predict(board, valids):
            with torch.no_grad():
                pi, v = nnet.forward_folded(W, torch.from_numpy(board.copy()).view(1, -1, 7), torch.from_numpy(valids.astype(np.uint8)).view(1, -1))
            return pi[0].numpy(), v[0].numpy()
    m = po.MCTSOracle(n, sims, cpuct=1.0, fpu=0.0, predict=predict)
    b = po.Board(n); b.init_philox(seed, 100000 + idx)
    rng = np.random.default_rng(seed + idx)
    for _ in range(24):
        v = b.valid_moves(0)
        b.make_move(int(rng.choice(np.flatnonzero(v))), 0, -2, seed, 100000 + idx, 0); b.swap_players(1)
    m.get_action_prob(b.state, full_search=True)     # warm-up move (allocations, caches)
    done = 0
    t0 = time.perf_counter()
    for _ in range(moves):
        if b.check_end_game().any():
            b.init_philox(seed, 200000 + idx + done); m.reset()
        o = m.get_action_prob(b.state, full_search=True)
        done += sims
        a = int(rng.choice(406, p=o["probs"]))
        b.make_move(a, 0, -2, seed, 100000 + idx, 0); b.swap_players(1)
    dt = time.perf_counter() - t0
    print(json.dumps({"sims": done, "seconds": dt}), flush=True)


def reference_dir():
    """the reference's own files (patched copy, oracle/refgen/build_patched_ref.py): built from /root/reference where that exists,
    else the travelling copy oracle/_ref/pyref; None if neither is there"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refgen"))
    try:
        import build_patched_ref
        return build_patched_ref.find_ref(callers=True)
    except Exception:
        return None


def cpu_ref_worker(spec):
    """one process = one core: the REFERENCE ITSELF - its MCTS.py, its Numba Board behind SplendorGame, its NNetWrapper.predict
    (GenericNNetWrapper.py:141-168: torch CPU float32, batch 1, one thread) - playing `moves` moves of `sims` simulations"""
    n, sims, moves, seed, idx, _ = [int(x) for x in spec.split(",")]
    import warnings
    warnings.filterwarnings("ignore")
    t_jit = time.perf_counter()
    if reference_dir() is None:
        print(json.dumps({"unavailable": "no copy of the reference"}), flush=True)
        return
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from MCTS import MCTS
    from splendor.SplendorGame import SplendorGame
    from splendor.NNet import NNetWrapper
    from utils import dotdict
    import azg_b200  # noqa: F401  (only for the weights both arms share)
    from azg_b200 import nnet
    game = SplendorGame(n)
    net = NNetWrapper(game, dict(lr=0.001, dropout=0.3, epochs=1, batch_size=32, nn_version=1), use_exchange=True)
    net.nnet.load_state_dict(nnet.random_state_dict(n, seed), strict=True)
    args = dotdict(numMCTSSims=sims, prob_fullMCTS=1.0, ratio_fullMCTS=5, forced_playouts=False, cpuct=1.0, fpu=0.0, no_mem_optim=False)
    mcts = MCTS(game, net, args)
    rng = np.random.default_rng(seed + idx)
    board, cur = game.getInitBoard(), 0
    for _ in range(24):
        canon = game.getCanonicalForm(board, cur)
        v = game.getValidMoves(canon, 0)
        board, cur = game.getNextState(board, cur, int(rng.choice(np.flatnonzero(v))))
    mcts.args = dotdict(args, numMCTSSims=32)
    mcts.getActionProb(game.getCanonicalForm(board, cur), temp=1, force_full_search=True)     # JIT warm-up of every njit helper
    mcts.args = args
    MCTS.reset_all_search_trees()
    t_jit = time.perf_counter() - t_jit
    done = 0
    t0 = time.perf_counter()
    for _ in range(moves):
        if game.getGameEnded(board, cur).any():
            board, cur = game.getInitBoard(), 0
            MCTS.reset_all_search_trees()
        canon = game.getCanonicalForm(board, cur)
        pi, _, _ = mcts.getActionProb(canon, temp=1, force_full_search=True)
        done += sims
        board, cur = game.getNextState(board, cur, int(rng.choice(len(pi), p=np.array(pi) / np.sum(pi))))
    dt = time.perf_counter() - t0
    print(json.dumps({"sims": done, "seconds": dt, "jit_seconds": t_jit}), flush=True)


def cpu_mcts(n, sims, moves, seed, fixed, procs=None, worker="--cpu-mcts-worker"):
    """-> (sims/s aggregate, cores, total sims, wall seconds) with one worker process per host core"""
    procs = procs or len(os.sched_getaffinity(0))
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), worker, f"{n},{sims},{moves},{seed},{i},{int(fixed)}"],
                           stdout=subprocess.PIPE, env=env, text=True) for i in range(procs)]
    tot, rate, worst = 0, 0.0, 0.0
    for p in ps:
        out, _ = p.communicate()
        r = json.loads(out.strip().splitlines()[-1])
        if "unavailable" in r:
            return None
        tot += r["sims"]; rate += r["sims"] / r["seconds"]; worst = max(worst, r["seconds"])
    return rate, procs, tot, worst


def workload_mcts(args):
    net = "fixed stand-in network" if args.fixed_net else "SplendorNNet (random-init)"
    return (f"configs[1]: {args.players}p self-play, {args.sims} MCTS sims/move with {net}, {args.trees} parallel games per GPU "
            f"(the metric's '64k games' = 65,536 games over 4 or more GPUs; exact tree cleaning)")


def workload_env(args):
    return f"configs[4]: {args.players}p env-step throughput, random legal moves, {args.lanes} game lanes per GPU"


def run_reference(args):
    """the reference's own CPU implementation of the path on this box's host cores (oracle port; the reference is
    Python + Numba and cannot travel to the GPU box)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.players
    cores = len(os.sched_getaffinity(0))
    if args.workload == "env":
        per_step = max(0.5, min(10.0, 120.0 / max(1, args.steps + args.warmup)))
        for _ in range(args.warmup):
            cpu_rollouts(n, args.seed, min(per_step, 1.0))
        tot, tdt, games = 0, 0.0, 0
        for _ in range(args.steps):
            v, cores, g, plies, dt = cpu_rollouts(n, args.seed, per_step)
            tot += plies; tdt += dt; games += g
        kind = "port"
        value, metric, unit, wl = tot / tdt, METRIC_ENV, UNIT_ENV, workload_env(args)
        sample = f"{games} whole random {n}p games (oracle port of the SplendorLogicNumba rules, C -O2), {cores} threads"
    else:
        # step = one move (getActionProb, the full budget) per worker process, one process per host core
        sims = args.sims
        moves = max(1, args.steps)
        res, kind = None, "reference"
        if reference_dir() is not None:
            res = cpu_mcts(n, sims, moves, args.seed, args.fixed_net, worker="--cpu-ref-worker")
        if res is None:
            kind = "port"
            res = cpu_mcts(n, sims, moves, args.seed, args.fixed_net)
        rate, cores, tot, tdt = res
        value, metric, unit, wl = rate, METRIC_MCTS, UNIT_MCTS, workload_mcts(args)
        what = ("the reference's own MCTS.py + Numba Board + NNetWrapper.predict (torch CPU float32, batch 1)" if kind == "reference" else
                "C port of MCTS.py + per-leaf torch-CPU float32 SplendorNNet predict")
        sample = f"{cores} processes x {moves} moves x {sims} sims/move = {tot} simulations ({what}, 1 thread each, JIT / warm-up move excluded)"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tdt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32 tree statistics, f32 network" if metric == METRIC_MCTS else "int8", "data": "synthetic",
        "config": {"workload": wl, "players": n},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind if args.workload != "env" else "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields via NVML) during the timed region
# ----------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self.loop, daemon=True)

    def loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.th.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def load_traffic(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except Exception:
        return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------
# MCTS leg (configs[1])
# ----------------------------------------------------------------------------------------------
def bench_mcts(args, torch, dist, azg, world, rank, local, dev, barrier):
    n, T, sims = args.players, args.trees, args.sims
    torch.backends.cuda.matmul.allow_tf32 = False      # the float32 network is evaluated in float32
    torch.backends.cudnn.allow_tf32 = False
    if args.nn_dtype == "fused":
        net = azg.FusedSplendorNNet(n, seed=args.seed, device=local)
    else:
        net = azg.SplendorNNetB200(n, seed=args.seed, device=local, dtype=torch.float32 if args.nn_dtype == "fp32" else torch.bfloat16)
    reach = args.gc == "reachable"
    cap = args.node_cap or 20 * sims          # node limit of ONE tree: a line of ~20 moves without a revealed card
    pool_nodes = int(args.pool_nodes or 4.8 * sims)  # the shared page pool holds this many records per tree on average (32-bit offsets x 32 B: <= 128 GB)
    G0 = args.tick_waves or (args.graph_waves if args.graph_waves > 0 else 16)
    # exact mode: begin retires / cleans a tree when it has to; reachable mode: periodic cleaning of the over-full trees between waves
    clean_every = max(1, int(args.clean_moves * sims / G0)) if (args.async_moves and reach) else 0
    eng = azg.SelfPlayEngine(n, T, None, sims, device=local, seed=args.seed, game_base=rank * T, cpuct=1.0, fpu=0.0, node_cap=cap,
                             pool_nodes=pool_nodes, gc_reachable=reach, graph_waves=args.graph_waves, rounds=args.rounds, max_levels=args.max_levels,
                             clean_every=clean_every, clean_percent=45, overlap_nnet=None if args.overlap else False, tick_graph=True)
    if args.fixed_net:
        pi_buf = torch.empty((T, 406), dtype=torch.float32, device=dev); v_buf = torch.empty((T, n), dtype=torch.float32, device=dev)
        eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v, pi_buf, v_buf)
        nn_launches = 1
    else:
        eng.evaluator = net
        nn_launches = None
    eng.env.rollout(args.opening_plies, rotate=True)    # mid-game positions, lanes de-synchronised by the random openings

    W = max(args.warmup, 3)
    G = G0
    ticks_per_step = -(-sims // G)
    if args.async_moves:
        eng.start_async()

        def one_step():                      # a step = as many waves as a move has simulations, lanes advancing on their own
            for _ in range(ticks_per_step):
                eng.tick(G)

        def sims_now():
            return int(eng.sims_completed.item()) + int(eng.sims_in_flight().item())
    else:
        def one_step():
            eng.play_move()

        def sims_now():
            return int(eng.sims_total.item())
    for _ in range(W):
        one_step()
    barrier()
    eng.env.counters.zero_()
    sims0 = sims_now()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            one_step()
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    sims_done = sims_now() - sims0
    if not args.async_moves:
        assert sims_done == T * sims * args.steps
    sims_all = torch.tensor([sims_done], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(sims_all, op=dist.ReduceOp.SUM)       # what every rank really simulated, not world x rank 0
    sims_all = int(sims_all.item())
    st = eng.arena.root_stats(want_arrays=False)
    pool = eng.arena.pool_stats()
    truncated_now = 0 if args.async_moves else int((st["sims_done"] < sims).sum())
    value = sims_all / (ms_total * 1e-3)
    # our own kernels inside the timed region (graph replays re-launch the captured ones): per wave 3 x rounds selection
    # kernels + the evaluator + expand; per tick / move the stats, policy, env step, resets, unpack and begin kernels
    per_wave = 3 * args.rounds + 1 + (1 if (args.fixed_net or args.nn_dtype == "fused") else 0)
    if args.async_moves:
        # per wave: descent (with the expansion, and the rules step up to 6144 trees), attach, network; per tick: sample_moves, env step,
        # two resets, unpack, begin
        in_wave = eng.arena.wave_nnet_launches if eng.overlap_nnet else per_wave - 1
        own_launches_total = args.steps * ticks_per_step * (G * in_wave + 6)
    else:
        own_launches_total = args.steps * ((ticks_per_step * G + eng.extra_waves // max(1, W + args.steps)) * per_wave + 10)

    # ---- kernel breakdown of one wave, measured live with CUDA events on plain (non-graph) launches
    ar = eng.arena
    if args.async_moves:
        ar.finish(eng.evaluator)
    eng.env.states(out=eng.roots)
    ar.begin(eng.roots, eng.sims, eng.flags)
    nb = 200
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nb)]
    for i in range(nb):
        evs[i][0].record(); ar.select()
        evs[i][1].record(); pi, v = eng.evaluator(ar.leaf_states, ar.leaf_valids)
        evs[i][2].record(); ar.expand(pi, v, None)
        evs[i][3].record()
    torch.cuda.synchronize()
    sel = sum(e[0].elapsed_time(e[1]) for e in evs[20:]) / (nb - 20)
    nnt = sum(e[1].elapsed_time(e[2]) for e in evs[20:]) / (nb - 20)
    exp = sum(e[2].elapsed_time(e[3]) for e in evs[20:]) / (nb - 20)
    ovl = None
    if eng.overlap_nnet:      # the wave as the timed region runs it: network next to the attach kernel
        for _ in range(20):
            ar.wave_nnet(eng.evaluator)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nb):
            ar.wave_nnet(eng.evaluator)
        e1.record()
        torch.cuda.synchronize()
        ovl = e0.elapsed_time(e1) / nb
        ar.drain_nnet()
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # B_SIM covers a whole simulation (selection + backup + expansion + network I/O), so it is divided by the whole wave: the
    # one-call wave when the fused evaluator is in use, else selection + network + expansion
    wave_ms = ovl if ovl is not None else sel + nnt + exp
    achieved = B_SIM[n] * T / (wave_ms * 1e-3) / 1e9
    tr = load_traffic(f"mcts_wave_n{n}_bytes_per_sim")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None if tr is None else tr * T,
                "kernel": "steady-state wave: mcts_expand_descend_kernel (expansion + descent + rules step) -> nnet_forward_kernel || mcts_attach_kernel",
                "avg_launch_ms": wave_ms,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_sim": B_SIM[n],
                "note": "one launch = one simulation of every tree; the search is bound by the latency of sequential waves, not bytes",
                "wave_breakdown_ms": {"selection (descend+rules+attach kernels x rounds)": sel, "network_forward": nnt, "mcts_expand_kernel": exp,
                                      "whole wave, network next to the attach kernel (plain launches)": ovl}}

    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1409.8))
    nn_tflops = NN_FLOPS_PER_LEAF[n] * T / (nnt * 1e-3) / 1e12
    roofline_nn = {"bound": "tensor", "achieved": nn_tflops, "peak": bf16_peak, "unit": "TFLOP/s", "frac": nn_tflops / bf16_peak, "traffic": None,
                   "kernel": "nn2::nnet2_forward_kernel (one launch = the whole SplendorNNet forward for every tree's leaf)", "avg_launch_ms": nnt,
                   "flops_per_leaf": NN_FLOPS_PER_LEAF[n], "leaves_per_launch": T, "leaves_per_s": T / (nnt * 1e-3),
                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (a kernel timed inside a long step)" if peaks else "fallback 1409.8 TFLOP/s"}
    nn_name = {"fused": "bf16 network (bf16 x bf16 -> fp32 accumulate, tensor cores)", "bf16": "bf16 network (torch)", "fp32": "f32 network (torch)"}[args.nn_dtype]
    line = {
        "metric": METRIC_MCTS, "value": value, "unit": UNIT_MCTS, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"f64/f32 tree statistics, {nn_name}", "data": "synthetic",
        "config": {"workload": workload_mcts(args), "players": n, "trees_per_gpu": T, "sims_per_move": sims, "cpuct": 1.0, "fpu": 0.0,
                   "network": "fixed" if args.fixed_net else f"SplendorNNet random-init seed {args.seed} ({args.nn_dtype}, tf32 off)",
                   "gc": args.gc, "node_limit_per_tree": cap, "pool_nodes_per_tree": pool_nodes, "graph_waves": args.graph_waves, "waves_per_tick": G, "rounds_per_wave": args.rounds, "max_levels_per_descend": args.max_levels, "async_moves": bool(args.async_moves), "network_overlaps_attach": bool(eng.overlap_nnet),
                   "moves_completed": int(eng.moves_completed.item()) if args.async_moves else args.steps * T, "extra_waves": eng.extra_waves, "opening_plies": args.opening_plies,
                   "parallelism": f"games sharded dp{world}, no collective on the path",
                   "l2": f"inputs larger than L2: tree arena {eng.arena.arena_bytes / 1e9:.1f} GB per GPU vs 126 MB L2"},
        "roofline": roofline, "roofline_network": roofline_nn, "gpu_launches": own_launches_total, "wall_s": wall,
        "tree_stats": {"truncated_searches": int(st["truncated"].sum()) + truncated_now, "lossy_resets": int(st["resets"].sum()), "cleanings": int(st["cleanings"].sum()),
                       "mean_nodes": float(st["nodes"].float().mean()), "max_nodes": int(st["nodes"].max()),
                       "pool_gb": pool["pages"] * pool["page_bytes"] / 1e9, "pool_peak_fill": 1.0 - pool["min_free"] / max(1, pool["pages"]), "mean_path_length": float(st["depth_sum"].sum()) / max(1.0, float(sims_now())), "mean_edges_per_node": float(st["edges"].sum()) / max(1.0, float(st["nodes"].sum())),
                       "early_fetch_hit_rate": float(st["spec_hits"].sum()) / max(1.0, float(st["depth_sum"].sum())),
                       "games_finished": int(eng.games_finished.item()), "network_rows_per_sim": float(st["nn_calls"].sum()) / max(1, sims_now())},
    }
    if rank == 0:
        line["clocks"] = clk.summary()

    # ---- e2e: the Coach-style loop through the public API with HOST buffers (boards in, probs out, boards out)
    game = azg.SplendorGame(n, seed=args.seed, device=local)
    import numpy as np
    boards = eng.env.states().cpu().numpy()
    ar.reset()
    ar.set_params(max_levels=args.e2e_max_levels, rounds=args.e2e_rounds)     # lock-step calls wait for the slowest tree (0: no yielding inside a descent)
    ke = 3
    h2d = d2h = 0
    launches0 = ar.launches
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        probs, q = ar.get_action_prob_batch(boards, sims, eng.evaluator)
        acts = probs.numpy().argmax(1).astype(np.int16)
        boards, valids, ended = game.getNextStateBatch(boards, 0, acts)
        boards = boards.copy()
        h2d += T * (eng.env.S + eng.env.S + 2); d2h += T * (406 * 8 + n * 8 + eng.env.S + 406 + 4 * n)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line["e2e"] = {"value": world * T * sims * ke / float(t.item()), "unit": UNIT_MCTS, "h2d_bytes_per_step": h2d // ke, "d2h_bytes_per_step": d2h // ke,
                   "call": "MCTSArena.get_action_prob_batch(host boards) -> host probs/q, then SplendorGame.getNextStateBatch(host boards, actions)",
                   "steps": ke, "waves_per_move": (ar.launches - launches0) / float(ar.wave_nnet_launches if eng.overlap_nnet else 5) / ke,
                   "note": "lock-step: every wave lasts as long as the deepest descent of any tree and the call returns when the slowest tree "
                           "has spent its budget (its ~1600 sequential simulations bound the call from below)"}
    if world > 1:
        line["example_exchange"] = check_example_exchange(args, torch, dist, azg, world, rank, local)
    return line, eng


def check_example_exchange(args, torch, dist, azg, world, rank, local):
    """multi-GPU only: a short recorded self-play on every rank, then the per-iteration all-gather of the
    finished-game examples over NCCL; every rank must end up with the same concatenation"""
    n, T = args.players, 256
    net = azg.FusedSplendorNNet(n, seed=args.seed, device=local)
    eng = azg.SelfPlayEngine(n, T, net, 8, device=local, seed=args.seed, game_base=rank * T, node_cap=128, record_examples=True)
    eng.env.rollout(40 * n, rotate=True)
    eng.examples.cur_player.zero_()
    for _ in range(40):
        eng.play_move()
    mine = eng.drain_examples(symmetries=True)
    allx = azg.examples.gather_examples(mine)
    local_n, total = int(mine["board"].shape[0]), int(allx["board"].shape[0])
    chk = torch.tensor([float(allx["pi"].double().sum().item()), float(allx["board"].double().abs().sum().item())], dtype=torch.float64,
                       device=f"cuda:{local}")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([local_n], dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(cnt)
    # the accepted network goes from rank 0 to everybody (Coach.py:156-165); every rank then evaluates the same rows identically
    sd = azg.multigpu.broadcast_weights(azg.nnet.random_state_dict(n, seed=1000 + rank), src=0, device=f"cuda:{local}")
    net2 = azg.FusedSplendorNNet(n, state_dict={k: v.cpu() for k, v in sd.items()}, device=local)
    pi, _ = net2(mine["board"][:64].contiguous(), mine["valids"][:64].contiguous()) if local_n >= 64 else (torch.zeros(1, device=f"cuda:{local}"), None)
    w = torch.tensor([float(sum(float(v.double().sum()) for v in sd.values()))], dtype=torch.float64, device=f"cuda:{local}")
    wl, wh = w.clone(), w.clone()
    dist.all_reduce(wl, op=dist.ReduceOp.MIN); dist.all_reduce(wh, op=dist.ReduceOp.MAX)
    return {"examples_local_rank0": local_n, "examples_gathered": total, "sum_of_local_counts": int(cnt.item()),
            "identical_on_all_ranks": bool(torch.equal(lo, hi)), "weights_broadcast_identical": bool(torch.equal(wl, wh)), "backend": "nccl"}


# ----------------------------------------------------------------------------------------------
# the other BASELINE.json configs, each as a nested block of the JSON line
# ----------------------------------------------------------------------------------------------
def _timed_async(eng, torch, dist, world, dev, barrier, G, ticks, steps, warm=2):
    """`steps` x `ticks` ticks of G waves of an asynchronous self-play engine -> (simulations of all ranks, ms max over ranks)"""
    def sims_now():
        return int(eng.sims_completed.item()) + int(eng.sims_in_flight().item())
    for _ in range(warm * ticks):
        eng.tick(G)
    barrier()
    s0 = sims_now()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps * ticks):
        eng.tick(G)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    done = torch.tensor([sims_now() - s0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(done, op=dist.ReduceOp.SUM)
    return int(done.item()), float(t.item())


def bench_config2_genbu(args, torch, dist, azg, world, rank, local, dev, barrier):
    """SURVEY config 2 as the reference runs it (main.py:111-117 defaults): the shipped checkpoint genbu.pt as the network, 4096 lanes,
    1600 sims for a full search, playout-cap randomisation (prob_fullMCTS 0.25, ratio 5), Dirichlet noise alpha 0.2 after a root
    softmax with temperature 1.25, forced playouts + policy-target pruning; exact cleaning"""
    path = os.path.join(ROOT, "tests", "golden", "genbu_n2.npz")
    if args.players != 2 or not os.path.isfile(path):
        return {"skipped": "genbu.pt is a 2-player network"}
    T, sims, G = 4096, args.sims, args.graph_waves or 16
    net = azg.FusedSplendorNNet(2, state_dict=azg.nnet.state_dict_from_npz(path), device=local)
    eng = azg.SelfPlayEngine(2, T, net, sims, device=local, seed=args.seed, game_base=rank * T, cpuct=1.0, fpu=0.0, prob_full=0.25, ratio_full=5,
                             forced_playouts=True, dirichlet_noise=True, dirichlet_alpha=0.2, temperature0=1.25, node_cap=20 * sims,
                             pool_nodes=4 * sims, graph_waves=args.graph_waves, max_levels=args.max_levels, tick_graph=True)
    eng.env.rollout(args.opening_plies, rotate=True)
    eng.start_async()
    ticks = -(-sims // G)
    done, ms = _timed_async(eng, torch, dist, world, dev, barrier, G, ticks, steps=3)
    st = eng.arena.root_stats(want_arrays=False)
    return {"value": done / (ms * 1e-3), "unit": UNIT_MCTS, "workload": "SURVEY config 2: genbu.pt weights, 4096 lanes per GPU, 1600 / 320 sims per move "
            "(prob_fullMCTS 0.25, ratio 5), dirichletAlpha 0.2, temperature[0] 1.25, forced playouts, cpuct 1.0, fpu 0", "trees_per_gpu": T,
            "moves_completed": int(eng.moves_completed.item()), "games_finished": int(eng.games_finished.item()), "ms": ms,
            "truncated_searches": int(st["truncated"].sum()), "lossy_resets": int(st["resets"].sum())}


def bench_virtual_loss(args, torch, dist, azg, world, rank, local, dev, barrier):
    """north star (3), an explicit NON-parity option: leaves_per_tree = 4 simulations of a tree in flight per wave (virtual-loss
    leaf batching) - the batch of the evaluator comes from 4096 trees x 4 leaves instead of 16,384 trees, at a quarter of the memory"""
    n, T, sims, K = args.players, 4096, args.sims, 4
    G = args.graph_waves or 16
    net = azg.FusedSplendorNNet(n, seed=args.seed, device=local)
    eng = azg.SelfPlayEngine(n, T, net, sims, device=local, seed=args.seed, game_base=rank * T, node_cap=20 * sims, pool_nodes=int(5.5 * sims),
                             graph_waves=args.graph_waves, max_levels=args.max_levels, tick_graph=True, leaves_per_tree=K)
    eng.env.rollout(args.opening_plies, rotate=True)
    eng.start_async()
    done, ms = _timed_async(eng, torch, dist, world, dev, barrier, G, max(1, sims // (2 * G)), steps=4)
    st = eng.arena.root_stats(want_arrays=False)
    return {"value": done / (ms * 1e-3), "unit": UNIT_MCTS, "trees_per_gpu": T, "leaves_per_tree": K, "sims_per_move": sims, "ms": ms,
            "arena_gb_per_gpu": eng.arena.arena_bytes / 1e9, "truncated_searches": int(st["truncated"].sum()), "lossy_resets": int(st["resets"].sum()),
            "note": "not the reference's sequential search: visit counts differ from MCTS.py (tests/test_gpu_mcts.py::test_virtual_loss_leaf_batching "
                    "measures how far); the headline value is the parity mode, one leaf per tree"}


def bench_fp32_network(args, torch, dist, azg, world, rank, local, dev, barrier):
    """the same search with the float32 evaluator (torch kernels, tf32 off): what the bf16 tensor-core kernel buys"""
    n, T, sims, G = args.players, 4096, args.sims, 16
    net = azg.SplendorNNetB200(n, seed=args.seed, device=local, dtype=torch.float32)
    eng = azg.SelfPlayEngine(n, T, net, sims, device=local, seed=args.seed, game_base=rank * T, node_cap=20 * sims, pool_nodes=4 * sims,
                             graph_waves=0, max_levels=args.max_levels, tick_graph=False)
    eng.env.rollout(args.opening_plies, rotate=True)
    eng.start_async()
    done, ms = _timed_async(eng, torch, dist, world, dev, barrier, G, 8, steps=2, warm=1)
    return {"value": done / (ms * 1e-3), "unit": UNIT_MCTS, "trees_per_gpu": T, "network": "SplendorNNetB200 float32 (torch, tf32 off), plain launches", "ms": ms}


def bench_config3_arena(args, torch, dist, azg, world, rank, local, dev, barrier):
    """BASELINE configs[3]: 3-player pit of two random-init networks (seeds 1, 2; seats [A,B,B] / [B,A,A], 1-2-2-1 order), playout-cap
    randomisation on (prob_fullMCTS 0.25, ratio 5), every game a lane of BatchedArena; games/s and sims/s of whole games"""
    n, T, sims = 3, 2048, 200
    nets = [azg.FusedSplendorNNet(n, seed=1, device=local), azg.FusedSplendorNNet(n, seed=2, device=local)]
    pit = azg.BatchedArena(n, nets, num_sims=sims, device=local, seed=args.seed, game_base=rank * T, prob_full=0.25, ratio_full=5, node_cap=16 * sims, pool_nodes=6 * sims)
    barrier()
    t0 = time.perf_counter()
    one, two, draws, d = pit.play_games(T)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    tot = torch.tensor([d["total_sims"], T - d["unfinished"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dt = float(dt.item())
    return {"games_per_s": int(tot[1]) / dt, "sims_per_s": int(tot[0]) / dt, "games_per_gpu": T, "sims_per_full_move": sims, "seconds": dt,
            "one_two_draws_rank0": [one, two, draws], "plies": d["plies"], "unfinished_rank0": d["unfinished"],
            "workload": "configs[3]: 3p Arena pit of two random-init SplendorNNet (fused bf16), playout cap 0.25 / ratio 5, lock-step moves, Philox reveals"}


def bench_config2_env4p(args, torch, dist, azg, world, rank, local, dev, barrier):
    """BASELINE configs[2]: 4 players, 16,384 lanes, Philox reveals keyed (seed 1234, game, ply), canonical rotation after every ply"""
    import copy
    a = copy.copy(args)
    a.players, a.lanes, a.seed, a.no_sweep, a.e2e_lanes, a.env_steps = 4, 16384, 1234, True, 16384, 20
    out = bench_env(a, torch, dist, azg, world, rank, local, dev, barrier)
    out["config"]["workload"] = "configs[2]: 4p, 16,384 game lanes per GPU, Philox reveals (seed 1234), canonical rotation every ply, random legal moves"
    out["note"] = "16,384 lanes are 512 lane tiles = 3.5 per SM: this point is launch- and latency-bound, configs[4] (1 Mi lanes) is the throughput point"
    return out


def bench_mcts_wide(args, torch, dist, azg, world, rank, local, dev, barrier):
    """the "64k games" point of the BASELINE metric: many more trees than configs[1], a short budget per move so that the
    arena fits (tree pools scale with the budget), same kernels, same network"""
    n, T, sims = args.players, args.wide_trees, args.wide_sims
    net = azg.FusedSplendorNNet(n, seed=args.seed, device=local)
    cap = 32 * sims
    G = min(args.wide_tick_waves, args.graph_waves) if args.graph_waves > 0 else args.wide_tick_waves      # short budgets: look for finished lanes often
    reach = args.gc == "reachable"
    eng = azg.SelfPlayEngine(n, T, net, sims, device=local, seed=args.seed, game_base=rank * T, node_cap=cap, pool_nodes=6 * sims,
                             gc_reachable=reach, graph_waves=G if args.graph_waves > 0 else 0, rounds=args.rounds,
                             max_levels=args.max_levels, clean_every=max(1, int(args.clean_moves * sims / G)) if (args.async_moves and reach) else 0, clean_percent=45,
                             tick_graph=True)
    eng.env.rollout(args.opening_plies, rotate=True)
    ticks = -(-sims // G)
    if args.async_moves:
        eng.start_async()

    def one_step():
        if args.async_moves:
            for _ in range(ticks):
                eng.tick(G)
        else:
            eng.play_move()

    def sims_now():
        return int(eng.sims_completed.item()) + int(eng.sims_in_flight().item()) if args.async_moves else int(eng.sims_total.item())
    for _ in range(3):
        one_step()
    barrier()
    s0 = sims_now()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    barrier()
    ev0.record()
    for _ in range(steps):
        one_step()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    done = sims_now() - s0
    st = eng.arena.root_stats(want_arrays=False)
    return {"value": world * done / (ms * 1e-3), "unit": UNIT_MCTS, "trees_per_gpu": T, "sims_per_move": sims, "steps": steps, "ms_per_step": ms / steps,
            "arena_gb_per_gpu": eng.arena.arena_bytes / 1e9, "truncated_searches": int(st["truncated"].sum()), "lossy_resets": int(st["resets"].sum())}


# ----------------------------------------------------------------------------------------------
# env leg (configs[4])
# ----------------------------------------------------------------------------------------------
def bench_env(args, torch, dist, azg, world, rank, local, dev, barrier):
    n, L, P = args.players, args.lanes, (args.plies if args.mode == "rollout" else 1)
    steps = args.env_steps
    env = azg.SplendorEnv(n, L, device=local, seed=args.seed, game_base=rank * L, use_tma=not args.no_tma)
    env.reset()
    env.rollout(args.burn_in, rotate=True)
    if args.mode == "step":
        env.step(None, want_next=True, want_ended=False, want_status=False)

    def one_step():
        if args.mode == "rollout":
            env.rollout(P, rotate=True)
        else:
            env.step(env.next_actions, player=0, chance="philox", rotate=True, auto_reset=True, want_next=True,
                     want_status=False, count=True)

    for _ in range(3):
        one_step()
    barrier()
    env.counters.zero_()
    launches0 = env.launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with Clocks(local) as clk:
        barrier()
        evs[0].record()
        for i in range(steps):
            one_step()
            evs[i + 1].record()
        barrier()
    ms_total = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cnt = env.counters.cpu().tolist()
    assert cnt[1] == L * P * steps, f"ply accounting {cnt[1]} != {L * P * steps}"
    value = world * L * P * steps / (ms_total * 1e-3)
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = B_STEP[n] * L * P / (avg_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC_ENV, "value": value, "unit": UNIT_ENV, "steps": steps, "ms_per_step": ms_total / steps, "dtype": "int8",
        "config": {"workload": workload_env(args), "lanes_per_gpu": L, "plies_per_launch": P, "mode": args.mode, "burn_in_plies": args.burn_in,
                   "tma": not args.no_tma, "l2": f"inputs larger than L2: {L * env.S / 1e6:.0f} MB of lane tiles per GPU vs 126 MB L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (lambda tr: None if tr is None else tr * L * P)(load_traffic(f"{args.mode}_n{n}_bytes_per_lane_ply")),
                     "kernel": "spl_rollout_kernel" if args.mode == "rollout" else "spl_step_kernel",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                     "algorithmic_bytes_per_step": B_STEP[n], "avg_launch_ms": avg_ms},
        "gpu_launches": env.launches - launches0, "games_finished": cnt[0] * world, "clocks": clk.summary(),
    }
    # e2e through the reference-facing calls with HOST buffers (pinned): every call uploads the boards (+ actions) and downloads the results
    game = azg.SplendorGame(n, seed=args.seed, device=local)
    Le = args.e2e_lanes
    src = azg.SplendorEnv(n, Le, device=local, seed=args.seed, game_base=rank * Le)
    src.reset(); src.rollout(20, rotate=True); src.step(None, want_next=True)
    boards0 = src.states().cpu().numpy(); acts0 = src.next_actions.cpu().numpy()
    pipe = game._pipe_for(Le)

    def timed(fn, ke=5):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / ke
    pipe.h_in.numpy()[...] = boards0.reshape(Le, -1); pipe.h_act.numpy()[...] = acts0
    dt1 = timed(lambda: pipe.step(0, False, True))                      # getNextStateBatch(..., packed_masks=True) on its pinned buffers
    Kr = args.plies
    dtk = timed(lambda: pipe.rollout(Kr))                               # rolloutBatch(boards, plies): one round trip per Kr plies
    eenv = game._env_for(Le)
    eenv._h_in.copy_(torch.from_numpy(boards0)); eenv._h_act.copy_(torch.from_numpy(acts0))
    dt0 = timed(lambda: game._step_pinned(eenv, 0, False, True))        # the one-stream call with bool[L,406] masks
    h1, d1 = pipe.bytes_per_call(False)
    hk, dk = pipe.bytes_per_call(True)
    out["e2e"] = {"value": world * Le * Kr / dtk, "unit": UNIT_ENV, "h2d_bytes_per_step": hk, "d2h_bytes_per_step": dk,
                  "call": f"SplendorGame.rolloutBatch(host int8[L,R,7] boards, plies={Kr}): the Arena.playGame loop with random players, {Kr} plies per "
                          f"host round trip (pinned boards up, boards + counters down, {len(pipe.parts)} lane chunks on their own streams)",
                  "lanes_per_call": Le, "plies_per_call": Kr,
                  "single_step": {"value": world * Le / dt1, "unit": UNIT_ENV, "h2d_bytes_per_step": h1, "d2h_bytes_per_step": d1,
                                  "call": "SplendorGame.getNextStateBatch(host boards, actions, packed_masks=True): one ply per round trip; next canonical boards, "
                                          "52-byte legal masks and end vectors come back; chunked over streams (upload / kernels / download overlap)"},
                  "single_step_bool_masks": {"value": world * Le / dt0, "unit": UNIT_ENV, "h2d_bytes_per_step": Le * (env.S + 2),
                                             "d2h_bytes_per_step": Le * (env.S + 406 + 4 * n),
                                             "call": "SplendorGame.getNextStateBatch(host boards, actions): the reference's bool[L,406] masks, one stream"}}
    del env
    if rank == 0 and not args.no_sweep:
        out["sweep"] = sweep(azg, torch, n, local, args)
    return out


def sweep(azg, torch, n, local, args):
    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    for Ls in (1 << 10, 1 << 13, 1 << 16, 1 << 18):
        e = azg.SplendorEnv(n, Ls, device=local, seed=args.seed, use_tma=not args.no_tma)
        e.reset()
        e.rollout(args.burn_in, rotate=True)
        for _ in range(3):
            e.rollout(args.plies, rotate=True)
        tot, reps = 0.0, 5
        for _ in range(reps):
            flush.fill_(1)   # evict L2 between timed launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e.rollout(args.plies, rotate=True); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        out.append({"lanes": Ls, "plies_per_launch": args.plies, "steps_per_s": Ls * args.plies * reps / (tot * 1e-3), "l2": "flushed"})
        del e
    return out


# ----------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.cpu_mcts_worker:
        return cpu_mcts_worker(args.cpu_mcts_worker)
    if args.cpu_ref_worker:
        return cpu_ref_worker(args.cpu_ref_worker)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import azg_b200 as azg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator comes up; stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.players
    line = None
    if args.workload in ("both", "mcts"):
        line, eng = bench_mcts(args, torch, dist, azg, world, rank, local, dev, barrier)
        del eng
        torch.cuda.empty_cache()
        if args.wide_trees > 0:
            line["mcts_wide"] = bench_mcts_wide(args, torch, dist, azg, world, rank, local, dev, barrier)
            torch.cuda.empty_cache()
        if not args.no_extra:
            for key, fn in (("config2_genbu", bench_config2_genbu), ("virtual_loss_4_leaves", bench_virtual_loss), ("fp32_network", bench_fp32_network),
                            ("config3_arena_3p", bench_config3_arena)):
                line[key] = fn(args, torch, dist, azg, world, rank, local, dev, barrier)
                torch.cuda.empty_cache()
    if args.workload in ("both", "env"):
        envres = bench_env(args, torch, dist, azg, world, rank, local, dev, barrier)
        if line is None:
            line = {"metric": METRIC_ENV, "value": envres["value"], "unit": UNIT_ENV, "n_gpus": world, "steps": envres["steps"],
                    "warmup": 3, "ms_per_step": envres["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "int8", "data": "synthetic", "config": dict(envres["config"], players=n,
                                                                          parallelism=f"games sharded dp{world}, no collective on the path"),
                    "roofline": envres["roofline"], "gpu_launches": envres["gpu_launches"], "e2e": envres["e2e"],
                    "sweep": envres.get("sweep"), "clocks": envres["clocks"]}
        else:
            line["env_steps"] = envres
        if not args.no_extra:
            line["config2_env_4p"] = bench_config2_env4p(args, torch, dist, azg, world, rank, local, dev, barrier)
    if rank == 0:
        if not args.no_cpu:
            if line["metric"] == METRIC_MCTS:
                sims_cpu = args.sims
                cpu_moves = max(3, int(args.cpu_seconds * 4000 / sims_cpu))     # ~cpu_seconds of work per core at ~4k sims/s
                rate, cores, tot, secs = cpu_mcts(n, sims_cpu, cpu_moves, args.seed, args.fixed_net)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT_MCTS, "cores": cores, "kind": "port",
                                        "sample": f"{cores} processes x {cpu_moves} moves x {sims_cpu} sims/move = {tot} simulations in {secs:.1f} s (C port of "
                                                  f"MCTS.py + per-leaf torch-CPU float32 SplendorNNet predict, one thread per core)"}
                if reference_dir() is not None and not args.no_extra:
                    ref = cpu_mcts(n, sims_cpu, max(2, cpu_moves // 2), args.seed, args.fixed_net, worker="--cpu-ref-worker")
                    if ref is not None:
                        line["cpu_baseline_reference"] = {"value": ref[0], "unit": UNIT_MCTS, "cores": ref[1], "kind": "reference",
                                                          "sample": f"{ref[1]} processes x {max(2, cpu_moves // 2)} moves x {sims_cpu} sims/move = {ref[2]} simulations in "
                                                                    f"{ref[3]:.1f} s (the reference's own MCTS.py + Numba Board + NNetWrapper.predict, one thread per core, "
                                                                    f"JIT and one warm-up move excluded)"}
            if args.workload in ("both", "env"):
                v, cores, games, plies, dtc = cpu_rollouts(n, args.seed, args.cpu_seconds)
                cb = {"value": v, "unit": UNIT_ENV, "cores": cores, "kind": "port",
                      "sample": f"{games} whole random {n}p games = {plies} plies in {dtc:.1f} s (oracle port of the SplendorLogicNumba "
                                f"rules, C -O2, one thread per core)"}
                if "env_steps" in line:
                    line["env_steps"]["cpu_baseline"] = cb
                else:
                    line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
