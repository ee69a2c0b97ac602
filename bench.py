#!/usr/bin/env python3
"""bench.py - Splendor env steps/s on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--lanes L] [--plies P] [--mode rollout|step]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: oracle port of the reference rules on all host cores

A "step" of the bench is one pass of the hot path over the whole batch of game lanes: one launch of the
persistent ply kernel (`--mode rollout`, P plies per lane per launch: legality mask -> uniform random legal
action -> move + Philox deck reveal -> ply++ -> canonical rotation -> end-game check -> auto reset), or of the
single-ply kernel (`--mode step`, P = 1, state read from and written back to HBM every ply).
value = env steps (lane-plies) per second over all ranks. Games shard over GPUs with no collective on the path
(weak scaling: lanes per GPU fixed).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "splendor_env_steps_per_s"
UNIT = "steps/s"
B_STEP = {2: 846, 3: 1060, 4: 1302}   # algorithmic bytes per env step: 2S + 52 + 2 + 4n


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--players", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=1 << 20, help="game lanes per GPU")
    ap.add_argument("--plies", type=int, default=16, help="plies per lane per launch (rollout mode)")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "step"])
    ap.add_argument("--e2e-lanes", type=int, default=65536)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--burn-in", type=int, default=300, help="untimed plies per lane before warm-up (de-synchronises the games)")
    ap.add_argument("--seed", type=int, default=20261018)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference rules (oracle/), all host cores
# ----------------------------------------------------------------------------------------------
def cpu_rollouts(n_players, seed, seconds, threads=None):
    """plays whole random games (same Philox policy as the GPU path) on `threads` host threads for about
    `seconds`; returns (steps/s, threads, games, plies)"""
    from oracle import pyoracle as po
    po.lib()
    threads = threads or len(os.sched_getaffinity(0))
    # calibrate one thread
    t0 = time.perf_counter()
    tot, _, _ = po.rollout(n_players, seed, 0, 200)
    dt = time.perf_counter() - t0
    per_game = dt / 200
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.players
    per_step = max(0.5, min(10.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rollouts(n, args.seed, min(per_step, 1.0))
    tot_plies, tot_dt, cores, games = 0, 0.0, 0, 0
    for _ in range(args.steps):
        v, cores, g, plies, dt = cpu_rollouts(n, args.seed, per_step)
        tot_plies += plies; tot_dt += dt; games += g
    value = tot_plies / tot_dt
    sample = f"{games} whole random {n}p games (oracle port of SplendorLogicNumba rules, C, -O2), {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": workload_name(args), "players": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"config5: {args.players}p env-step throughput, random legal moves, {args.lanes} game lanes per GPU "
            f"(64k-lane point of the BASELINE metric reported in `sweep`)")


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


# ----------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import azg_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, L, P = args.players, args.lanes, (args.plies if args.mode == "rollout" else 1)

    # games shard over ranks by global game id: lane l of rank r is game r*L + l (results independent of N)
    env = azg_b200.SplendorEnv(n, L, device=local, seed=args.seed, game_base=rank * L, use_tma=not args.no_tma)
    env.reset()
    # burn-in: games start in lock-step; play until the lanes are spread over all game phases (steady state)
    env.rollout(args.burn_in, rotate=True)
    if args.mode == "step":
        env.step(None, want_next=True, want_ended=False, want_status=False)

    def one_step():
        if args.mode == "rollout":
            env.rollout(P, rotate=True)
        else:
            env.step(env.next_actions, player=0, chance="philox", rotate=True, auto_reset=True, want_next=True,
                     want_status=False, count=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    env.counters.zero_()
    launches0 = env.launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with Clocks(local) as clk:
        barrier()
        t0 = time.perf_counter()
        evs[0].record()
        for i in range(args.steps):
            one_step()
            evs[i + 1].record()
        barrier()
        wall = time.perf_counter() - t0
    ms_total = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    launches = env.launches - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cnt = env.counters.cpu().tolist()
    assert cnt[1] == L * P * args.steps, f"ply accounting {cnt[1]} != {L * P * args.steps}"
    steps_all = world * L * P * args.steps
    value = steps_all / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (the ply kernel itself; events bracket exactly one launch each)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = B_STEP[n] * L * P / (avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = prof.get(f"{args.mode}_n{n}_bytes_per_lane_ply")
        if traffic is not None:
            traffic = traffic * L * P
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "spl_rollout_kernel" if args.mode == "rollout" else "spl_step_kernel",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_step": B_STEP[n], "avg_launch_ms": avg_ms}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": workload_name(args), "players": n, "lanes_per_gpu": L, "plies_per_launch": P, "mode": args.mode, "burn_in_plies": args.burn_in,
                   "tma": not args.no_tma, "parallelism": f"games sharded dp{world}, no collective on the path",
                   "l2": f"inputs larger than L2: {L * env.S / 1e6:.0f} MB of lane tiles per GPU vs 126 MB L2"},
        "roofline": roofline, "gpu_launches": launches, "wall_s": wall, "games_finished": cnt[0] * world,
    }
    if rank == 0:
        line["clocks"] = clk.summary()

    # ---- e2e through the reference-facing call with HOST buffers (rank-local, all ranks in parallel)
    game = azg_b200.SplendorGame(n, seed=args.seed, device=local)
    Le = args.e2e_lanes
    eenv = game._env_for(Le)
    src = azg_b200.SplendorEnv(n, Le, device=local, seed=args.seed, game_base=rank * Le)
    src.reset(); src.rollout(20, rotate=True); src.step(None, want_next=True)
    eenv._h_in.copy_(src.states().cpu()); eenv._h_act.copy_(src.next_actions.cpu())
    for _ in range(3):
        game._step_pinned(eenv, 0, False, True)
    barrier()
    ke = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(ke):
        game._step_pinned(eenv, 0, False, True)     # includes H2D of boards+actions and D2H of boards+masks+end vectors
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line["e2e"] = {"value": world * Le * ke / float(t.item()), "unit": UNIT,
                   "h2d_bytes_per_step": Le * (env.S + 2), "d2h_bytes_per_step": Le * (env.S + 406 + 4 * n),
                   "call": "SplendorGame.getNextStateBatch (pinned host int8[L,R,7] boards + actions in; next canonical boards, "
                           "bool[L,406] masks, float32[L,n] end vectors out)", "lanes_per_call": Le}

    # ---- sweep point named by the metric: 64k lanes, L2 flushed between timed launches
    if rank == 0:
        if not args.no_sweep:
            line["sweep"] = sweep(azg_b200, torch, n, local, args)
        if not args.no_cpu:
            v, cores, games, plies, dtc = cpu_rollouts(n, args.seed, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{games} whole random {n}p games = {plies} plies in {dtc:.1f} s "
                                              f"(oracle port of the SplendorLogicNumba rules, C -O2, one thread per core)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sweep(azg_b200, torch, n, local, args):
    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    for Ls in (1 << 10, 1 << 13, 1 << 16, 1 << 18):
        e = azg_b200.SplendorEnv(n, Ls, device=local, seed=args.seed, use_tma=not args.no_tma)
        e.reset()
        e.rollout(args.burn_in, rotate=True)
        for _ in range(3):
            e.rollout(args.plies, rotate=True)
        tot = 0.0
        reps = 5
        for _ in range(reps):
            flush.fill_(1)   # evict L2 between timed launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e.rollout(args.plies, rotate=True); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        out.append({"lanes": Ls, "plies_per_launch": args.plies, "steps_per_s": Ls * args.plies * reps / (tot * 1e-3), "l2": "flushed"})
        del e
    return out


if __name__ == "__main__":
    main()
