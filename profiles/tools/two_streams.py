#!/usr/bin/env python3
"""Does it pay to run the games of one GPU as SEVERAL independent groups on their own streams (kernels of one group fill the tails and the
idle SMs of the other's)?   python profiles/tools/two_streams.py --trees 18944 --groups 1 2 4"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200 as azg

ap = argparse.ArgumentParser()
ap.add_argument("--trees", type=int, default=18944)
ap.add_argument("--sims", type=int, default=1600)
ap.add_argument("--groups", type=int, nargs="+", default=[1, 2])
ap.add_argument("--graph-waves", type=int, default=128)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
n, dev = 2, 0
for ng in a.groups:
    T = a.trees // ng
    engs, streams = [], []
    for g in range(ng):
        net = azg.FusedSplendorNNet(n, seed=20261018, device=dev)
        e = azg.SelfPlayEngine(n, T, net, a.sims, device=dev, seed=20261018, game_base=g * T, cpuct=1.0, fpu=0.0, node_cap=20 * a.sims,
                               pool_nodes=int(4.8 * a.sims), graph_waves=a.graph_waves, max_levels=16, tick_graph=True)
        e.env.rollout(24, rotate=True)
        s = torch.cuda.Stream(dev)
        with torch.cuda.stream(s):
            e.start_async()
        engs.append(e); streams.append(s)
    G = a.graph_waves
    ticks = -(-a.sims // G)

    def step():
        for _ in range(ticks):
            for e, s in zip(engs, streams):
                with torch.cuda.stream(s):
                    e.tick(G)

    def sims_now():
        torch.cuda.synchronize()
        return sum(int(e.sims_completed.item()) + int(e.sims_in_flight().item()) for e in engs)
    for _ in range(3):
        step()
    s0 = sims_now()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    s1 = sims_now()
    dt = time.perf_counter() - t0
    print(f"groups {ng} x {T} trees: {(s1 - s0) / dt:.3e} sims/s ({dt / a.steps * 1e3:.1f} ms per {a.sims} waves)", flush=True)
    del engs
    torch.cuda.empty_cache()
