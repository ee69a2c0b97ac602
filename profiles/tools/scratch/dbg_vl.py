import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import azg_b200 as az
n, T, sims = 2, 64, 200
env = az.SplendorEnv(n, T, seed=17); env.reset(); env.rollout(24, rotate=True)
roots = env.states()
simt = torch.full((T,), sims, dtype=torch.int32, device=roots.device)
for K in (1, 2, 4):
    ar = az.MCTSArena(n, T, node_cap=1024, cpuct=1.5, fpu=0.2, leaves_per_tree=K)
    ar.begin(roots, simt)
    hist = []
    for w in range(400):
        ar.counters.zero_()
        ar.select(count=True)
        c = ar.counters.cpu().tolist()
        if c[1] == 0: break
        hist.append(c[0])
        pi, v = ar.fixed_net(); ar.expand(pi, v)
    st = ar.root_stats(want_arrays=False)
    print("K", K, "waves", len(hist), "leaves/wave first 12", hist[:12], "mean", sum(hist)/len(hist), "sims_done min/max", int(st["sims_done"].min()), int(st["sims_done"].max()), "nn_calls", int(st["nn_calls"].sum()))
