import sys, os, numpy as np, torch
sys.path.insert(0,'/root/repo')
import azg_b200 as az
sd = az.nnet.state_dict_from_npz('/root/repo/tests/golden/genbu_n2.npz')
n, T, sims = 2, 256, 48
net = az.FusedSplendorNNet(n, state_dict=sd)
eng = az.SelfPlayEngine(n, T, net, sims, seed=3, cpuct=1.0, fpu=0.0, prob_full=0.25, ratio_full=5, forced_playouts=True,
                        dirichlet_noise=True, dirichlet_alpha=0.2, temperature0=1.25, node_cap=4096)
eng.env.rollout(20, rotate=True)
probs, q, is_full = eng.search(1.0)
torch.cuda.synchronize()
bad = ~torch.isfinite(probs).all(1) | (probs.sum(1) <= 0) | (probs < 0).any(1)
print("bad lanes", int(bad.sum()), "full among bad", is_full[bad].tolist()[:10])
st = eng.arena.root_stats()
i = int(torch.nonzero(bad)[0]) if bad.any() else 0
print("lane", i, "sims", int(eng.sims[i]), "done", int(st["sims_done"][i]), "ns", int(st["ns"][i]), "status", int(st["status"][i]))
print("nsa", st["nsa"][i][st["nsa"][i] > 0].tolist())
print("ps finite", bool(torch.isfinite(st["ps"][i]).all()), "ps sum", float(st["ps"][i].sum()), "qs", float(st["qs"][i]))
print("probs", probs[i][probs[i] != 0].tolist()[:10])
pi, v = net(eng.roots, eng.env.valids()) if False else (None, None)
