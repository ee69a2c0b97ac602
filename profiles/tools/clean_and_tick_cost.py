import sys, torch
sys.path.insert(0, "/root/repo")
import azg_b200 as azg
n = 2
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
G = int(sys.argv[3]) if len(sys.argv) > 3 else 128
net = azg.FusedSplendorNNet(n, seed=1)
cap = 8 * sims
eng = azg.SelfPlayEngine(n, T, net, sims, seed=1, node_cap=cap, edge_cap=cap * 36, gc_reachable=True, graph_waves=G, max_levels=16, clean_every=0, tick_graph=True)
eng.env.rollout(24, rotate=True)
eng.start_async()
ticks_per_move = max(1, sims // G)
for mv in range(12):
    for _ in range(ticks_per_move):
        eng.tick(G)
    if mv % 4 == 3:
        st = eng.arena.root_stats(want_arrays=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.arena.clean(45); b.record(); torch.cuda.synchronize()
        st2 = eng.arena.root_stats(want_arrays=False)
        print("move", mv + 1, "clean ms %.2f" % a.elapsed_time(b), "nodes before mean/max", float(st["nodes"].float().mean()), int(st["nodes"].max()),
              "after", float(st2["nodes"].float().mean()), int(st2["nodes"].max()), "cleaned trees", int((st2["cleanings"] - st["cleanings"]).sum()))
# tick tail cost
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    eng._tick_graph.replay()
b.record(); torch.cuda.synchronize()
print("tick tail graph replay ms", a.elapsed_time(b) / 20)

# one graph replay of G waves
a.record()
for _ in range(5):
    eng._graph.replay()
b.record(); torch.cuda.synchronize()
print("wave graph replay ms (%d waves)" % G, a.elapsed_time(b) / 5)
# steady state, tick by tick: the waves and the move logic timed separately
tw = tt = 0.0
K = 40
for _ in range(K):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); eng._run_waves(G); e1.record(); eng._tick_graph.replay(); e2.record()
    torch.cuda.synchronize()
    tw += e0.elapsed_time(e1); tt += e1.elapsed_time(e2)
print("steady state per tick: waves %.3f ms (%.1f us per wave), move logic %.3f ms" % (tw / K, 1e3 * tw / K / G, tt / K))
