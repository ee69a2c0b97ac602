"""where a tick's time goes in the asynchronous self-play loop: the captured waves, the per-tick move logic (sample -> env step ->
resets -> budgets -> begin, replayed as a graph) and, eagerly with events in between, the stages of that move logic"""
import sys, torch
sys.path.insert(0, "/root/repo")
import azg_b200 as azg
n = 2
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
G = int(sys.argv[3]) if len(sys.argv) > 3 else 128
net = azg.FusedSplendorNNet(n, seed=1)
eng = azg.SelfPlayEngine(n, T, net, sims, seed=1, node_cap=20 * sims, pool_nodes=4 * sims, graph_waves=G, max_levels=16, clean_every=0, tick_graph=True)
eng.env.rollout(24, rotate=True)
eng.start_async()
ticks_per_move = max(1, sims // G)
for mv in range(6):
    for _ in range(ticks_per_move):
        eng.tick(G)
st = eng.arena.root_stats(want_arrays=False)
print("nodes mean/max", float(st["nodes"].float().mean()), int(st["nodes"].max()), "pool", eng.arena.pool_stats())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tw = tt = 0.0
K = 40
for _ in range(K):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); eng._run_waves(G); e1.record(); eng._tick_graph.replay(); e2.record()
    torch.cuda.synchronize()
    tw += e0.elapsed_time(e1); tt += e1.elapsed_time(e2)
print("steady state per tick: waves %.3f ms (%.1f us per wave), move logic %.3f ms" % (tw / K, 1e3 * tw / K / G, tt / K))
# the stages of the move logic, eagerly
names = ["sample_moves", "env.step", "episodes+env.reset", "arena.reset", "budgets", "env.states", "arena.begin"]
acc = [0.0] * len(names)
for _ in range(K):
    eng._run_waves(G)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    eng.arena.sample_moves(1.0, eng.env.episodes, eng.actions, eng._fin8, eng._move_counters)
    fin = eng._fin8.to(torch.bool)
    ev[1].record()
    eng.env.step(eng.actions, player=0, chance="philox", rotate=True, auto_reset=False, want_mask=False, want_status=False, count=True)
    done = fin & (eng.env.ended != 0).any(dim=1)
    done8 = done.to(torch.uint8)
    ev[2].record()
    eng.env.episodes += done.to(torch.int32)
    eng.env.reset(done8)
    ev[3].record()
    eng.arena.reset(done8)
    ev[4].record()
    eng.games_finished += done.sum()
    eng._assign_budgets(fin)
    ev[5].record()
    eng.env.states(out=eng.roots)
    ev[6].record()
    eng.arena.begin(eng.roots, eng.sims, eng.flags, eng._fin8)
    ev[7].record()
    torch.cuda.synchronize()
    for i in range(len(names)):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
print("move logic, eager, ms per tick:", {k: round(v / K, 3) for k, v in zip(names, acc)})
