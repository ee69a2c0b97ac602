import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle/refgen')
import azg_b200 as az
import gen_callers_golden as gcg
from oracle import pyoracle as po
n=3
g=np.load('/root/repo/tests/golden/callers_coach_n3.npz')
A,T=gcg.COACH_ARGS,1
eng = az.SelfPlayEngine(n, T, None, A["numMCTSSims"], seed=1, cpuct=A["cpuct"], fpu=A["fpu"], ratio_full=A["ratio_fullMCTS"],
                        forced_playouts=A["forced_playouts"], dirichlet_noise=True, dirichlet_alpha=A["dirichletAlpha"],
                        temperature0=A["temperature"][0], node_cap=8192, record_examples=True)
eng.evaluator = lambda s, v: eng.arena.fixed_net(s, v)
dev=eng.device
eng.env.set_states(torch.from_numpy(np.repeat(g["init"][None], T, 0)))
# oracle replay
b=po.Board(n); b.set_state(g["init"])
full = g["coins"] < A["prob_fullMCTS"]
k=0
for mv in range(len(g["actions"])):
    is_full = torch.full((T,), bool(full[mv]), dtype=torch.bool, device=dev)
    dirv=None
    if full[mv]:
        dirv = torch.from_numpy(np.repeat(g["dirs"][k][None], T, 0)).to(dev).contiguous(); k+=1
    code=int(g["reveals"][mv])
    st_before = eng.env.states().cpu().numpy()[0]
    if not np.array_equal(st_before, b.state):
        print("state mismatch before move", mv); d=np.argwhere(st_before!=b.state); print(d[:10]); break
    probs,q,isf,ended = eng.play_move(1.0, is_full=is_full, dir_values=dirv, forced_actions=torch.full((T,), int(g["actions"][mv]), dtype=torch.int16),
                  reveals=torch.full((T,), 255 if code < 0 else code, dtype=torch.uint8))
    nxt=b.make_move(int(g["actions"][mv]),0,code if code>=0 else -1); b.swap_players(nxt)
    e=b.check_end_game()
    if mv>len(g["actions"])-4: print(mv, "engine ended", ended.cpu().numpy(), "oracle", e, "finished", int(eng.games_finished))
