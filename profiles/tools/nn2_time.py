"""Fused evaluator (csrc/spl_nnet2.cu), 2 players: phase stamps of CTA 0 and back-to-back launch times by batch size; a quick accuracy
check against the float32 folded pass. SPL_B200_LIB selects a variant build."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200
from azg_b200 import nnet as nn

dev = torch.device("cuda", 0)
lib = azg_b200._native.lib(); lib.spl_nnet_debug_stamps.argtypes = [C.c_void_p]
n = 2
sd = nn.random_state_dict(n, seed=5)
net = azg_b200.FusedSplendorNNet(n, state_dict=sd)
W = nn.fold(sd, dev)
env = azg_b200.SplendorEnv(n, 4096, seed=3); env.reset(); env.rollout(30, rotate=True)
st = env.states(); env.step(None, store_state=False, want_ended=False, want_status=False); va = env.valids()
pi, val = net(st, va); torch.cuda.synchronize()
rp, rv = nn.forward_folded(W, st, va)
print(f"lib {os.environ.get('SPL_B200_LIB', 'base')}: max|dp| {float((pi - rp).abs().max()):.4f} max|dv| {float((val - rv).abs().max()):.4f} "
      f"top1 {float((pi.argmax(1) == rp.argmax(1)).float().mean()):.3f}", flush=True)
buf = (C.c_longlong * 32)(); lib.spl_nnet_debug_stamps(buf)
names = ["input", "L1", "L2", "G1", "L3+flatten", "L4", "G4", "L5a", "L5b", "G5", "head", "softmax"]
print("  error flag", buf[31], "phases us:", {names[i]: round((buf[i + 1] - buf[i]) / 1.965e3, 2) for i in range(12) if buf[i + 1] and buf[i]},
      "total", round((buf[12] - buf[0]) / 1.965e3, 2), flush=True)
for B in (4096, 9472, 16384, 18944, 65536):
    envb = azg_b200.SplendorEnv(n, B, seed=4); envb.reset(); envb.rollout(30, rotate=True)
    sb = envb.states(); envb.step(None, store_state=False, want_ended=False, want_status=False); vb = envb.valids()
    for _ in range(5): net(sb, vb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(40): net(sb, vb)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 40
    print(f"  B={B}: {us:.1f} us per launch, {B / us:.1f} leaves/us, {B * 1.24e6 / us / 1e6:.1f} TFLOP/s", flush=True)
