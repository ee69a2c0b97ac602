"""Development check of the fused evaluator (csrc/spl_nnet2.cu): outputs against the float32 folded pass on mid-game positions for
2 / 3 / 4 players and ragged batch sizes, phase stamps of CTA 0 and back-to-back launch times. SPL_NNET_IMPL=1 runs the first version."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200
from azg_b200 import nnet as nn

dev = torch.device("cuda", 0)
lib = azg_b200._native.lib(); lib.spl_nnet_debug_stamps.argtypes = [C.c_void_p]
impl = os.environ.get("SPL_NNET_IMPL", "2")
for n in (2, 3, 4):
    sd = nn.random_state_dict(n, seed=5)
    net = azg_b200.FusedSplendorNNet(n, state_dict=sd)
    W = nn.fold(sd, dev)
    env = azg_b200.SplendorEnv(n, 4096, seed=3); env.reset(); env.rollout(30, rotate=True)
    st = env.states(); env.step(None, store_state=False, want_ended=False, want_status=False); va = env.valids()
    for B in (1, 3, 63, 64, 65, 200, 4096):
        s, v_ = st[:B].contiguous(), va[:B].contiguous()
        pi, val = net(s, v_)
        torch.cuda.synchronize()
        rp, rv = nn.forward_folded(W, s, v_)
        dp, dv = (pi - rp).abs(), (val - rv).abs()
        top1 = float((pi.argmax(1) == rp.argmax(1)).float().mean())
        ok = bool(torch.isfinite(pi).all()) and float(dp.max()) < 0.08 and float(dv.max()) < 0.06
        print(f"impl {impl} n={n} B={B}: max|dp| {float(dp.max()):.4f} mean|dp| {float(dp.mean()):.2e} max|dv| {float(dv.max()):.4f} top1 {top1:.3f} "
              f"sum(pi) {float(pi.sum(1).min()):.4f}..{float(pi.sum(1).max()):.4f} {'OK' if ok else 'FAIL'}", flush=True)
    buf = (C.c_longlong * 32)(); lib.spl_nnet_debug_stamps(buf)
    print("  error flag", buf[31], "stamps us:", [round((buf[i + 1] - buf[i]) / 1.965e3, 2) for i in range(12) if buf[i + 1] and buf[i]], "total",
          round((buf[12] - buf[0]) / 1.965e3, 2) if impl == "2" else round((buf[11] - buf[0]) / 1.965e3, 2), flush=True)
    if impl == "2":
        print("  G1 step, us after its first stamp (per half: accumulator ready, TMEM read, stores done, published):",
              [round((buf[i] - buf[16]) / 1.965e3, 2) for i in range(16, 24)], flush=True)
    if impl == "2":
        print("  prologue us: setup %.2f, row flags %.2f, loads issued + legality bits %.2f, converted + published %.2f" % tuple(
            (buf[b] - buf[a]) / 1.965e3 for a, b in ((0, 13), (13, 14), (14, 15), (15, 1))), flush=True)
        tb = (C.c_longlong * 144)(); lib.spl_nnet_debug_tile_stamps.argtypes = [C.c_void_p]; lib.spl_nnet_debug_tile_stamps(tb)
        t = np.array(tb[:], dtype=np.int64).reshape(3, 48)
        nt = int((t[0] != 0).sum())
        print("  tiles: requested / landed / issued, us after kernel start:", flush=True)
        print("   ", " ".join("%d:%.1f/%.1f/%.1f" % (i, (t[0, i] - buf[0]) / 1.965e3, (t[1, i] - buf[0]) / 1.965e3, (t[2, i] - buf[0]) / 1.965e3) for i in range(nt)), flush=True)
    if n == 2:
        for B in (64, 4096, 9472, 16384, 18944, 65536):
            envb = azg_b200.SplendorEnv(n, B, seed=4); envb.reset(); envb.rollout(30, rotate=True)
            sb = envb.states(); envb.step(None, store_state=False, want_ended=False, want_status=False); vb = envb.valids()
            for _ in range(5): net(sb, vb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(40): net(sb, vb)
            e1.record(); torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 40
            print(f"  impl {impl} B={B}: {us:.1f} us per launch, {B / us:.1f} leaves/us, {B * 1.24e6 / us / 1e6:.1f} TFLOP/s", flush=True)
