import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, __import__("os").path.realpath(__import__("os").path.join(__import__("os").path.dirname(__file__), "..", "..")))
import azg_b200
net = azg_b200.FusedSplendorNNet(2, seed=1)
env = azg_b200.SplendorEnv(2, 4096, seed=3); env.reset(); env.rollout(30, rotate=True)
st = env.states(); env.step(None, store_state=False, want_ended=False, want_status=False); va = env.valids()
for _ in range(5): net(st, va)
torch.cuda.synchronize()
buf = (C.c_longlong * 32)()
lib = azg_b200._native.lib(); lib.spl_nnet_debug_stamps.argtypes=[C.c_void_p]
lib.spl_nnet_debug_stamps(buf)
s = np.array(buf[:12], dtype=np.int64)
names = ["start->input", "input+L1", "L2", "G1", "L3", "flatten", "L4", "G4..V0", "PI1", "V1", "softmax"]
d = np.diff(s) / 1.965e3
s2 = np.array(buf[:16], dtype=np.int64)
print("L5a step: acquire->gemm+epilogue", round((s2[13]-s2[12])/1.965e3, 3), "release", round((s2[14]-s2[13])/1.965e3, 3), "step start (after G4) -> acquire done", "n/a")
s3 = np.array(buf[:32], dtype=np.int64)
print("L2 (tcgen05 layer) us: weights+barrier %.3f, MMA issue %.3f, wait for the accumulator %.3f, epilogue %.3f, rest %.3f" % (tuple(
    (s3[i + 1] - s3[i]) / 1.965e3 for i in (16, 17, 18, 19)) + ((s3[3] - s3[20]) / 1.965e3,)))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
evs[0].record()
for i in range(40):
    net(st, va); evs[i + 1].record()
torch.cuda.synchronize()
print("kernel, back-to-back launches, us:", round(1e3 * evs[0].elapsed_time(evs[40]) / 40, 2))
print("us per phase:", dict(zip(names, np.round(d, 2))), "total", round((s[11]-s[0])/1.965e3, 2))

cta = (C.c_longlong * 320)()
lib.spl_nnet_debug_cta_times.argtypes = [C.c_void_p]
lib.spl_nnet_debug_cta_times(cta)
c = np.array(cta[:256], dtype=np.int64).reshape(128, 2)
t0 = c[:, 0].min()
print("CTA start after the first CTA, us: p50 %.2f max %.2f; CTA duration us: min %.2f p50 %.2f max %.2f; last end - first start %.2f" % (
    np.median(c[:, 0] - t0) / 1e3, (c[:, 0] - t0).max() / 1e3, (c[:, 1] - c[:, 0]).min() / 1e3, np.median(c[:, 1] - c[:, 0]) / 1e3,
    (c[:, 1] - c[:, 0]).max() / 1e3, (c[:, 1].max() - t0) / 1e3))
