"""Cycles per tcgen05.mma (M = 128, K = 16, bf16) by N and by the layout of the B operand (csrc/spl_umma.cu: spl_umma_mma_cycles)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200
from azg_b200 import _native as nat
lib = nat.lib(); dev = torch.device("cuda", 0)
h = C.c_void_p(); nat.check(lib.spl_ctx_create(2, 10, nat.RULES_DEFAULT, 0, C.byref(h)))
out = torch.zeros(2, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for n in (32, 64, 128, 224, 256):
    for b_mn, sbo, name in ((0, 0, "B K-major"), (1, 2048, "B MN-major, SBO 2048"), (1, 256, "B MN-major, SBO 256"), (1, 128, "B MN-major, SBO 128 (n contiguous), LBO 128")):
        for ks in (8,):
            reps = 16
            for _ in range(2):
                nat.check(lib.spl_umma_mma_cycles(h, n, ks, b_mn, sbo, reps, C.c_void_p(out.data_ptr()), st))
                torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"N={n:3d} {name:45s}: {o[1] / (reps * ks):7.1f} cycles per MMA (issue {o[0] / (reps * ks):.1f}); peak rate would be {n / 2:.0f}", flush=True)
