"""tcgen05.mma pacing with a tcgen05.commit after every group of k-steps (csrc/spl_umma.cu: spl_umma_mma_cycles, b_mn bit 1)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.realpath(os.path.join(os.path.dirname(__file__), "..", "..")))
import azg_b200
from azg_b200 import _native as nat
lib = nat.lib(); dev = torch.device("cuda", 0)
h = C.c_void_p(); nat.check(lib.spl_ctx_create(2, 10, nat.RULES_DEFAULT, 0, C.byref(h)))
out = torch.zeros(2, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for n, sbo in ((64, 2048),):
    for mode, name in ((1 + 4, "lean, one thread"), (1 + 4 + 2, "+ commit per 4"), (1 + 4 + 2 + 16, "+ commit + wait on a completed barrier"),
                       (1 + 4 + 2 + 32, "+ commit + fence::after_thread_sync"), (1 + 4 + 2 + 16 + 32, "+ commit + wait + fence"),
                       (1 + 4 + 2 + 16 + 32 + 64, "+ commit + wait + fence + clock store")):
        reps, ks = 16, 8
        for _ in range(2):
            nat.check(lib.spl_umma_mma_cycles(h, n, ks, mode, sbo, reps, C.c_void_p(out.data_ptr()), st))
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        print(f"N={n:3d} {name:48s}: {o[1] / (reps * ks):7.1f} cycles per MMA (issue {o[0] / (reps * ks):.1f}); per group of 4: {4 * o[1] / (reps * ks):.0f}", flush=True)
