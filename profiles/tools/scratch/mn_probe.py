import ctypes as C, torch, sys
sys.path.insert(0, "/root/repo")
import azg_b200
from azg_b200 import _native as nat
lib = nat.lib(); dev = torch.device("cuda", 0)
h = C.c_void_p(); nat.check(lib.spl_ctx_create(2, 10, nat.RULES_DEFAULT, 0, C.byref(h)))
def run(n, k, sk8, sn8, lbo, sbo):
    g = torch.Generator(device="cpu").manual_seed(100 * n + k)
    a = torch.randn((128, k), generator=g).to(torch.bfloat16).to(dev)
    b = torch.randn((n, k), generator=g).to(torch.bfloat16).to(dev)
    out = torch.full((128, n), float("nan"), dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nat.check(lib.spl_umma_selftest_mn(h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), n, k, sk8, sn8, lbo, sbo, C.c_void_p(err.data_ptr()), st))
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    scale = (a.double().abs() @ b.double().abs().T).clamp_min(1e-6)
    e = float(((out.double() - ref).abs() / scale).nan_to_num(9.9).max())
    print(f"n={n} k={k} stride_k8={sk8} stride_n8={sn8} lbo={lbo} sbo={sbo} err_flag={int(err.item())} max_rel={e:.3e}", flush=True)
for n, k in [(224, 128), (64, 64), (32, 16), (64, 112)]:
    sk8, sn8 = 128, (k // 8) * 128
    run(n, k, sk8, sn8, sk8, sn8)      # LBO = K-direction core-matrix stride, SBO = N-direction
    run(n, k, sk8, sn8, sn8, sk8)      # swapped
# K core matrices far apart (stride_k8 = N/8 * 128: N-adjacent contiguous)
for n, k in [(224, 128), (64, 64)]:
    sn8, sk8 = 128, (n // 8) * 128
    run(n, k, sk8, sn8, sk8, sn8)
    run(n, k, sk8, sn8, sn8, sk8)
