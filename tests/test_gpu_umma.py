"""GPU test of the tcgen05 / TMEM building blocks (csrc/spl_umma.cuh) through spl_umma_selftest: a 128 x N x K bf16 product on the
5th-generation tensor cores (shared-memory descriptors, tcgen05.mma, tcgen05.commit -> mbarrier, tcgen05.ld) against a
float64 product of the same bf16 inputs. Tolerance: fp32 accumulation of exact bf16 products, 1e-4 relative to the row scale."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(128, 128), (128, 64), (64, 96), (256, 128), (128, 16), (32, 112), (192, 256)])
def test_umma_product(n, k):
    import azg_b200
    from azg_b200 import _native as nat
    lib = nat.lib()
    dev = torch.device("cuda", 0)
    h = C.c_void_p()
    nat.check(lib.spl_ctx_create(2, 10, nat.RULES_DEFAULT, 0, C.byref(h)))
    try:
        g = torch.Generator(device="cpu").manual_seed(100 * n + k)
        a = torch.randn((128, k), generator=g).to(torch.bfloat16).to(dev)
        b = torch.randn((n, k), generator=g).to(torch.bfloat16).to(dev)
        out = torch.full((128, n), float("nan"), dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        nat.check(lib.spl_umma_selftest(h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), n, k,
                                        C.c_void_p(err.data_ptr()), st))
        torch.cuda.synchronize()
        assert int(err.item()) == 0, "completion barrier timed out"
        ref = a.double() @ b.double().T
        scale = (a.double().abs() @ b.double().abs().T).clamp_min(1e-6)
        assert torch.isfinite(out).all()
        assert float(((out.double() - ref).abs() / scale).max()) < 1e-4
    finally:
        lib.spl_ctx_destroy(h)
