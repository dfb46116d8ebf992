"""tcgen05 GEMM tiles (engine 1) against an fp64 matmul: single bf16 product and the two-term split."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, nprod, seed=0):
    from avatarclip_b200 import _lib
    L = _lib.lib()
    L.avc_tc_gemm_nt_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]
    L.avc_tc_gemm_nt_test.restype = C.c_int
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) * 0.1).cuda()
    Cm = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(4 * (M + N) * ((K + 7) // 8 * 8) + 8192, dtype=torch.uint8, device="cuda")
    _lib.check(L.avc_tc_gemm_nt_test(A.data_ptr(), B.data_ptr(), M, N, K, nprod, Cm.data_ptr(), ws.data_ptr(),
                                     ws.numel(), _lib.stream_ptr()), "avc_tc_gemm_nt_test")
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (Cm.double() - ref).abs().max().item() / ref.abs().max().item()
    return err


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 256), (1000, 217, 256), (300, 39, 256), (4096, 256, 40),
                                   (513, 128, 320)])
def test_tc_gemm_split3(M, N, K):
    err = _run(M, N, K, 3)
    print(M, N, K, "split-3 rel-to-max err", err)
    assert err < 3e-5


@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (1000, 217, 256)])
def test_tc_gemm_single(M, N, K):
    err = _run(M, N, K, 1)
    print(M, N, K, "single-bf16 rel-to-max err", err)
    assert err < 2e-2


def _run_tn(P, N1, N2, nprod, seed=0):
    from avatarclip_b200 import _lib
    L = _lib.lib()
    L.avc_tc_gemm_tn_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.avc_tc_gemm_tn_test.restype = C.c_int
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(P, N1, generator=g).cuda()
    B = (torch.randn(P, N2, generator=g) * 0.1).cuda()
    base = torch.randn(N1, N2, generator=g).cuda()
    Cm = base.clone()
    r8 = lambda n: (n + 7) // 8 * 8
    ws = torch.empty(4 * P * (r8(N1) + r8(N2)) + 8192, dtype=torch.uint8, device="cuda")
    cs = torch.ones(N1, device="cuda")
    _lib.check(L.avc_tc_gemm_tn_test(A.data_ptr(), B.data_ptr(), P, N1, N2, nprod, Cm.data_ptr(), cs.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "avc_tc_gemm_tn_test")
    torch.cuda.synchronize()
    ref = base.double() + A.double().t() @ B.double()
    cref = 1.0 + A.double().sum(0)
    cerr = (cs.double() - cref).abs().max().item() / cref.abs().max().item()
    assert cerr < 3e-5, ("fused column sum", cerr)
    return (Cm.double() - ref).abs().max().item() / ref.abs().max().item()


@pytest.mark.parametrize("P,N1,N2", [(64, 128, 64), (640, 256, 256), (5000, 217, 256), (3000, 257, 39), (9000, 40, 128)])
def test_tc_gemm_tn_split3(P, N1, N2):
    err = _run_tn(P, N1, N2, 3)
    print(P, N1, N2, "TN split-3 rel-to-max err", err)
    assert err < 3e-5
