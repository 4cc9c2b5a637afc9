"""tcgen05 3xTF32 GEMM vs a float64 reference: the split must recover fp32-level accuracy (plain TF32 would be ~1e-3)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tc_gemm(A, B, bias=None, relu=False):
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    Bp = torch.empty((lib.mapf_tc_packed_b_size(N, K),), dtype=torch.float32, device=A.device)
    _lib.check(lib.mapf_tc_pack_b(B.data_ptr(), K, None, 0, N, Bp.data_ptr(), _lib.current_stream()), "tc_pack_b")
    rc = lib.mapf_tc_gemm(M, N, K, A.data_ptr(), A.stride(0), Bp.data_ptr(), out.data_ptr(), N,
                          None if bias is None else bias.data_ptr(), int(relu), _lib.current_stream())
    _lib.check(rc, "tc_gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (128, 64, 64), (128, 128, 128), (300, 64, 400), (1000, 128, 192),
                                   (77, 40, 36), (4096, 128, 256), (20480, 64, 400)])
def test_matches_float64(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ B.double().t() + bias.double())
    fp32 = torch.addmm(bias, A, B.t())
    err32 = (fp32.double() - ref).abs().max().item() / ref.abs().max().item()
    out = tc_gemm(A, B, bias)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M} N={N} K={K}: 3xTF32 rel err {err:.2e}, cuBLAS fp32 rel err {err32:.2e}")
    assert err < 4e-6, err
    out = tc_gemm(A, B, bias, relu=True)
    err = (out.double() - ref.clamp_min(0)).abs().max().item() / ref.abs().max().item()
    assert err < 4e-6, err


def test_structured_values_are_exact():
    """Small integers are exactly representable in TF32: the product must be exact (layout / descriptor check)."""
    M, N, K = 128, 64, 32
    A = torch.arange(M * K, device="cuda", dtype=torch.float32).reshape(M, K) % 7
    B = (torch.arange(N * K, device="cuda", dtype=torch.float32).reshape(N, K) % 5) - 2
    out = tc_gemm(A, B)
    assert torch.equal(out, A @ B.t())
