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


@pytest.mark.parametrize("M,N,K", [(640, 400, 128), (300, 256, 128), (5000, 144, 64), (130, 128, 32), (257, 513, 40)])
def test_column_tiled_gemm_matches_float64(M, N, K):
    """mapf_tc_pack_b_tiles + mapf_tc_gemm_tiles: any number of output columns in one launch (grid.y = 128-column tile)."""
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    tiles = (N + 127) // 128
    Bp = torch.empty((tiles * lib.mapf_tc_packed_b_size(128, K),), dtype=torch.float32, device="cuda")
    _lib.check(lib.mapf_tc_pack_b_tiles(B.data_ptr(), K, N, Bp.data_ptr(), _lib.current_stream()), "pack_b_tiles")
    out = torch.full((M, N + 3), float("nan"), dtype=torch.float32, device="cuda")      # ldc > N: the pad stays untouched
    _lib.check(lib.mapf_tc_gemm_tiles(M, N, K, A.data_ptr(), K, Bp.data_ptr(), out.data_ptr(), N + 3, bias.data_ptr(), 0,
                                      _lib.current_stream()), "gemm_tiles")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t() + bias.double()
    err = (out[:, :N].double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 4e-6, err
    assert torch.isnan(out[:, N:]).all()


def tc_gemm_tn(A, B, transpose_c=False, accumulate_into=None):
    """C = A^T B through mapf_tc_gemm_tn (A [K, M], B [K, N] row-major, possibly strided views)."""
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    K, M = A.shape
    N = B.shape[1]
    if accumulate_into is not None:
        out = accumulate_into
    else:
        out = torch.full((N, M) if transpose_c else (M, N), float("nan"), dtype=torch.float32, device=A.device)
    rc = lib.mapf_tc_gemm_tn(M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(), out.stride(0),
                             int(transpose_c), int(accumulate_into is not None), _lib.current_stream())
    _lib.check(rc, "tc_gemm_tn")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,tr", [(128, 192, 40960, False), (128, 64, 40960, False), (400, 64, 5000, True),
                                      (128, 5, 3000, False), (64, 400, 2048, False), (5, 128, 777, True), (16, 16, 7, False),
                                      (130, 257, 33, False), (128, 128, 16, False)])
def test_tn_weight_gradient_gemm_matches_float64(M, N, K, tr):
    """dW = A^T B with the contraction over the row index (split K, transposing producers): fp32-level accuracy against
    float64, every tile / split / tail shape, row-major and transposed output, unaligned leading dimensions."""
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + K)
    A = torch.randn(K, M, device="cuda", generator=g)
    B = torch.randn(K, N, device="cuda", generator=g) * 0.1
    ref = A.double().t() @ B.double()
    out = tc_gemm_tn(A, B, transpose_c=tr)
    if tr:
        out = out.t()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    err32 = ((A.t() @ B).double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M} N={N} K={K}: 3xTF32 TN rel err {err:.2e}, cuBLAS fp32 {err32:.2e}")
    assert err < 4e-6, err
    # strided operands (columns of wider matrices) and accumulation into an existing C
    wide_a = torch.randn(K, M + 12, device="cuda", generator=g)
    wide_b = torch.randn(K, N + 8, device="cuda", generator=g)
    Av, Bv = wide_a[:, 4:4 + M], wide_b[:, 8:8 + N]
    base = torch.randn(M, N, device="cuda", generator=g)
    out2 = tc_gemm_tn(Av, Bv, accumulate_into=base.clone())
    ref2 = base.double() + Av.double().t() @ Bv.double()
    assert (out2.double() - ref2).abs().max().item() / ref2.abs().max().item() < 4e-6


def test_tn_structured_values_are_exact():
    K, M, N = 64, 128, 128
    A = (torch.arange(K * M, device="cuda", dtype=torch.float32).reshape(K, M) % 7) - 3
    B = (torch.arange(K * N, device="cuda", dtype=torch.float32).reshape(K, N) % 5) - 2
    assert torch.equal(tc_gemm_tn(A, B), A.t() @ B)


@pytest.mark.parametrize("B,K,self_loop", [(7, 4, True), (2, 2, False), (33, 3, True), (1, 4, True)])
def test_tensor_core_hops_match_ffma_taps_and_float64(B, K, self_loop):
    """Filter taps of 64-agent teams (gnn.py:87-109): the tensor-core hops kernel (tc_hops.cu, block-diagonal S per pair of
    episodes, 3xTF32) against the FFMA taps kernel and a float64 restatement; odd episode counts cover the half tile."""
    from mapf_gnn_b200 import _lib
    lib = _lib.load()
    N, F = 64, 64
    g = torch.Generator(device="cuda").manual_seed(B * 10 + K)
    x = torch.randn(B * N, F, device="cuda", generator=g)
    adj = (torch.rand(B, N, N, device="cuda", generator=g) < 0.08).float()
    adj[:, torch.arange(N), torch.arange(N)] = 0
    if not self_loop:
        adj[0, 5, :] = 0          # zero-degree row: inf -> 0 (gnn.py:185-199)

    def run(fn):
        z = torch.full((B * N, K * F), float("nan"), device="cuda")
        a = _lib.GraphFilterFwdArgs()
        a.batch, a.agents, a.in_dim, a.out_dim, a.taps = B, N, F, 128, K
        a.add_self_loop, a.has_skip = int(self_loop), 0
        a.x, a.x_sb, a.x_sn, a.x_sf = x.data_ptr(), N * F, F, 1
        a.gso, a.z = adj.data_ptr(), z.data_ptr()
        _lib.check(fn(a, _lib.current_stream()), "graph_taps")
        torch.cuda.synchronize()
        return z

    z_tc, z_ff = run(lib.mapf_graph_taps), run(lib.mapf_graph_taps_ffma)
    A = adj.double() + (torch.eye(N, device="cuda", dtype=torch.float64) if self_loop else 0)
    dinv = A.sum(dim=2).pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    S = dinv[:, :, None] * A * dinv[:, None, :]
    xk = x.double().view(B, N, F)
    taps = [xk]
    for _ in range(1, K):
        taps.append(torch.einsum("bij,bif->bjf", S, taps[-1]))
    ref = torch.cat(taps, dim=2).reshape(B * N, K * F)
    scale = ref.abs().max().item()
    assert not torch.isnan(z_tc).any()
    assert (z_tc.double() - ref).abs().max().item() / scale < 5e-6
    assert (z_ff.double() - ref).abs().max().item() / scale < 5e-6
    assert torch.equal(z_tc[:, :F], x) and torch.equal(z_ff[:, :F], x)
