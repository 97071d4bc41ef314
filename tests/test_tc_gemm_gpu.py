"""GPU: the tcgen05 tile engine as a plain GEMM against torch (bf16 operands, fp32 accumulate), all four operand-major
combinations, split-K, ragged edges.  This pins the UMMA shared-memory / instruction descriptors every tensor-core
kernel of the learner relies on.  Tolerance: fp32 accumulation of exact bf16 products -> 1e-5 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_gemm(M, N, K, a_mn, b_mn, splits, seed=0):
    from isdqn_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
    a_store = A.t().contiguous() if a_mn else A.contiguous()   # [K][M] or [M][K]
    b_store = B.t().contiguous() if b_mn else B.contiguous()   # [K][N] or [N][K]
    C = torch.full((splits, M, N), float("nan"), device="cuda")
    _lib.check(lib.isdqn_tc_gemm_bf16(a_store.data_ptr(), a_store.stride(0), int(a_mn), b_store.data_ptr(), b_store.stride(0),
                                      int(b_mn), C.data_ptr(), M, N, K, splits, _lib.stream_ptr()), "tc_gemm")
    torch.cuda.synchronize()
    want = A.double() @ B.double().t()
    n_chunks = -(-K // 64)
    cps = -(-n_chunks // splits)
    real = -(-n_chunks // cps)
    got = C[:real].double().sum(0)
    err = float((got - want).abs().max() / want.abs().max())
    return err


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_majors_single_tile(a_mn, b_mn):
    for (M, N, K) in [(128, 64, 64), (128, 32, 128), (128, 256, 192), (128, 128, 64)]:
        err = run_gemm(M, N, K, a_mn, b_mn, 1)
        assert err < 1e-5, f"M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}"


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_ragged_and_multi_tile(a_mn, b_mn):
    for (M, N, K) in [(64, 512, 7744), (7744, 512, 32), (32, 7744, 512), (200, 72, 104), (1000, 520, 24)]:
        err = run_gemm(M, N, K, a_mn, b_mn, 1, seed=1)
        assert err < 2e-5, f"M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}"


def test_split_k_partials():
    for splits in (2, 7, 32):
        err = run_gemm(64, 512, 7744, 0, 1, splits, seed=2)
        assert err < 2e-5, f"splits {splits}: rel err {err:.3e}"
