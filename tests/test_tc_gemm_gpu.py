"""tcgen05 + TMA GEMM building block (pmg_tc_gemm) against torch fp64 matmul on the same bf16 operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a_hi, a_lo, b, bn, splits=1):
    from pmg_b200 import _lib
    M, K = a_hi.shape
    N = b.shape[0]
    total_kb = K // 64
    per = -(-total_kb // splits)
    eff = -(-total_kb // per)
    out = torch.full((eff, M, N), float("nan"), device="cuda", dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib.pmg_tc_gemm(a_hi.data_ptr(), a_lo.data_ptr() if a_lo is not None else None, b.data_ptr(), M, N, K, bn,
                              splits, out.data_ptr(), st)
    _lib.check(rc, "pmg_tc_gemm")
    torch.cuda.synchronize()
    return out.sum(0) if splits > 1 else out[0]


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 32, 64, 32), (300, 96, 128, 96), (1000, 128, 512, 128),
                                      (77, 192, 192, 192), (640, 576, 192, 192), (513, 384, 384, 192)])
def test_single_plane(M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.double() @ b.double().T
    got = _gemm(a, None, b, bn)
    err = (got.double() - ref).norm() / ref.norm()
    assert torch.isfinite(got).all()
    assert err.item() < 1e-5, err.item()


@pytest.mark.parametrize("M,N,K,bn,splits", [(200, 64, 512, 64, 1), (384, 192, 9216, 192, 4), (90, 192, 9216, 192, 16)])
def test_split_activation_and_split_k(M, N, K, bn, splits):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g)
    hi = a.to(torch.bfloat16)
    lo = (a - hi.float()).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.double() @ b.double().T
    got = _gemm(hi, lo, b, bn, splits)
    err = (got.double() - ref).norm() / ref.norm()
    # two bf16 planes carry ~16 significant bits of the fp32 activation
    assert err.item() < 2e-5, err.item()
    single = _gemm(hi, None, b, bn, splits)
    assert ((single.double() - ref).norm() / ref.norm()).item() > 1e-4   # sanity: lo plane matters


@pytest.mark.parametrize("M,K,splits,planes", [(256, 64, 1, 1), (256, 128, 1, 2), (300, 512, 2, 2), (9216, 6144, 2, 2), (90, 3072, 16, 1)])
def test_cta_pair_gemm(M, K, splits, planes):
    """tcgen05 cta_group::2 split-K GEMM (tc_gemm2.cuh), N = 192."""
    from pmg_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    hi = a.to(torch.bfloat16)
    lo = (a - hi.float()).to(torch.bfloat16)
    b = torch.randn(192, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = (a.double() if planes == 2 else hi.double()) @ b.double().T
    total_kb = K // 64
    per = -(-total_kb // splits)
    eff = -(-total_kb // per)
    out = torch.full((eff, M, 192), float("nan"), device="cuda", dtype=torch.float32)
    rc = _lib.lib.pmg_tc_gemm2(hi.data_ptr(), lo.data_ptr() if planes == 2 else None, b.data_ptr(), M, K, splits, out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "pmg_tc_gemm2")
    torch.cuda.synchronize()
    got = out.sum(0)
    assert torch.isfinite(got).all()
    err = ((got.double() - ref).norm() / ref.norm()).item()
    assert err < 3e-5, err
