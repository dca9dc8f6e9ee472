"""GPU tests of the tcgen05 GEMM layer (gemm_img.cuh) through the C-ABI self-test hook: the three
operand orientations of the path (forward, data gradient, weight gradient) against float64
matmul, at the path's shapes and at ragged ones.  bf16x3 is fp32-grade: tolerance 2e-5 of the
output scale (plain bf16 would sit near 4e-3), widened by sqrt(K/4096) for very long reductions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(got, want, what, K):
    """fp32 accumulation noise grows ~sqrt(K): the bound is 2e-5 of the output scale up to K = 4096
    and widens with sqrt(K/4096) beyond (the weight gradient reduces over 105,600 token rows)."""
    scale = want.abs().max().item()
    err = (got.double() - want).abs().max().item()
    tol = 2e-5 * scale * max(1.0, (K / 4096.0) ** 0.5)
    assert err <= tol, f"{what}: max err {err:.3e} vs scale {scale:.3e} (tol {tol:.3e})"


@pytest.mark.parametrize("M,N,K", [(128, 240, 64), (300, 900, 300), (1000, 200, 300), (77, 60, 20), (4096, 900, 300)])
def test_forward_orientation(M, N, K, built_lib):
    from pytorch_news_recommender_b200 import ops
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    _check(ops.gemm_selftest(0, A, B), A.double() @ B.double().t(), f"NT {M}x{N}x{K}", K)


@pytest.mark.parametrize("M,N,K", [(128, 320, 64), (300, 300, 900), (1000, 300, 200), (77, 20, 60), (4096, 300, 900)])
def test_data_gradient_orientation(M, N, K, built_lib):
    from pytorch_news_recommender_b200 import ops
    torch.manual_seed(M + N + K + 1)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(K, N, device="cuda")
    _check(ops.gemm_selftest(1, A, B), A.double() @ B.double(), f"NN {M}x{N}x{K}", K)


@pytest.mark.parametrize("M,N,K", [(128, 320, 64), (900, 300, 3000), (200, 300, 3000), (60, 20, 130), (900, 300, 105600)])
def test_weight_gradient_orientation(M, N, K, built_lib):
    from pytorch_news_recommender_b200 import ops
    torch.manual_seed(M + N + K + 2)
    A = torch.randn(K, M, device="cuda")
    B = torch.randn(K, N, device="cuda")
    _check(ops.gemm_selftest(2, A, B), A.double().t() @ B.double(), f"TN {M}x{N}x{K}", K)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 64), (300, 900, 300), (1000, 200, 300), (77, 60, 20),
                                   (4096, 960, 300), (105600, 960, 300)])
def test_forward_orientation_on_cta_pairs(M, N, K, built_lib):
    """The forward projection on clusters of two CTAs (tcgen05.mma.cta_group::2): odd and even token-tile
    counts, a ragged last tile, several N tiles, and the benchmark's own shape."""
    from pytorch_news_recommender_b200 import ops
    torch.manual_seed(M + N + K + 3)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    _check(ops.gemm_selftest(3, A, B), A.double() @ B.double().t(), f"NT pair {M}x{N}x{K}", K)
