"""Hardware self-test of the tcgen05 building blocks (tile layout, UMMA descriptors, TMEM round trip) that the fused
field kernels are made of: one 128xNxK tile against torch.matmul, K-major and MN-major operands."""
import itertools

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [(N, K, a, b) for (N, K), (a, b) in itertools.product(
    [(128, 128), (128, 64), (128, 32), (128, 16), (64, 128), (32, 128), (16, 128), (16, 16), (32, 32)],
    [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 1)])]  # a_mn == 2: A operand staged in tensor memory


@pytest.mark.parametrize("N,K,a_mn,b_mn", CASES)
def test_single_tile_gemm(N, K, a_mn, b_mn, built_lib, cuda):
    from stable_nerf_b200 import _lib
    from stable_nerf_b200._lib import check, ptr, stream
    built_lib = _lib.load_debug()  # the self-test of the tcgen05 building blocks lives in the debug build only
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K * 10 + a_mn * 2 + b_mn)
    A = torch.randn(128, K, generator=g).to(cuda)
    B = torch.randn(N, K, generator=g).to(cuda)
    D = torch.full((128, N), float("nan"), device=cuda)
    check(built_lib.snerf_tc_selftest(ptr(A), ptr(B), ptr(D), N, K, a_mn, b_mn, stream()), "tc selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    err = (D - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, f"N={N} K={K} a_mn={a_mn} b_mn={b_mn}: rel err {err}"
