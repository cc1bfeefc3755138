"""tcgen05 / TMA building blocks in isolation: one 128x128x128 bf16 GEMM per operand form
(include/nnop_b200.h, nnop_selftest_umma) against a torch fp32 matmul of the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", [0, 1, 2, 3, 4])
def test_umma_operand_forms(nnop, which):
    g = torch.Generator().manual_seed(which)
    a = torch.randn(128, 128, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(128, 128, generator=g).to(torch.bfloat16).cuda()
    d = nnop.selftest_umma(a, b, which)
    torch.cuda.synchronize()
    af, bf = a.float(), b.float()
    ref = {0: af @ bf.T, 1: af @ bf, 2: af @ bf.T, 3: af.T @ bf, 4: af @ bf}[which]
    err = (d - ref).abs().max().item()
    assert err < 1e-3, f"which={which}: max abs err {err}"
