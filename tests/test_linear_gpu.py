"""csrc/pcgrl_linear.cu (tcgen05.mma + TMEM + TMA) against a plain PyTorch fp32 reference of the same op:
relu(x @ w.T + b) on bf16-rounded inputs.  Tolerance: bf16 products are exact in fp32 and the accumulation is fp32, so the
only difference is summation order -- |err| <= 2e-3 * (1 + |ref|) for K <= 4096 with unit-scale inputs (stated here)."""
import pytest

pytestmark = pytest.mark.gpu

CASES = [(4096, 512, 1024, True), (128, 128, 64, False), (300, 132, 200, True), (1, 4, 8, False), (1000, 512, 3136, True),
         # enough 128 x 256 tiles to fill the SMs: the BLOCK_N = 256 instantiation (full and ragged in M and N)
         (32768, 512, 256, True), (20001, 300, 136, False)]


@pytest.mark.parametrize("m,n,k,relu", CASES)
def test_tcgen05_linear_matches_torch(m, n, k, relu):
    import torch
    from gym_pcgrl_b200 import _native
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    x = torch.randn((m, k), generator=g, device="cuda").bfloat16()
    w = (torch.randn((n, k), generator=g, device="cuda") / k ** 0.5).bfloat16()
    b = torch.randn(n, generator=g, device="cuda")
    y = _native.linear_bf16(x, w, b, relu=relu)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if relu:
        ref = torch.relu(ref)
    err = (y - ref).abs()
    tol = 2e-3 * (1 + ref.abs())
    assert bool((err <= tol).all()), (float(err.max()), float(ref.abs().max()))
    assert y.dtype == torch.float32 and tuple(y.shape) == (m, n)


def test_tcgen05_linear_inside_the_policy():
    """Cnn2's fc1 through the hand-written kernel (PCGRL_TCGEN05_FC semantics: inference only) == the torch layer."""
    import torch
    from gym_pcgrl_b200.models import Cnn2
    net = Cnn2((28, 28, 1)).cuda()
    obs = torch.randint(0, 2, (512, 28, 28, 1), dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        want = net(obs)
        net.use_tcgen05_fc = True
        got = net(obs)
    assert float((got - want).abs().max()) <= 2e-2 * (1 + float(want.abs().max()))
