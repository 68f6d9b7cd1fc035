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


@pytest.mark.parametrize("kind,obs_shape,n_actions,n", [
    ("CustomPolicyBigMap", (28, 28, 1), 3, 700), ("CustomPolicySmallMap", (10, 10, 5), 9, 130),
    ("CustomPolicyBigMap", (22, 22, 8), 12, 257), ("FullyConvPolicyBigMap", (14, 14, 1), 14 * 14 * 2, 96),
    ("FullyConvPolicySmallMap", (5, 5, 5), 125, 64)])
def test_native_policy_forward_matches_torch(kind, obs_shape, n_actions, n):
    """policy_native.NativePolicy (im2col + tcgen05 GEMM for every layer, bf16 activations, fp32 accumulation) against the
    torch fp32 modules of models.py.  Tolerance: bf16 rounding of weights and of 4-9 activation tensors; stated as
    |err| <= 0.03 * max|ref| + 0.02 per output tensor."""
    import torch
    from gym_pcgrl_b200.models import ActorCritic
    from gym_pcgrl_b200.policy_native import NativePolicy
    torch.manual_seed(n)
    net = ActorCritic(kind, obs_shape, n_actions).cuda()
    with torch.no_grad():
        for p in net.parameters():          # non-trivial biases and heads
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    hi = 2 if obs_shape[2] > 1 else 2
    obs = torch.randint(0, hi, (n,) + obs_shape, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        want_logits, want_value = net(obs)
    got_logits, got_value = NativePolicy(net)(obs)
    torch.cuda.synchronize()
    assert tuple(got_logits.shape) == tuple(want_logits.shape) and tuple(got_value.shape) == tuple(want_value.shape)
    for got, want in ((got_logits, want_logits), (got_value, want_value)):
        tol = 0.03 * float(want.abs().max()) + 0.02
        assert float((got.float() - want).abs().max()) <= tol, (kind, float((got.float() - want).abs().max()), tol)


@pytest.mark.parametrize("n,h,w,cout,relu", [(3, 5, 5, 64, 1), (7, 14, 14, 64, 1), (2, 16, 11, 5, 0), (33, 7, 11, 8, 1), (1, 1, 1, 64, 1),
                                             (130, 14, 14, 2, 1)])
def test_implicit_gemm_conv3x3_matches_torch(n, h, w, cout, relu):
    """pcgrl_conv3x3_bf16 (tcgen05 implicit GEMM on a zero-bordered NHWC buffer: nine shifted TMA loads per tile, weights
    resident in shared memory) against torch conv2d in fp32 on the same bf16-rounded inputs.  Tolerance: fp32 accumulation of
    576 bf16 products + one bf16 rounding of the output: |err| <= 0.01 * max|ref| + 0.01.  The border of the output must be
    exactly zero (it is the next layer's padding)."""
    import torch
    import torch.nn.functional as F
    from gym_pcgrl_b200 import _native
    torch.manual_seed(n * 100 + h)
    cin = 64
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (0.1 * torch.randn(cout, cin, 3, 3, device="cuda")).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda")
    npad = 64 if cout > 32 else (cout + 7) // 8 * 8
    xp = torch.zeros(n, h + 2, w + 2, cin, dtype=torch.bfloat16, device="cuda")
    xp[:, 1:-1, 1:-1] = x
    w2 = torch.zeros(npad, 3, 3, cin, dtype=torch.bfloat16, device="cuda")
    w2[:cout] = wt.permute(0, 2, 3, 1)
    b2 = torch.zeros(npad, device="cuda")
    b2[:cout] = bias
    yp = torch.full((n, h + 2, w + 2, npad), 7.0, dtype=torch.bfloat16, device="cuda")      # poisoned: the kernel must write the border
    rc = _native.lib().pcgrl_conv3x3_bf16(xp.data_ptr(), w2.reshape(npad, -1).contiguous().data_ptr(), b2.data_ptr(), yp.data_ptr(),
                                          n, h, w, cin, npad, relu, _native.stream_ptr(torch.device("cuda", 0)))
    assert rc == 0, _native.lib().pcgrl_linear_last_error()
    torch.cuda.synchronize()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    if relu:
        want = want.relu()
    want = want.permute(0, 2, 3, 1)
    got = yp[:, 1:-1, 1:-1, :cout].float()
    tol = 0.01 * float(want.abs().max()) + 0.01
    assert float((got - want).abs().max()) <= tol, (float((got - want).abs().max()), tol)
    assert float(yp[:, 0].abs().max()) == 0 and float(yp[:, -1].abs().max()) == 0
    assert float(yp[:, :, 0].abs().max()) == 0 and float(yp[:, :, -1].abs().max()) == 0
    if npad > cout:
        assert float(yp[..., cout:].abs().max()) == 0
