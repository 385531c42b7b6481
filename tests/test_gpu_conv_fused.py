"""GPU: the fused conv epilogues through the C-ABI against plain PyTorch fp32 on the same fp16-rounded inputs.

* cy4_conv_fwd_fused: y = act(conv(x, w) + shift[c]) (+ residual) -- eval-mode Conv2d+BatchNorm2d(+Mish /
  LeakyReLU)(+shortcut) of models/darknet2pytorch.py:247-278,208-219 as one kernel;
* cy4_conv_dgrad_fused: the input gradient fused with the first pass of the producer's BN/activation backward
  (dz = dA_total * act'(scale*Y + shift), per-channel sum dz and sum dz*Y)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACTS = {0: lambda z: z, 1: lambda z: F.leaky_relu(z, 0.1), 2: lambda z: z * torch.tanh(F.softplus(z))}


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,act,res", [
    (2, 38, 128, 256, 3, 1, 2, False), (2, 38, 256, 128, 1, 1, 2, True), (4, 76, 64, 64, 3, 1, 1, True),
    (2, 76, 64, 128, 3, 2, 2, False), (3, 19, 512, 1024, 3, 1, 1, False), (2, 40, 32, 64, 3, 1, 0, True),
    (8, 152, 64, 64, 1, 1, 2, True),          # more tiles than SMs, packed k-blocks, 4 rotating output slabs
])
def test_conv_fwd_fused(B, H, Cin, Cout, k, stride, act, res):
    from cy4 import convops as co
    torch.manual_seed(B + H + Cin + Cout + act)
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
    shift = torch.randn(Cout, device="cuda") * 0.5
    big = torch.randn(B, Ho, Ho, Cout + 32, device="cuda").half()
    r = big[..., 32:] if res else None                        # residual read from a channel slice (ld > C)
    outbuf = torch.zeros(B, Ho, Ho, Cout + 64, device="cuda", dtype=torch.float16)
    y = co.conv_fwd_fused(x, co.pack_fprop(w.float()), Cout, k, stride, pad, shift, act, residual=r, out=outbuf[..., 64:])
    ref = ACTS[act](F.conv2d(_nchw(x), w.float(), None, stride, pad) + shift[None, :, None, None]).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    assert float(outbuf[..., :64].abs().max()) == 0.0


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,act,accum", [
    (2, 38, 128, 256, 3, 1, 2, False),        # TMA-store epilogue
    (2, 38, 256, 128, 1, 1, 2, True),         # read-modify-write epilogue (last writer of a shared gradient)
    (2, 76, 64, 128, 3, 2, 1, False),         # stride 2: four parity-class launches share the statistics
    (2, 76, 64, 128, 3, 2, 2, True),
    (4, 76, 64, 64, 3, 1, 0, False),          # linear activation: dz = dA, sums only
    (8, 152, 64, 64, 1, 1, 2, True),
])
def test_conv_dgrad_fused(B, H, Cin, Cout, k, stride, act, accum):
    from cy4 import convops as co
    torch.manual_seed(B + H + Cin + Cout + act + 1)
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    dy = torch.randn(B, Ho, Ho, Cout, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cout * k * k) ** 0.5).half()
    yprod_big = (torch.randn(B, H, H, Cin + 32, device="cuda") * 1.5).half()
    yprod = yprod_big[..., :Cin]                                  # ld > C
    sc = torch.rand(Cin, device="cuda") + 0.5
    sh = torch.randn(Cin, device="cuda") * 0.3
    old = torch.randn(B, H, H, Cin, device="cuda").half()
    xr = torch.zeros(B, Cin, H, H, device="cuda", requires_grad=True)
    F.conv2d(xr, w.float(), None, stride, pad).backward(_nchw(dy))
    dA = xr.grad.permute(0, 2, 3, 1) + (old.float() if accum else 0)
    z = (yprod.float() * sc + sh).requires_grad_(True)
    ACTS[act](z).backward(torch.ones_like(z))
    dz_ref = dA * z.grad
    s1 = torch.zeros(Cin, device="cuda"); s2 = torch.zeros(Cin, device="cuda")
    out = old.clone() if accum else torch.empty(B, H, H, Cin, device="cuda", dtype=torch.float16)
    co.conv_dgrad(dy, co.pack_dgrad(w.float()), H, H, Cin, k, stride, pad, out=out, accumulate=accum, fuse=(yprod, sc, sh, act, s1, s2))
    err = (out.float() - dz_ref).abs().max().item()
    assert err <= 2e-3 * dz_ref.abs().max().item() + 1e-3, err
    r1 = dz_ref.sum((0, 1, 2)); r2 = (dz_ref * yprod.float()).sum((0, 1, 2))
    n = dz_ref.numel() / Cin
    assert (s1 - r1).abs().max().item() <= 2e-3 * dz_ref.abs().mean().item() * n ** 0.5 + 1e-2 * r1.abs().max().item()
    assert (s2 - r2).abs().max().item() <= 2e-3 * (dz_ref * yprod.float()).abs().mean().item() * n ** 0.5 + 1e-2 * r2.abs().max().item()
