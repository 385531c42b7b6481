"""GPU: every distinct conv shape of complex_yolov4.cfg @608 at the BENCH batch
(B=32): fprop (fp32-output epilogue), dgrad and wgrad through the C-ABI against a plain PyTorch fp32
reference of the same op (F.conv2d / its autograd, TF32 off) on identical fp16-rounded inputs.

This is the parity check AT the bench configuration: tile counts far above 148 (persistent loop, TMEM
double buffering, k-block packing), the wave-aware split-K of the weight gradient, and the four parity
classes of the stride-2 input gradient all take different branches here than at the small test shapes.
Tolerances: fprop |err| <= 1e-3 absolute (BASELINE.json: conv activations 1e-3 fp32) on unit-variance
outputs; dgrad is stored in fp16 (2e-3 of max|ref| = 2 ulp of fp16); wgrad accumulates in fp32
(1e-3 of max|ref|, split-K summation order)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# Cin, Cout, k, stride, Hin  (the stem 3->32 goes through the im2col-matrix path and is covered by test_stem_path
# and the engine tests; here it appears as its GEMM form 32->32 1x1 @608)
SHAPES = [
    (512, 1024, 3, 1, 19), (256, 512, 3, 1, 38), (128, 128, 3, 1, 76), (256, 256, 3, 1, 38), (128, 256, 3, 1, 76),
    (512, 512, 3, 1, 19), (32, 64, 3, 2, 608), (32, 64, 3, 1, 304), (64, 128, 3, 2, 304), (64, 64, 3, 1, 152),
    (128, 256, 3, 2, 152), (256, 512, 3, 2, 76), (512, 1024, 3, 2, 38), (512, 256, 1, 1, 38), (1024, 512, 1, 1, 19),
    (64, 64, 1, 1, 304), (256, 128, 1, 1, 76), (128, 128, 1, 1, 76), (256, 256, 1, 1, 38), (128, 64, 1, 1, 304),
    (512, 512, 1, 1, 19), (128, 256, 3, 2, 76), (256, 512, 3, 2, 38), (128, 64, 1, 1, 152), (128, 128, 1, 1, 152),
    (256, 256, 1, 1, 76), (512, 512, 1, 1, 38), (1024, 1024, 1, 1, 19), (2048, 512, 1, 1, 19), (64, 64, 1, 1, 152),
    (64, 32, 1, 1, 304), (512, 256, 1, 1, 19), (256, 128, 1, 1, 38), (256, 30, 1, 1, 76), (512, 30, 1, 1, 38),
    (1024, 30, 1, 1, 19),
]
B = 32


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


@pytest.mark.parametrize("Cin,Cout,k,stride,H", SHAPES)
def test_bench_shape_fprop_dgrad_wgrad(Cin, Cout, k, stride, H):
    from cy4 import convops as co
    torch.manual_seed(Cin * 7 + Cout * 3 + k + stride + H)
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
    # ---- fprop, fp32 output
    y = co.conv_fwd(x, co.pack_fprop(w.float()), Cout, k, stride, pad, out_f32=True)
    ref = F.conv2d(_nchw(x), w.float(), None, stride, pad).permute(0, 2, 3, 1)
    err = (y[..., :Cout] - ref).abs().max().item()
    assert err <= 1e-3, "fprop %g" % err
    # ---- fprop, fp16 output through the TMA-store epilogue + fused BN statistics
    if Cout % 32 == 0:
        s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
        y16 = co.conv_fwd(x, co.pack_fprop(w.float()), Cout, k, stride, pad, stats=(s1, s2))
        assert (y16.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
        n = ref.numel() / Cout
        torch.testing.assert_close(s1 / n, ref.sum((0, 1, 2)) / n, rtol=1e-3, atol=2e-4)
        torch.testing.assert_close(s2 / n, (ref * ref).sum((0, 1, 2)) / n, rtol=1e-3, atol=1e-4)
        del y16
    del y
    # ---- dgrad / wgrad against autograd of the same conv
    cpad = (Cout + 63) // 64 * 64
    dyb = torch.zeros(B, Ho, Ho, cpad, device="cuda", dtype=torch.float16)
    dyb[..., :Cout] = (torch.randn(B, Ho, Ho, Cout, device="cuda") / 8).half()
    xr = _nchw(x).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    F.conv2d(xr, wr, None, stride, pad).backward(_nchw(dyb[..., :Cout]))
    c32 = (Cout + 31) // 32 * 32
    wpad = torch.zeros(c32, Cin, k, k, device="cuda")
    wpad[:Cout] = w.float()
    dx = co.conv_dgrad(dyb[..., :c32], co.pack_dgrad(wpad), H, H, Cin, k, stride, pad)
    gref = xr.grad.permute(0, 2, 3, 1)
    err = (dx.float() - gref).abs().max().item()
    assert err <= 2e-3 * gref.abs().max().item() + 1e-4, "dgrad %g (max %g)" % (err, gref.abs().max().item())
    # accumulate form (what the engine uses for tensors with several consumers)
    co.conv_dgrad(dyb[..., :c32], co.pack_dgrad(wpad), H, H, Cin, k, stride, pad, out=dx, accumulate=True)
    assert (dx.float() - 2 * gref).abs().max().item() <= 4e-3 * gref.abs().max().item() + 2e-4
    del dx, gref
    acc = co.conv_wgrad(x, dyb, Cin, Cout, k, stride, pad)        # dY rows are 64-channel padded (TMA box width)
    gw = co.unpack_wgrad(acc, Cout, Cin, k)
    err = (gw - wr.grad).abs().max().item()
    assert err <= 1e-3 * wr.grad.abs().max().item() + 1e-5, "wgrad %g (max %g)" % (err, wr.grad.abs().max().item())
