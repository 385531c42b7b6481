"""GPU: tcgen05 implicit-GEMM convolution (fprop, dgrad) through the C-ABI against a plain PyTorch
fp32 reference of the same op on the CPU (torch.nn.functional.conv2d), which is what the
reference's nn.Conv2d (src/models/darknet2pytorch.py:258-264) computes.

Tolerance: inputs are rounded to fp16 before BOTH implementations, so the only differences are the
fp32 accumulation order and the fp16 rounding of the stored output: |err| <= 2e-3 * max|y| is far
above that and far below any indexing / layout mistake (which gives O(1) errors)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

FWD_CASES = [
    # B, H, W, Cin, Cout, k, stride
    (2, 19, 19, 64, 64, 1, 1),      # 1x1, one k-block
    (2, 19, 19, 128, 256, 1, 1),    # 1x1, block_n 256
    (1, 38, 38, 64, 128, 3, 1),     # 3x3 pad 1 (im2col halo, image borders)
    (3, 19, 19, 256, 512, 3, 1),    # M=1083 (tail tile), two n tiles
    (2, 38, 38, 64, 128, 3, 2),     # stride 2
    (2, 40, 24, 32, 64, 3, 1),      # Cin=32: 64B swizzle path, non-square
    (2, 40, 24, 32, 64, 3, 2),
    (2, 19, 19, 128, 30, 1, 1),     # head: Cout=30 padded to 32
    (1, 76, 76, 128, 128, 3, 1),    # many tiles per CTA? (46 tiles) exercises pipeline wrap
    (4, 76, 76, 64, 64, 1, 1),      # 181 tiles > 148 SMs: persistent loop + TMEM double buffer
]


def _ref_conv(x_nhwc16, w16, stride, pad):
    x = x_nhwc16.float().permute(0, 3, 1, 2).cpu()
    return F.conv2d(x, w16.float().cpu(), None, stride, pad).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", FWD_CASES)
def test_conv_fwd(B, H, W, Cin, Cout, k, stride):
    from cy4 import convops as co
    torch.manual_seed(B * 1000 + H + Cin + Cout + k)
    pad = (k - 1) // 2
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
    wp = co.pack_fprop(w.float())
    y = co.conv_fwd(x, wp, Cout, k, stride, pad)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, stride, pad)
    got = y[..., :Cout].float().cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    if Cout % 32:
        assert float(y[..., Cout:].abs().max()) == 0.0


def test_conv_fwd_matrix_mode_and_stats():
    from cy4 import convops as co
    torch.manual_seed(1)
    B, H, W, Cin, Cout = 2, 38, 38, 64, 128
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda") / 8).half()
    wp = co.pack_fprop(w.float())
    s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
    y = co.conv_fwd(x, wp, Cout, 1, 1, 0, stats=(s1, s2), a_matrix=True)
    ref = _ref_conv(x, w, 1, 0)
    assert (y.float().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    np.testing.assert_allclose(s1.cpu().numpy(), ref.sum((0, 1, 2)).numpy(), rtol=1e-3, atol=2e-2)
    np.testing.assert_allclose(s2.cpu().numpy(), (ref * ref).sum((0, 1, 2)).numpy(), rtol=1e-3)


def test_conv_fwd_f32_bias_and_slices():
    """fp32 head output with bias; input read from / output written into channel slices (ld > C)."""
    from cy4 import convops as co
    torch.manual_seed(2)
    B, H, W = 2, 19, 19
    big = torch.randn(B, H, W, 192, device="cuda").half()
    x = big[..., 64:192]                              # Cin=128 slice, ld=192
    w = (torch.randn(30, 128, 1, 1, device="cuda") / 11).half()
    bias = torch.randn(32, device="cuda"); bias[30:] = 0
    y = co.conv_fwd(x, co.pack_fprop(w.float()), 30, 1, 1, 0, out_f32=True, bias=bias)
    ref = _ref_conv(x.contiguous(), w, 1, 0) + bias[:30].cpu()
    assert (y[..., :30].cpu() - ref).abs().max().item() < 2e-3
    w2 = (torch.randn(64, 128, 3, 3, device="cuda") / 34).half()
    outbuf = torch.zeros(B, H, W, 160, device="cuda", dtype=torch.float16)
    co.conv_fwd(x, co.pack_fprop(w2.float()), 64, 3, 1, 1, out=outbuf[..., 96:160])
    ref2 = _ref_conv(x.contiguous(), w2, 1, 1)
    assert (outbuf[..., 96:160].float().cpu() - ref2).abs().max().item() <= 2e-3 * ref2.abs().max().item() + 1e-3
    assert float(outbuf[..., :96].abs().max()) == 0.0


def test_stem_path():
    """3-channel stem: explicit im2col to [M,32] then the tensor-core GEMM (reference conv1)."""
    from cy4 import convops as co
    torch.manual_seed(3)
    for stride in (1, 2):
        x = torch.rand(2, 3, 64, 48, device="cuda")
        w = torch.randn(32, 3, 3, 3, device="cuda") / 5
        cols = co.stem_im2col(x, 3, stride, 1)
        # weights in (r, s, c) order padded to 32
        wp = torch.zeros(32, 32, device="cuda", dtype=torch.float16)
        wp[:, :27] = w.permute(0, 2, 3, 1).reshape(32, 27).half()
        y = co.conv_fwd(cols, wp, 32, 1, 1, 0, a_matrix=True)
        ref = F.conv2d(x.half().float().cpu(), w.half().float().cpu(), None, stride, 1).permute(0, 2, 3, 1)
        assert (y.float().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


DGRAD_CASES = [
    (2, 19, 19, 64, 128, 1, 1),
    (2, 38, 38, 64, 128, 3, 1),
    (1, 19, 19, 256, 512, 3, 1),
    (2, 38, 38, 64, 128, 3, 2),
    (2, 40, 24, 32, 64, 3, 2),
    (2, 76, 76, 128, 256, 3, 2),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", DGRAD_CASES)
def test_conv_dgrad(B, H, W, Cin, Cout, k, stride):
    from cy4 import convops as co
    torch.manual_seed(B + H + Cin + Cout + k + stride)
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    dy = torch.randn(B, Ho, Wo, Cout, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cout * k * k) ** 0.5).half()
    dx = co.conv_dgrad(dy, co.pack_dgrad(w.float()), H, W, Cin, k, stride, pad)
    x = torch.zeros(B, Cin, H, W, requires_grad=True)
    yref = F.conv2d(x, w.float().cpu(), None, stride, pad)
    yref.backward(dy.float().cpu().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1)
    err = (dx.float().cpu() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    # accumulate mode: dx2 = dx + dgrad
    dx2 = dx.clone()
    co.conv_dgrad(dy, co.pack_dgrad(w.float()), H, W, Cin, k, stride, pad, out=dx2, accumulate=True)
    assert (dx2.float().cpu() - 2 * ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 2e-3


WGRAD_CASES = [
    # B, H, W, Cin, Cout, k, stride
    (2, 19, 19, 64, 128, 1, 1),
    (2, 19, 19, 128, 64, 3, 1),      # Cout=64: second dY box out of range
    (1, 38, 38, 256, 256, 3, 1),     # N=256 (4 boxes), M two tiles
    (2, 38, 38, 64, 128, 3, 2),      # stride 2
    (3, 19, 19, 512, 256, 1, 1),     # two n tiles
    (2, 24, 40, 64, 64, 3, 1),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", WGRAD_CASES)
def test_conv_wgrad(B, H, W, Cin, Cout, k, stride):
    from cy4 import convops as co
    torch.manual_seed(B + H + Cin + Cout + k + stride + 5)
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    dy = (torch.randn(B, Ho, Wo, Cout, device="cuda") / (B * Ho * Wo) ** 0.5).half()
    acc = co.conv_wgrad(x, dy, Cin, Cout, k, stride, pad)
    gw = co.unpack_wgrad(acc, Cout, Cin, k)
    w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    y = F.conv2d(x.float().cpu().permute(0, 3, 1, 2), w, None, stride, pad)
    y.backward(dy.float().cpu().permute(0, 3, 1, 2))
    ref = w.grad
    err = (gw.cpu() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, err


def test_conv_pair_variant():
    """The CTA-pair (cta_group::2) kernel that serves every eligible launch by default against the 1-CTA kernel
    (cy4_set_option("conv_pair", 0); same K order: identical fp16 outputs) and the fp32 reference, for fprop (+BN statistics),
    stride-1 / stride-2 dgrad, odd m-tile counts."""
    from cy4 import _lib, convops as co
    L = _lib.lib()
    torch.manual_seed(77)
    for (B, H, W, Cin, Cout, k, stride) in [(2, 38, 38, 256, 512, 3, 1), (3, 19, 19, 512, 256, 1, 1), (2, 76, 76, 128, 128, 3, 1),
                                            (1, 38, 38, 256, 256, 3, 1), (2, 76, 76, 128, 256, 3, 2), (16, 38, 38, 256, 512, 3, 1)]:
        pad = (k - 1) // 2
        x = torch.randn(B, H, W, Cin, device="cuda").half()
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        dy = torch.randn(B, Ho, Wo, Cout, device="cuda").half()
        wp, wd = co.pack_fprop(w.float()), co.pack_dgrad(w.float())
        res = []
        for pair in (0, 1):
            _lib.check(L.cy4_set_option(b"conv_pair", pair))
            try:
                s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
                y = co.conv_fwd(x, wp, Cout, k, stride, pad, stats=(s1, s2))
                dx = co.conv_dgrad(dy, wd, H, W, Cin, k, stride, pad)
                torch.cuda.synchronize()
                res.append((y.clone(), dx.clone(), s1.clone(), s2.clone()))
            finally:
                _lib.check(L.cy4_set_option(b"conv_pair", 1))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        assert (res[0][2] - res[1][2]).abs().max().item() <= 1e-3 * res[0][2].abs().max().item() + 1e-3
        ref = _ref_conv(x, w, stride, pad)
        assert (res[1][0][..., :Cout].float().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_conv_wgrad_pair_variant():
    """The CTA-pair (cta_group::2) weight-gradient kernel that serves the eligible launches by default (Cout % 256 == 0, X tile
    of 128 / 256 channels) against the 1-CTA kernel (cy4_set_option("wgrad_pair", 0); fp32 split-K atomics: summation order
    only) and autograd."""
    from cy4 import _lib, convops as co
    L = _lib.lib()
    torch.manual_seed(41)
    for (B, H, W, Cin, Cout, k, stride) in [(2, 38, 38, 256, 512, 3, 1), (3, 19, 19, 512, 256, 1, 1), (4, 19, 19, 512, 1024, 3, 1),
                                            (2, 76, 76, 128, 256, 3, 2), (2, 38, 38, 128, 256, 3, 1), (32, 38, 38, 256, 512, 3, 1)]:
        pad = (k - 1) // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        x = torch.randn(B, H, W, Cin, device="cuda").half()
        dy = (torch.randn(B, Ho, Wo, Cout, device="cuda") / (B * Ho * Wo) ** 0.5).half()
        outs = []
        for pair in (0, 1):
            _lib.check(L.cy4_set_option(b"wgrad_pair", pair))
            try:
                outs.append(co.unpack_wgrad(co.conv_wgrad(x, dy, Cin, Cout, k, stride, pad), Cout, Cin, k))
                torch.cuda.synchronize()
            finally:
                _lib.check(L.cy4_set_option(b"wgrad_pair", 1))
        assert (outs[0] - outs[1]).abs().max().item() <= 1e-3 * outs[0].abs().max().item() + 1e-6, (B, H, Cin, Cout, k, stride)
        if B <= 4:
            w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
            F.conv2d(x.float().cpu().permute(0, 3, 1, 2), w, None, stride, pad).backward(dy.float().cpu().permute(0, 3, 1, 2))
            assert (outs[1].cpu() - w.grad).abs().max().item() <= 2e-3 * w.grad.abs().max().item() + 1e-4


def test_conv_wgrad_stem_and_narrow():
    """Stem cols matrix (32 wide, 64B-swizzle B operand) and 32-channel tensors stored with ld=64."""
    from cy4 import convops as co
    torch.manual_seed(9)
    xin = torch.rand(2, 3, 32, 48, device="cuda")
    cols = co.stem_im2col(xin, 3, 1, 1)                              # [2,32,48,32]
    dyb = torch.zeros(2, 32, 48, 64, device="cuda", dtype=torch.float16)
    dyb[..., :32] = (torch.randn(2, 32, 48, 32, device="cuda") / 40).half()
    acc = co.conv_wgrad(cols, dyb[..., :32], 32, 32, 1, 1, 0, a_matrix=True)
    ref = torch.einsum("bhwo,bhwi->oi", dyb[..., :32].float().cpu(), cols.float().cpu())
    assert (acc[:32, 0, :].cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-4
    # 32-channel activation kept in a 64-wide buffer (pad channels arbitrary but finite)
    xb = torch.randn(2, 20, 20, 64, device="cuda").half()
    dy2 = torch.zeros(2, 20, 20, 64, device="cuda", dtype=torch.float16)
    dy2[..., :64] = (torch.randn(2, 20, 20, 64, device="cuda") / 28).half()
    acc2 = co.conv_wgrad(xb[..., :32], dy2, 32, 64, 3, 1, 1)
    w = torch.zeros(64, 32, 3, 3, requires_grad=True)
    y = F.conv2d(xb[..., :32].float().cpu().permute(0, 3, 1, 2), w, None, 1, 1)
    y.backward(dy2.float().cpu().permute(0, 3, 1, 2))
    gw = co.unpack_wgrad(acc2, 64, 32, 3)
    assert (gw.cpu() - w.grad).abs().max().item() <= 2e-3 * w.grad.abs().max().item() + 1e-4


def test_conv_fp32_output_parity():
    """BASELINE.json: conv activations within 1e-3 of the fp32 reference on identical inputs.  With the
    fp32-output epilogue nothing but the accumulation order differs from F.conv2d."""
    from cy4 import convops as co
    torch.manual_seed(11)
    for (B, H, W, Cin, Cout, k, stride) in [(2, 38, 38, 256, 512, 3, 1), (2, 76, 76, 128, 128, 3, 1), (2, 38, 38, 512, 256, 1, 1),
                                            (2, 76, 76, 128, 256, 3, 2)]:
        pad = (k - 1) // 2
        x = torch.randn(B, H, W, Cin, device="cuda").half()
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
        y = co.conv_fwd(x, co.pack_fprop(w.float()), Cout, k, stride, pad, out_f32=True)
        ref = _ref_conv(x, w, stride, pad)
        assert (y.cpu() - ref).abs().max().item() <= 1e-3          # measured ~1e-5


def test_kblocks_per_slot():
    """Narrow layers pack up to 4 k-blocks into one pipeline slot (one barrier round trip per slot): bit-identical
    to one k-block per slot (same MMA order), checked on shapes with more tiles than SMs (3x3 s1/s2, 1x1, Cin=32,
    a k loop that is not a multiple of the packing, stride-2 dgrad parity classes)."""
    from cy4 import _lib, convops as co
    L = _lib.lib()
    torch.manual_seed(31)
    for (B, H, W, Cin, Cout, k, stride) in [(4, 152, 152, 64, 64, 3, 1), (4, 152, 152, 32, 64, 3, 2), (8, 152, 152, 64, 64, 1, 1),
                                             (4, 152, 152, 128, 64, 1, 1)]:
        pad = (k - 1) // 2
        x = torch.randn(B, H, W, Cin, device="cuda").half()
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).half()
        Ho = (H + 2 * pad - k) // stride + 1
        dy = torch.randn(B, Ho, Ho, Cout, device="cuda").half()
        wp, wd = co.pack_fprop(w.float()), co.pack_dgrad(w.float())
        outs = []
        for kps in (4, 1):
            _lib.check(L.cy4_set_option(b"kblocks_per_slot", kps))
            outs.append((co.conv_fwd(x, wp, Cout, k, stride, pad), co.conv_dgrad(dy, wd, H, W, Cin, k, stride, pad)))
        _lib.check(L.cy4_set_option(b"kblocks_per_slot", 4))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        ref = _ref_conv(x, w, stride, pad)
        assert (outs[0][0].float().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("option,value", [("conv_cluster", 2), ("conv_cluster", 4), ("wgrad_cluster", 2), ("tma_store", 0)])
def test_conv_variants(option, value):
    """The optional code paths (TMA multicast of the weight / activation slabs across thread-block
    clusters, direct-store epilogue) give the same results as the default configuration."""
    from cy4 import _lib, convops as co
    L = _lib.lib()
    torch.manual_seed(21)
    B, H, W, Cin, Cout = 2, 38, 38, 256, 512
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / 48).half()
    dy = (torch.randn(B, H, W, Cout, device="cuda") / 50).half()
    wp, wd = co.pack_fprop(w.float()), co.pack_dgrad(w.float())
    base = (co.conv_fwd(x, wp, Cout, 3, 1, 1), co.conv_dgrad(dy, wd, H, W, Cin, 3, 1, 1), co.conv_wgrad(x, dy, Cin, Cout, 3, 1, 1))
    default = {"conv_cluster": 1, "wgrad_cluster": 1, "tma_store": 1}[option]
    try:
        _lib.check(L.cy4_set_option(option.encode(), value))
        alt = (co.conv_fwd(x, wp, Cout, 3, 1, 1), co.conv_dgrad(dy, wd, H, W, Cin, 3, 1, 1), co.conv_wgrad(x, dy, Cin, Cout, 3, 1, 1))
    finally:
        _lib.check(L.cy4_set_option(option.encode(), default))
    assert torch.equal(base[0], alt[0]) and torch.equal(base[1], alt[1])
    assert (base[2] - alt[2]).abs().max().item() <= 1e-3 * base[2].abs().max().item()     # split-K atomics: order only


def test_batched_pack_and_unpack_match_the_single_layer_kernels():
    """cy4_pack_weights_batched / cy4_unpack_wgrad_batched (tiled through shared memory, one launch for all layers, optional
    BatchNorm fold scale) against the per-layer kernels and plain torch permutes."""
    import ctypes
    from cy4 import _lib, convops as co
    from cy4._sigs_engine import PackItem, UnpackItem
    L = _lib.lib()
    torch.manual_seed(5)
    shapes = [(64, 32, 3), (30, 256, 1), (128, 64, 1), (512, 256, 3), (96, 160, 3)]
    ws = [torch.randn(o, i, k, k, device="cuda") for o, i, k in shapes]
    scales = [torch.rand(o, device="cuda") + 0.5 if n % 2 else None for n, (o, i, k) in enumerate(shapes)]
    rup = lambda x, m: (x + m - 1) // m * m
    wf = [torch.full((rup(o, 32), k * k * i), 7.0, device="cuda", dtype=torch.float16) for o, i, k in shapes]
    wd = [torch.full((rup(i, 32), k * k * rup(o, 32)), 7.0, device="cuda", dtype=torch.float16) for o, i, k in shapes]
    items, tiles = [], 0
    for w, f, d, sc, (o, i, k) in zip(ws, wf, wd, scales, shapes):
        it = PackItem()
        it.tile_begin = tiles
        tiles += (rup(o, 32) // 32) * (rup(i, 32) // 32)
        it.w_oihw, it.w_fprop, it.w_dgrad = w.data_ptr(), f.data_ptr(), d.data_ptr()
        it.Cout, it.Cin, it.ksize, it.cout_pad, it.cin_pad = o, i, k, rup(o, 32), rup(i, 32)
        it.fold_scale = sc.data_ptr() if sc is not None else None
        items.append(it)
    arr = (PackItem * len(items))(*items)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().cuda()
    _lib.check(L.cy4_pack_weights_batched(table.data_ptr(), len(items), _lib.stream()))
    for w, f, d, sc, (o, i, k) in zip(ws, wf, wd, scales, shapes):
        wsc = w * sc[:, None, None, None] if sc is not None else w
        ref_f = torch.zeros(rup(o, 32), k * k * i, device="cuda")
        ref_f[:o] = wsc.permute(0, 2, 3, 1).reshape(o, -1)
        assert torch.equal(f, ref_f.half())
        ref_d = torch.zeros(rup(i, 32), k * k, rup(o, 32), device="cuda")
        ref_d[:i, :, :o] = w.permute(1, 2, 3, 0).reshape(i, k * k, o)
        assert torch.equal(d, ref_d.reshape(rup(i, 32), -1).half())
        if sc is None:
            assert torch.equal(f, co.pack_fprop(w)) and torch.equal(d[:, :], co.pack_dgrad(torch.cat([w, torch.zeros(rup(o, 32) - o, i, k, k, device="cuda")])))
    # unpack
    accs = [torch.randn(rup(o, 32), k * k, i, device="cuda") for o, i, k in shapes]
    gws = [torch.full((o, i, k, k), 3.0, device="cuda") for o, i, k in shapes]
    uitems, tiles = [], 0
    for a, g, (o, i, k) in zip(accs, gws, shapes):
        it = UnpackItem()
        it.tile_begin = tiles
        tiles += (rup(o, 32) // 32) * (rup(i, 32) // 32)
        it.dw_acc, it.gw_oihw, it.Cout, it.Cin, it.ksize = a.data_ptr(), g.data_ptr(), o, i, k
        uitems.append(it)
    uarr = (UnpackItem * len(uitems))(*uitems)
    utable = torch.frombuffer(bytearray(bytes(uarr)), dtype=torch.uint8).clone().cuda()
    dscale = torch.tensor([0.25], device="cuda")
    _lib.check(L.cy4_unpack_wgrad_batched(utable.data_ptr(), len(uitems), dscale.data_ptr(), _lib.stream()))
    for a, g, (o, i, k) in zip(accs, gws, shapes):
        assert torch.equal(g, 0.25 * a[:o].reshape(o, k, k, i).permute(0, 3, 1, 2))


def test_shifted_bn_statistics_have_no_cancellation():
    """cy4_conv_fwd_stats: the BatchNorm sums taken about a per-channel shift c.  The statistics are those of the STORED fp16
    tensor (the epilogue reads them off the staged output slab: what BatchNorm then normalises).  On a layer whose outputs have
    |mean| ~ 1000 sigma the plain E[y^2] - E[y]^2 in fp32 loses the variance entirely; with c = a previous estimate of the mean
    it is exact to fp32 rounding.  (The engine passes last step's batch mean as c.)  In a regime where fp16 storage resolves
    sigma (|mean| ~ 25 sigma) the same statistics also match those of the fp32 convolution."""
    import ctypes
    from cy4 import _lib, convops as co
    from cy4._sigs_engine import CONV_STATS
    L = _lib.lib()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(3)
        B, H, Cin, Cout = 8, 76, 64, 64
        n = B * H * H
        w = ((1.0 + 0.2 * torch.rand(Cout, Cin, 1, 1, device="cuda")) / Cin).half()
        wp = co.pack_fprop(w.float())
        for regime, (mu, sd) in (("extreme", (8.0, 0.05)), ("moderate", (2.0, 0.64))):
            x = (mu + sd * torch.randn(B, H, H, Cin, device="cuda")).half()
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float()).double()
            mean_ref, var_ref = ref.mean((0, 2, 3)), ref.var((0, 2, 3), unbiased=False)
            ratio = (mean_ref.abs() / var_ref.sqrt()).min().item()
            assert ratio > (500 if regime == "extreme" else 15)               # the regime under test
            y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
            out = {}
            for tag, shift in (("plain", None), ("shifted", (mean_ref + 0.01).float().contiguous())):
                s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
                d = co.conv_desc(B, H, H, Cin, Cout, 1, 1, 0, Cin, Cout, CONV_STATS)
                _lib.check(L.cy4_conv_fwd_stats(ctypes.byref(d), x.data_ptr(), wp.data_ptr(), y.data_ptr(), s1.data_ptr(), s2.data_ptr(),
                                                shift.data_ptr() if shift is not None else None, _lib.stream()))
                yd = y.double()
                mean_y, var_y = yd.mean((0, 1, 2)), yd.var((0, 1, 2), unbiased=False)      # the stored tensor's own statistics
                ms = s1.double() / n
                var = s2.double() / n - ms * ms
                mean = ms + (shift.double() if shift is not None else 0.0)
                out[tag] = ((mean - mean_y).abs().max().item(), ((var - var_y).abs() / var_y).max().item(),
                            ((var - var_ref).abs() / var_ref).max().item())
            print(regime, "|mean|/sigma >= %.0f" % ratio, " plain (mean err, var rel err vs stored, vs fp32 conv):", out["plain"], " shifted:", out["shifted"])
            assert out["shifted"][0] <= 1e-4 and out["shifted"][1] <= 1e-4
            if regime == "extreme":
                assert out["plain"][1] > 10 * out["shifted"][1]                  # what the shift buys
            else:
                assert out["shifted"][2] <= 1e-3                                 # fp16 storage resolves sigma: same as the fp32 conv's
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
