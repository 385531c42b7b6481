"""CPU: the host-side tiling / split-K choice of the weight-gradient kernel (cy4_conv_wgrad_plan, no launch) over every
conv layer of complex_yolov4 at the BASELINE batch.  One 193 KB CTA fits per SM, so a grid runs in waves of 148 CTAs:
the chosen split must not leave a nearly empty trailing wave, and must never cost more than the earlier
ceil(2 x 148 / items) rule under the kernel's cost model, waves x (k-blocks per CTA + 6 for the fixed prologue/epilogue)."""
import ctypes
import math

SMS = 148       # sm_count() without a device


def _convs(batch=32):
    from cy4 import netdefs
    H = W = 608
    C = 3
    outs, shapes = [], []
    ind = -1
    for b in netdefs.NETS["complex_yolov4"]():
        t = b["type"]
        if t == "net":
            continue
        ind += 1
        if t == "convolutional":
            k, s = int(b["size"]), int(b["stride"])
            pad = (k - 1) // 2 if int(b["pad"]) else 0
            Hi, Wi = H, W
            H, W = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
            shapes.append((batch, Hi, Wi, C, H, W, int(b["filters"]), k, s, pad))
            C = int(b["filters"])
        elif t == "upsample":
            H, W = H * 2, W * 2
        elif t == "route":
            ls = [int(i) if int(i) > 0 else int(i) + ind for i in b["layers"].split(",")]
            H, W, C = outs[ls[0]][0], outs[ls[0]][1], sum(outs[l][2] for l in ls)
        outs.append((H, W, C))
    return shapes


def test_split_k_avoids_empty_waves():
    from cy4 import _lib, _sigs_engine as se
    L = _lib.lib()
    shapes = [s for s in _convs() if s[3] % 32 == 0]          # the 3-channel stem goes through the im2col matrix path
    assert len(shapes) == 109
    total_new = total_old = 0
    for (B, Hi, Wi, Cin, Ho, Wo, Cout, k, s, pad) in shapes:
        d = se.ConvDesc(B, Hi, Wi, Cin, Ho, Wo, Cout, k, s, pad, (Cin + 31) // 32 * 32, (Cout + 63) // 64 * 64, 0, 0)
        out = (ctypes.c_int32 * 8)()
        assert L.cy4_conv_wgrad_plan(ctypes.byref(d), out) == 0
        m_tiles, n_tiles, block_n, tpc, tap_groups, kblocks, ksplit, grid = list(out)
        items = m_tiles * n_tiles * tap_groups
        assert grid == items * ksplit and 1 <= ksplit <= kblocks
        assert m_tiles == (Cout + 127) // 128 and kblocks == (B * Ho * Wo + 127) // 128
        assert tpc * block_n <= 256 and tap_groups * tpc >= k * k            # TMEM columns / all taps covered
        waves = math.ceil(grid / SMS)
        cost = waves * (math.ceil(kblocks / ksplit) + 6)
        old_ks = max(1, min(kblocks, (2 * SMS + items - 1) // items))
        old_cost = math.ceil(items * old_ks / SMS) * (math.ceil(kblocks / old_ks) + 6)
        assert cost <= old_cost, (Cin, Cout, k, Ho, cost, old_cost)
        # the last wave is at least 60 % full whenever the split is free to choose
        if items <= SMS:
            assert grid - (waves - 1) * SMS >= 0.6 * SMS, (Cin, Cout, k, Ho, grid)
        total_new += cost; total_old += old_cost
    # model: k-block-times summed over the net drop by > 15 % against the old rule
    assert total_new < 0.85 * total_old, (total_new, total_old)
