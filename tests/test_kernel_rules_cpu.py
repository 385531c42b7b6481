"""CPU checks of two index rules that CUDA kernels rely on (numpy mirrors of the device code; the kernels themselves are checked
against torch / the oracle in the -m gpu tests):

* separable arg-max pooling (csrc/elementwise.cu maxpool_v_kernel + maxpool_h_kernel): a column pass that keeps the column
  maximum and the SMALLEST dy attaining it, then a row pass with the tie-break (greater value, else smaller dy, else smaller dx),
  must pick exactly torch's first maximum of the row-major window scan -- including on inputs full of ties;
* the work-unit order of the merged stride-2 input gradient (csrc/conv_tc.cuh unit_decode): every (tile, parity class) pair
  exactly once, and every CTA of the static round-robin sees a balanced mix of the 1 / 2 / 2 / 4-tap classes."""
import collections

import numpy as np
import torch
import torch.nn.functional as F


def _separable_argmax_pool(x, k):
    """x [H, W] -> (out [H, W], window offset dy*k+dx [H, W]) for a stride-1 'same' max pool, the device algorithm."""
    H, W = x.shape
    pad = k // 2
    vmax = np.full((H, W), -np.inf, np.float32); vdy = np.full((H, W), -1, np.int64)
    for oh in range(H):
        for w in range(W):
            for dy in range(k):
                h = oh - pad + dy
                if 0 <= h < H and (x[h, w] > vmax[oh, w] or vdy[oh, w] < 0):
                    vmax[oh, w] = x[h, w]; vdy[oh, w] = dy
    out = np.empty((H, W), np.float32); pos = np.empty((H, W), np.int64)
    for oh in range(H):
        for ow in range(W):
            best, bdy, bdx = -np.inf, 1 << 20, -1
            for dx in range(k):
                w = ow - pad + dx
                if not 0 <= w < W:
                    continue
                t, dy = vmax[oh, w], vdy[oh, w]
                if bdx < 0 or t > best or (t == best and dy < bdy):
                    best, bdy, bdx = t, dy, dx
            out[oh, ow] = best; pos[oh, ow] = bdy * k + bdx
    return out, pos


def test_separable_argmax_pool_picks_torchs_first_maximum():
    rng = np.random.default_rng(0)
    for k in (5, 9, 13):
        for coarse in (False, True):
            x = rng.standard_normal((19, 19)).astype(np.float32)
            if coarse:
                x = np.round(x * 2) / 2                      # many tied maxima per window
            out, pos = _separable_argmax_pool(x, k)
            ref, idx = F.max_pool2d(torch.tensor(x)[None, None], k, 1, k // 2, return_indices=True)
            assert np.array_equal(out, ref[0, 0].numpy())
            pad = k // 2
            oh, ow = np.meshgrid(np.arange(19), np.arange(19), indexing="ij")
            flat = (oh - pad + pos // k) * 19 + (ow - pad + pos % k)      # window offset -> input position
            assert np.array_equal(flat, idx[0, 0].numpy()), (k, coarse)


def _unit_decode(t, ncls):
    tt = t // ncls
    return (t - tt * ncls + tt) % ncls, tt


def test_interleaved_parity_class_order_is_a_bijection_and_balanced():
    taps = [1, 2, 2, 4]                                           # taps of the four output-parity classes of a 3x3 / stride-2 dgrad
    for cls_units, grid in ((5776, 148), (722, 74), (1444, 74), (2888, 148), (181, 74)):
        seen = set(); load = collections.Counter()
        for t in range(cls_units * 4):
            cls, tt = _unit_decode(t, 4)
            assert 0 <= cls < 4 and 0 <= tt < cls_units
            seen.add((cls, tt)); load[t % grid] += taps[cls]
        assert len(seen) == cls_units * 4                         # every (class, tile) exactly once
        if cls_units >= 722:                                      # the launches the host enables it for (dY larger than L2 keeps)
            mean = sum(load.values()) / grid
            assert max(load.values()) <= 1.03 * mean, (cls_units, grid, max(load.values()) / mean)
