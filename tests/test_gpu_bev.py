"""GPU: the LiDAR -> BEV rasteriser (csrc/bev.cu through the C-ABI and the drop-in data_process.kitti_bev_utils) against
the reference fixture and the CPU oracle.  Intensity and height are bit-exact (selection of
one input value / one float32 division); density is log(count+1)/log(64) in fp64 rounded to fp32: <= 1 ulp."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DENS_TOL = 6e-8


def _dense(g):
    m = np.zeros((3, 608 * 608))
    m[:, g["nz_cells"]] = g["nz_values"]
    return m.reshape(3, 608, 608)


def _check(got, ref):
    got = np.asarray(got, np.float64)
    assert got.shape == ref.shape
    assert np.array_equal(got[0], ref[0].astype(np.float32)) and np.array_equal(got[1], ref[1].astype(np.float32))
    assert np.abs(got[2] - ref[2]).max() <= DENS_TOL
    assert np.array_equal(got[2] != 0, ref[2] != 0)


def test_rasterize_vs_reference_golden(golden):
    from cy4 import bevops
    import data_process.kitti_bev_utils as kb
    g = golden("bev_raster.npz")
    ref = _dense(g)
    fused = bevops.rasterize([g["points"]], check=True)                      # removePoints fused into the kernel
    assert fused.shape == (1, 3, 608, 608) and fused.dtype == torch.float32 and fused.is_cuda
    _check(fused[0].cpu().numpy(), ref)
    b = kb.removePoints(g["points"].copy(), bevops.BOUNDARY)                 # the reference's two-call sequence
    rgb = kb.makeBVFeature(b, bevops.DISCRETIZATION, bevops.BOUNDARY)
    assert rgb.dtype == np.float64
    _check(rgb, ref)


def test_batch_vs_oracle():
    """KITTI-sized frames (120 k points), a batch with an empty frame, exact height ties and points on the bounds."""
    from cy4 import bevops, synth
    from oracle import bev_oracle as bo
    clouds = [synth.make_point_cloud(120000, seed=s) for s in (1, 2, 3)] + [np.zeros((0, 4), np.float32), synth.make_point_cloud(777, seed=9)]
    out = bevops.rasterize(clouds, check=True).cpu().numpy()
    assert out.shape == (5, 3, 608, 608)
    for i, c in enumerate(clouds):
        _check(out[i], bo.make_bv_feature(bo.remove_points(c)))
    assert not out[3].any()
    # same frames again: no state is carried between calls (workspace is re-zeroed)
    out2 = bevops.rasterize(clouds[:2]).cpu().numpy()
    assert np.array_equal(out2, out[:2])


def test_tie_order_and_unfiltered_input():
    from cy4 import bevops
    from oracle import bev_oracle as bo
    pts = np.array([[1.0, 0.0, 0.5, 0.11], [1.0, 0.0, 0.5, 0.22], [1.0, 0.0, 0.4, 0.33], [1.0, 0.0, -0.0, 0.5],
                    [50.0, 0.0, 0.0, 0.9], [10.0, 25.0, 0.0, 0.8], [0.0, -25.0, -2.73, 0.7]], np.float32)
    _check(bevops.rasterize([pts], check=True)[0].cpu().numpy(), bo.make_bv_feature(bo.remove_points(pts)))
    for order in ([1, 0, 2, 3, 4, 5, 6], [2, 1, 0, 3, 4, 5, 6]):              # file order decides among equal heights
        p = pts[order]
        _check(bevops.rasterize([p])[0].cpu().numpy(), bo.make_bv_feature(bo.remove_points(p)))
    # makeBVFeature contract: unfiltered points far outside the map are reported, not silently wrapped
    with pytest.raises(IndexError):
        bevops.rasterize([np.array([[-20.0, 0.0, 0.0, 1.0]], np.float32)], apply_filter=False, check=True)
