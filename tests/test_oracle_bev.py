"""CPU: the BEV rasteriser oracle against the fixture made by the unmodified reference's data_process/kitti_bev_utils.py
(tests/golden/bev_raster.npz, oracle/gen_golden.py gen_bev)."""
import numpy as np


def _dense(g):
    m = np.zeros((3, 608 * 608))
    m[:, g["nz_cells"]] = g["nz_values"]
    return m.reshape(3, 608, 608)


def test_rasteriser_oracle_vs_reference(golden):
    from oracle import bev_oracle as bo
    g = golden("bev_raster.npz")
    b = bo.remove_points(g["points"])
    assert b.shape[0] == int(g["filtered_rows"]) and np.array_equal(b[:64], g["filtered_head"])
    assert float(g["discretization"]) == bo.DISCRETIZATION
    rgb = bo.make_bv_feature(b)
    assert rgb.dtype == np.float64 and np.array_equal(rgb, _dense(g))           # bit for bit


def test_build_yolo_target_vs_reference(golden):
    from oracle import bev_oracle as bo
    import data_process.kitti_bev_utils as kb          # the drop-in's host half needs no GPU
    g = golden("bev_raster.npz")
    assert np.array_equal(bo.build_yolo_target(g["labels"]), g["yolo_target"])
    assert np.array_equal(kb.build_yolo_target(g["labels"]), g["yolo_target"])
    b = kb.removePoints(g["points"].copy(), bo.BOUNDARY)
    assert b.shape[0] == int(g["filtered_rows"]) and np.array_equal(b[:64], g["filtered_head"])


def test_tie_and_edge_semantics():
    """Equal heights in a cell: the first point in file order supplies the intensity; bounds are inclusive; row / column
    608 of the scratch map are cropped."""
    from oracle import bev_oracle as bo
    d = bo.DISCRETIZATION
    pts = np.array([[1.0, 0.0, 0.5, 0.11], [1.0, 0.0, 0.5, 0.22], [1.0, 0.0, 0.4, 0.33],          # one cell, tie on z
                    [50.0, 0.0, 0.0, 0.9], [10.0, 25.0, 0.0, 0.8], [0.0, -25.0, -2.73, 0.7]], np.float32)
    rgb = bo.make_bv_feature(bo.remove_points(pts))
    ix, iy = int(np.floor(np.float32(1.0) / np.float32(d))), 304
    assert rgb[0, ix, iy] == np.float32(0.11) and rgb[2, ix, iy] == min(1.0, np.log(4) / np.log(64))
    assert rgb[0, 0, 0] == np.float32(0.7) and rgb[1, 0, 0] == 0.0
    assert (rgb[0] != 0).sum() == 2                 # the x = 50 and y = 25 points land in the cropped row / column
