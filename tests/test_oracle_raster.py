"""CPU: the numpy restatement of the reference rasterizer (oracle/raster_np.py) against golden vectors
produced by the reference's own CUDA extension on a B200 (tests/golden/make_raster_golden.py)."""
import os

import numpy as np
import pytest
import torch

from deblurgs_b200 import synthetic
from oracle import raster_np as rn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(tag):
    return np.load(os.path.join(GOLD, "raster_golden_%s.npz" % tag))


def _scene():
    P, W, H, F, order = synthetic.CONFIGS["tiny"]
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(P, cam)
    return cam, [t.numpy() for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)], F


@pytest.mark.parametrize("tag,use_sigmoid", [("tiny", False), ("tiny_sigmoid", True)])
def test_oracle_forward_matches_reference_cuda(tag, use_sigmoid):
    g = _load(tag)
    cam, a, F = _scene()
    W, H = cam.width, cam.height
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    for s in range(F):
        fw = rn.forward(a[0], a[1], a[2], a[3], a[4], 3, g["view"][s], g["proj"][s], g["campos"][s], g["bg"], W, H,
                        cam.tanfovx, cam.tanfovy, use_sigmoid=use_sigmoid)
        pre = fw["pre"]
        k = "s%d_" % s
        vis = g[k + "radii"] > 0
        # integer / index work: bit-exact
        assert np.array_equal(pre["radii"], g[k + "radii"])
        assert np.array_equal(pre["tiles_touched"], g[k + "tiles_touched"])
        assert np.array_equal(pre["depths"].view(np.uint32)[vis], g[k + "depths"].view(np.uint32)[vis])
        assert np.array_equal(pre["means2D"].view(np.uint32)[vis], g[k + "means2D"].reshape(-1, 2).view(np.uint32)[vis])
        assert np.array_equal(pre["cov3D"].view(np.uint32)[vis], g[k + "cov3D"].reshape(-1, 6).view(np.uint32)[vis])
        co = g[k + "conic_opacity"].reshape(-1, 4)
        assert np.array_equal(pre["conic"].view(np.uint32)[vis], co[:, :3].view(np.uint32)[vis])
        assert np.array_equal(fw["keys"], g[k + "keys"].view(np.uint64))
        assert np.array_equal(fw["point_list"], g[k + "point_list"].view(np.uint32))
        assert fw["point_list"].size == int(g[k + "num_rendered"])
        assert np.array_equal(fw["ranges"], g[k + "ranges"].reshape(tiles, 2))
        # floating point: colours to 1e-6; images to 1e-4 except isolated alpha-threshold flips
        # (numpy exp != CUDA expf in the last bit)
        np.testing.assert_allclose(pre["rgb"][vis], g[k + "rgb"].reshape(-1, 3)[vis], atol=1e-6)
        err = np.abs(fw["color"] - g[k + "color"])
        assert np.quantile(err, 0.999) < 1e-4 and err.max() < 2e-2
        errd = np.abs(fw["depth"] - g[k + "depth"])
        assert np.quantile(errd, 0.999) < 1e-3
        nc = g[k + "n_contrib"].reshape(H, W)
        assert (fw["n_contrib"] != nc).mean() < 2e-3
        assert np.quantile(np.abs(fw["final_T"] - g[k + "final_T"].reshape(H, W)), 0.999) < 1e-5


@pytest.mark.parametrize("tag,use_sigmoid", [("tiny", False), ("tiny_sigmoid", True)])
def test_oracle_backward_matches_reference_cuda(tag, use_sigmoid):
    g = _load(tag)
    cam, a, F = _scene()
    W, H = cam.width, cam.height

    def rel(x, y):
        return np.abs(x - y).max() / max(np.abs(y).max(), 1e-30)

    for s in range(F):
        k = "s%d_" % s
        fw = rn.forward(a[0], a[1], a[2], a[3], a[4], 3, g["view"][s], g["proj"][s], g["campos"][s], g["bg"], W, H,
                        cam.tanfovx, cam.tanfovy, use_sigmoid=use_sigmoid)
        bw = rn.backward(fw, a[0], a[1], a[2], a[4], 3, g["view"][s], g["proj"][s], g["campos"][s], g["bg"], W, H,
                         cam.tanfovx, cam.tanfovy, g["dL_dpix"][s], g["dL_ddepth"][s], use_sigmoid=use_sigmoid)
        # tolerance: the reference accumulates with fp32 atomics in arbitrary order (ref-vs-ref noise ~1e-4)
        # and single-pixel alpha flips perturb a few Gaussians
        for name in ["dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dviewmatrix",
                     "dL_dprojmatrix", "dL_dcov3D"]:
            r = rel(bw[name], g[k + name].astype(np.float64).reshape(bw[name].shape))
            assert r < 2e-3, (name, s, r)
        r = rel(bw["dL_dmeans2D"], g[k + "dL_dmeans2D"][:, :2].astype(np.float64))
        assert r < 2e-3, ("dL_dmeans2D", r)
        # projection-matrix quirk: entries 3,7,11,15 all carry the same value (backward.cu:436-448)
        dp = bw["dL_dprojmatrix"].reshape(16)
        assert dp[3] == dp[7] == dp[11] == dp[15]
        # view-matrix quirk: last column untouched (backward.cu:279-293, 454-457)
        dv = bw["dL_dviewmatrix"].reshape(16)
        assert dv[3] == dv[7] == dv[11] == dv[15] == 0


def test_oracle_edge_cases():
    cam = synthetic.make_camera(40, 24)       # partial edge tiles
    sc = synthetic.make_scene(50, cam)
    a = [t.numpy() for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)]
    view = np.eye(4, dtype=np.float32).reshape(16)
    proj = cam.projection_matrix_t().numpy().reshape(16)
    bg = np.array([0.1, 0.5, 0.9], np.float32)
    # everything behind the camera: num_rendered = 0 -> background image, depth = z_far, T = 1, n_contrib = 0
    behind = a[0].copy()
    behind[:, 2] = -np.abs(behind[:, 2]) - 1
    fw = rn.forward(behind, a[1], a[2], a[3], a[4], 3, view, proj, np.zeros(3, np.float32), bg, 40, 24, cam.tanfovx,
                    cam.tanfovy)
    assert fw["point_list"].size == 0 and (fw["pre"]["radii"] == 0).all()
    assert np.allclose(fw["color"], bg[:, None, None]) and np.allclose(fw["depth"], 100.0)
    assert (fw["final_T"] == 1).all() and (fw["n_contrib"] == 0).all()
