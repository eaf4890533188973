"""CPU: the PyTorch-CPU evaluation of the blurry-view path (oracle/raster_torch_cpu.py, bench.py's cpu_baseline
leg) against the numpy restatement of the reference rasterizer (oracle/raster_np.py, itself pinned by golden vectors
of the reference CUDA extension): same lists, same images, and its autograd gradients equal the oracle's float64
backward where the reference's backward is the true derivative (Gaussian parameters)."""
import numpy as np
import torch

from deblurgs_b200 import synthetic
from oracle import pose_torch as pt, raster_np as rn, raster_torch_cpu as rt


def test_torch_cpu_path_matches_numpy_oracle():
    P, W, H, F, order = synthetic.CONFIGS["tiny"]
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(P, cam)
    traj = synthetic.make_trajectory(F, order)
    bg, gt = synthetic.make_background(), synthetic.make_target(cam)
    params = [t.clone() for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)]
    ct, cr = traj.ctrl_trans.clone(), traj.ctrl_rot.clone()
    loss, blurred = rt.blurry_view_step(params, ct, cr, traj.nu, cam.projection_matrix_t(), bg, gt, W, H,
                                        cam.tanfovx, cam.tanfovy)
    # numpy oracle on the same poses
    poses = pt.trajectory(traj.ctrl_trans, traj.ctrl_rot, traj.nu, cam.projection_matrix_t())
    a = [t.numpy() for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)]
    fws = [rn.forward(a[0], a[1], a[2], a[3], a[4], 3, v.numpy(), p.numpy(), c.numpy(), bg.numpy(), W, H,
                      cam.tanfovx, cam.tanfovy) for (v, p, c) in poses]
    ref_blur = np.mean(np.stack([f["color"] for f in fws]), axis=0)
    err = np.abs(blurred.numpy() - ref_blur)
    assert np.quantile(err, 0.999) < 1e-4 and err.max() < 2e-2
    # lists: same Gaussians per tile in the same order (first sub-frame)
    with torch.no_grad():
        v, p, c = poses[0]
        pre = rt.project(*params, 3, v.float(), p.float(), c.float(), W, H, cam.tanfovx, cam.tanfovy)
        pl, bounds = rt.bin_tiles(pre)
    assert np.array_equal(pl.numpy().astype(np.uint32), fws[0]["point_list"])
    assert np.array_equal(np.stack([bounds[:-1].numpy(), bounds[1:].numpy()], 1)[fws[0]["ranges"][:, 1] > fws[0]["ranges"][:, 0]],
                          fws[0]["ranges"][fws[0]["ranges"][:, 1] > fws[0]["ranges"][:, 0]])
    # gradients of the L1 loss: autograd vs the oracle's float64 backward
    dblur = np.sign(blurred.numpy() - gt.numpy()) / blurred.numel()
    acc = {}
    for s, (v, p, c) in enumerate(poses):
        bw = rn.backward(fws[s], a[0], a[1], a[2], a[4], 3, v.numpy(), p.numpy(), c.numpy(), bg.numpy(), W, H,
                         cam.tanfovx, cam.tanfovy, dblur / F, np.zeros((1, H, W)))
        for k in ("dL_dmeans3D", "dL_dopacity", "dL_dsh", "dL_dscales", "dL_drotations"):
            acc[k] = acc.get(k, 0) + bw[k]
    for k, t in zip(("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh"), params):
        g, r = t.grad.double().numpy().reshape(acc[k].shape), acc[k]
        assert np.abs(g - r).max() <= 5e-3 * np.abs(r).max(), (k, np.abs(g - r).max() / np.abs(r).max())
    assert ct.grad is not None and torch.isfinite(ct.grad).all() and float(ct.grad.abs().max()) > 0
    assert cr.grad is not None and torch.isfinite(cr.grad).all()
