#!/usr/bin/env python
"""A DeblurGS-style training loop on the deblurgs_b200 API, end to end on synthetic data (needs a B200).

What the reference's train.py does per iteration (train.py:120-215), with this repository's pieces:

  reference                                         here
  ------------------------------------------------  ---------------------------------------------------------
  CameraMotionModule(cam_infos, args)               CameraMotionModule.from_poses(cams, R, t, ...)   (se3_log_map init)
  gaussians.training_setup(opt)                     GaussianParams.training_setup(...)                (FusedAdam, 1 launch/step)
  camera_motion_module.add_training_setup(...)      same name
  query(cam_idx) -> F x render(), stack, mean       query(cam_idx): ONE batched forward for the F sub-frames
  l1_loss + lambda * batchwise_smoothness_loss      blur_photometric_loss (one kernel each way)
  loss.backward()                                   one batched backward (+ pose kernel backward)
  for pkg in render_pkgs: add_densification_stats   add_densification_stats_blurry(pkg)  (computed in the backward)
  clip_grad_value_; optimizer.step(); zero_grad     optimizer.step(clip_grad_value=...); zero_grad
  gaussians.densify_and_prune(threshold, extent)    same name (deblurgs_b200/densify.py; optimizer state follows)

The "ground truth" blurry images are rendered from a hidden scene / trajectories; training starts from perturbed
colours, opacities and trajectories and must bring the photometric loss down.

  python examples/train_blurry_views.py [--iters 200] [--views 4] [--gaussians 20000]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deblurgs_b200 import blur_photometric_loss, synthetic  # noqa: E402
from deblurgs_b200.motion import CameraMotionModule, GaussianParams  # noqa: E402
from deblurgs_b200.pose import bezier_se3_poses  # noqa: E402


class RefCamera:
    """The fields CameraMotionModule / render read from the reference's Camera objects."""

    def __init__(self, cam, device):
        self.image_width, self.image_height = cam.width, cam.height
        self.FoVx, self.FoVy, self.znear, self.zfar = cam.fovx, cam.fovy, cam.znear, cam.zfar
        self.projection_matrix = cam.projection_matrix_t().to(device)
        self.original_image = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--gaussians", type=int, default=20000)
    ap.add_argument("--subframes", type=int, default=8)
    ap.add_argument("--width", type=int, default=200)
    ap.add_argument("--height", type=int, default=136)
    ap.add_argument("--densify-every", type=int, default=100, help="0: never")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    order, F = 3, args.subframes

    cam = synthetic.make_camera(args.width, args.height)
    scene = synthetic.make_scene(args.gaussians, cam, seed=0).to(dev)
    bg = torch.zeros(3, device=dev)

    # ---- hidden truth: one trajectory per view around the base pose; blurry observations rendered from it
    truth = GaussianParams.from_scene(scene)
    cams = [RefCamera(cam, dev) for _ in range(args.views)]
    true_traj = [synthetic.make_trajectory(F, order, seed=10 + i).to(dev) for i in range(args.views)]
    base = torch.tensor([synthetic.BASE_SE3], dtype=torch.float32, device=dev).repeat(args.views, 1)
    oracle_cmm = CameraMotionModule(cams, base, curve_order=order, num_subframes=F)
    with torch.no_grad():
        for i, t in enumerate(true_traj):
            oracle_cmm._trans._control_points[i].copy_(t.ctrl_trans)
            oracle_cmm._rot._control_points[i].copy_(t.ctrl_rot)
    oracle_cmm.link_gaussian(truth)
    with torch.no_grad():
        for i in range(args.views):
            cams[i].original_image = oracle_cmm.query(i, "all", background=bg)["blurred"].clone()

    # ---- the model being trained: same geometry, perturbed appearance; trajectories initialised from the mid pose
    g = torch.Generator().manual_seed(123)
    gauss = GaussianParams.from_scene(scene)
    with torch.no_grad():
        gauss._features_dc.add_(0.5 * torch.randn(gauss._features_dc.shape, generator=g).to(dev))
        gauss._features_rest.mul_(0.0)
        gauss._opacity.mul_(0.7)
    opt = gauss.training_setup(position_lr_init=1.6e-5, feature_lr=0.01, opacity_lr=0.02, scaling_lr=0.002,
                               rotation_lr=0.001)
    # mid-trajectory camera pose of every view -> (c2w rotation, camera position), as COLMAP would give them
    mid_R, mid_t = [], []
    for t in true_traj:
        view, _, campos = bezier_se3_poses(t.ctrl_trans, t.ctrl_rot, torch.tensor([0.5], device=dev),
                                           cams[0].projection_matrix)
        mid_R.append(view[0, :3, :3])
        mid_t.append(campos[0])
    cmm = CameraMotionModule.from_poses(cams, torch.stack(mid_R), torch.stack(mid_t), curve_order=order,
                                        num_subframes=F)
    cmm.link_gaussian(gauss)
    cmm.add_training_setup(gauss, {"curve_rot": 2e-4, "curve_trans": 2e-3, "curve_alignment": 0.0})

    lam0, lam1 = 1e-3, 1e-5     # lambda_t_smooth_init / final (arguments/__init__.py:97-98)
    log, t0 = [], time.perf_counter()
    for it in range(args.iters):
        idx = it % args.views
        lam = lam0 * (lam1 / lam0) ** (it / max(args.iters - 1, 1))
        out = cmm.query(idx, "all", background=bg)
        loss = blur_photometric_loss(out["blurred"], out["subframes"], out["gt"], lam)
        loss.backward()
        with torch.no_grad():
            gauss.add_densification_stats_blurry(out["batched"])    # train.py:188-193, from the backward's epilogue
        opt.step(clip_grad_value=1.0)
        opt.zero_grad(set_to_none=True)
        if args.densify_every and it > 0 and it % args.densify_every == 0 and it < args.iters - 1:
            n_before = gauss.get_xyz.shape[0]
            gauss.densify_and_prune(max_grad=2e-4, extent=5.0)       # train.py:195-196
            print("iter %4d  densify/prune: %d -> %d Gaussians" % (it, n_before, gauss.get_xyz.shape[0]), flush=True)
        if it % max(args.iters // 10, 1) == 0 or it == args.iters - 1:
            log.append((it, float(loss.detach())))
            print("iter %4d  loss %.5f" % log[-1], flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d iterations in %.2f s (%.2f ms / blurry view incl. optimizer); loss %.5f -> %.5f; "
          "densification: %d Gaussians seen, max screen radius %d px"
          % (args.iters, dt, 1e3 * dt / args.iters, log[0][1], log[-1][1], int((gauss.denom > 0).sum()),
             int(gauss.max_radii2D.max())))
    assert log[-1][1] < 0.8 * log[0][1], "training did not reduce the loss"


if __name__ == "__main__":
    main()
