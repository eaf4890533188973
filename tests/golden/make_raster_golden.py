"""Generates tests/golden/raster_golden_<config>.npz on a GPU box by running the REFERENCE's own CUDA
rasterizer (oracle/_ref/libref_dgr.so = the reference sources compiled in place + our C shim) on the
seeded synthetic inputs of deblurgs_b200/synthetic.py.

Run (GPU box):  python tests/golden/make_raster_golden.py tiny   -> gpurun_out/raster_golden_tiny.npz
then copy the file into tests/golden/. The sub-frame matrices are stored too (they come from the pose
kernel), so that the fixture pins the rasterizer independently of the pose generator.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import parity_utils as pu  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def main(name, use_sigmoid=False):
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs(name)
    P, F, W, H = scene.means3D.shape[0], view.shape[0], cam.width, cam.height
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    g = torch.Generator().manual_seed(7)
    dL_dpix = (torch.randn(F, 3, H, W, generator=g) / (3 * H * W)).cuda()
    dL_ddepth = (torch.randn(F, 1, H, W, generator=g) / (H * W) * 0.1).cuda()
    out = dict(view=view.cpu().numpy(), proj=proj.cpu().numpy(), campos=campos.cpu().numpy(), bg=bg.cpu().numpy(),
               dL_dpix=dL_dpix.cpu().numpy(), dL_ddepth=dL_ddepth.cpu().numpy(), use_sigmoid=np.array(use_sigmoid))
    for s in range(F):
        r = ref_cuda.forward(scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None,
                             view[s].contiguous(), proj[s].contiguous(), campos[s].contiguous(), bg, W, H,
                             cam.tanfovx, cam.tanfovy, 3, use_sigmoid=use_sigmoid)
        b = ref_cuda.backward(r, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None,
                              view[s].contiguous(), proj[s].contiguous(), campos[s].contiguous(), bg, W, H,
                              cam.tanfovx, cam.tanfovy, 3, dL_dpix[s].contiguous(), dL_ddepth[s].contiguous(),
                              use_sigmoid=use_sigmoid)
        pre = "s%d_" % s
        out[pre + "num_rendered"] = np.array(r["num_rendered"])
        out[pre + "radii"] = r["radii"].cpu().numpy()
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched", "pre_sigmoid"):
            out[pre + k] = r["geom"][k].cpu().numpy()
        out[pre + "keys"] = r["binning"]["point_list_keys"].cpu().numpy()
        out[pre + "point_list"] = r["binning"]["point_list"].cpu().numpy()
        out[pre + "ranges"] = r["image"]["ranges"][:2 * tiles].cpu().numpy()
        out[pre + "final_T"] = r["image"]["accum_alpha"].cpu().numpy()
        out[pre + "n_contrib"] = r["image"]["n_contrib"].cpu().numpy()
        out[pre + "color"] = r["color"].cpu().numpy()
        out[pre + "depth"] = r["depth"].cpu().numpy()
        for k, v in b.items():
            out[pre + k] = v.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "raster_golden_%s%s.npz" % (name, "_sigmoid" if use_sigmoid else ""))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    for n in sys.argv[1:] or ["tiny"]:
        if n.endswith(":sigmoid"):
            main(n.split(":")[0], True)
        else:
            main(n)
