"""Small forward+backward used under compute-sanitizer (memcheck / racecheck / initcheck) on the GPU box:
  compute-sanitizer --tool memcheck python tests/sanitizer_case.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_utils as pu  # noqa: E402
from deblurgs_b200 import distCUDA2  # noqa: E402

cam, scene, traj, bg, view, proj, campos = pu.make_inputs("tiny")
F, W, H = view.shape[0], cam.width, cam.height
fw = pu.ours_forward(cam, scene, bg, view, proj, campos)
g = torch.Generator().manual_seed(0)
b = pu.ours_backward(cam, scene, bg, view, proj, campos, fw, (torch.randn(F, 3, H, W, generator=g) / (H * W)).cuda(),
                     (torch.randn(F, 1, H, W, generator=g) / (H * W)).cuda())
d = distCUDA2(scene.means3D)
torch.cuda.synchronize()
print("ok", fw["num_rendered"], float(b["dL_dmeans3D"].abs().sum()), float(d.mean()))
