"""Generates tests/golden/se3_log_golden.pt by running the REFERENCE's own se3_log_map / se3_exp_map
(utils/pytorch3d_functions.py:373-541) on the CPU in the authoring container, where /root/reference is
mounted: fp32 and fp64 inputs, including near-identity rotations (the eps-clamped branch), rotations
near pi, and the reference's _set_initial_parameters layout (scene/motion.py:196-204).

Run:  python tests/golden/make_se3log_golden.py      (needs /root/reference; not needed on the GPU box)
"""
import importlib.util
import os

import torch

REF = os.environ.get("DEBLURGS_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_p3d", os.path.join(REF, "utils", "pytorch3d_functions.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    g = torch.Generator().manual_seed(7)
    cases = {}
    for tag, dtype, rot_scale in [("f32", torch.float32, 0.8), ("f64", torch.float64, 0.8),
                                  ("tiny_f32", torch.float32, 2e-3), ("tiny_f64", torch.float64, 1e-5),
                                  ("large_f64", torch.float64, 1.7)]:
        v = torch.randn(48, 6, generator=g, dtype=torch.float64)
        v[:, 3:] *= rot_scale
        v = v.to(dtype)
        T = m.se3_exp_map(v)
        cases[tag] = dict(log_in=v, transform=T, log_out=m.se3_log_map(T))
    # the layout CameraMotionModule._set_initial_parameters feeds: c2w rotation transposed, position in row 3
    q = torch.linalg.qr(torch.randn(16, 3, 3, generator=g, dtype=torch.float64))[0]
    q = q * torch.sign(torch.det(q))[:, None, None]
    pos = torch.randn(16, 3, generator=g, dtype=torch.float64)
    c2w = torch.zeros(16, 4, 4, dtype=torch.float64)
    c2w[:, :3, :3] = q.transpose(-2, -1)
    c2w[:, 3, :3] = pos
    c2w[:, 3, 3] = 1.0
    cases["init_layout"] = dict(rotations=q, translations=pos, transform=c2w, log_out=m.se3_log_map(c2w))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "se3_log_golden.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
