"""Detailed parity report: libdgs_b200 vs the reference's own CUDA rasterizer (oracle/_ref) on identical
inputs. Run on the GPU box:  python tests/gpu_parity_report.py [config ...] > gpurun_out/parity.txt
(test infrastructure; the pytest -m gpu tests assert the same quantities)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_utils as pu  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def report(name, sh_degree=3, use_sigmoid=False, P=None, F=None, do_backward=True):
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs(name, sh_degree=sh_degree, P=P, F=F)
    P, F, W, H = scene.means3D.shape[0], view.shape[0], cam.width, cam.height
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    res = {"config": name, "P": P, "F": F, "W": W, "H": H, "sh_degree": sh_degree, "use_sigmoid": use_sigmoid}
    fw = pu.ours_forward(cam, scene, bg, view, proj, campos, use_sigmoid=use_sigmoid)
    dec = pu.ours_decode(fw, P, F, W, H)
    res["num_rendered"] = fw["num_rendered"]
    mism = {k: 0 for k in ["radii", "tiles_touched", "depth_bits", "means2D_bits", "conic_bits", "keys", "point_list",
                           "ranges", "n_contrib", "final_T_bits", "num_rendered"]}
    maxabs = {k: 0.0 for k in ["conic_opacity", "rgb", "color", "depth", "blur"]}
    refs = []
    base = 0
    tb = dec["tile_bits"]
    vis_total = 0
    for s in range(F):
        r = ref_cuda.forward(scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None,
                             view[s].contiguous(), proj[s].contiguous(), campos[s].contiguous(), bg, W, H,
                             cam.tanfovx, cam.tanfovy, sh_degree, use_sigmoid=use_sigmoid)
        refs.append(r)
        vis = r["radii"] > 0
        vis_total += int(vis.sum())
        mism["radii"] += int((r["radii"] != fw["radii"][s]).sum())
        mism["tiles_touched"] += int((r["geom"]["tiles_touched"] != dec["tiles_touched"][s]).sum())
        mism["depth_bits"] += int((r["geom"]["depths"].view(torch.int32)[vis] != dec["depths"][s].view(torch.int32)[vis]).sum())
        m2 = r["geom"]["means2D"].view(P, 2)
        mism["means2D_bits"] += int((m2.view(torch.int32)[vis] != dec["means2D"][s].view(torch.int32)[vis]).sum())
        co = r["geom"]["conic_opacity"].view(P, 4)
        mism["conic_bits"] += int((co.view(torch.int32)[vis] != dec["conic_opacity"][s].view(torch.int32)[vis]).sum())
        maxabs["conic_opacity"] = max(maxabs["conic_opacity"], float((co[vis] - dec["conic_opacity"][s][vis]).abs().max()) if vis.any() else 0.0)
        rgb = r["geom"]["rgb"].view(P, 3)
        maxabs["rgb"] = max(maxabs["rgb"], float((rgb[vis] - dec["rgb"][s][vis]).abs().max()) if vis.any() else 0.0)
        R = r["num_rendered"]
        seg_keys = dec["keys"][base:base + R]
        seg_list = dec["point_list"][base:base + R]
        low_mask = (1 << (32 + tb)) - 1
        if seg_keys.numel() != R:
            mism["num_rendered"] += 1
        else:
            mism["keys"] += int(((seg_keys & low_mask) != r["binning"]["point_list_keys"]).sum())
            mism["keys"] += int(((seg_keys >> (32 + tb)) != s).sum())
            mism["point_list"] += int((seg_list != r["binning"]["point_list"]).sum())
        rr = r["image"]["ranges"][:2 * tiles].view(tiles, 2)
        mine = dec["ranges"][s].clone()
        nz = (mine[:, 1] > mine[:, 0])
        mine[nz] -= base
        mism["ranges"] += int((mine != rr).any(dim=1).sum())
        mism["n_contrib"] += int((r["image"]["n_contrib"].view(H, W) != dec["n_contrib"][s]).sum())
        mism["final_T_bits"] += int((r["image"]["accum_alpha"].view(torch.int32).view(H, W) != dec["final_T"][s].view(torch.int32)).sum())
        maxabs["color"] = max(maxabs["color"], float((r["color"] - fw["color"][s]).abs().max()))
        maxabs["depth"] = max(maxabs["depth"], float((r["depth"] - fw["depth"][s]).abs().max()))
        base += R
    if base != fw["num_rendered"]:
        mism["num_rendered"] += 1
    ref_blur = torch.stack([r["color"] for r in refs]).mean(dim=0)
    maxabs["blur"] = float((ref_blur - fw["blur"]).abs().max())
    res["visible_total"] = vis_total
    res["mismatches"] = mism
    res["max_abs"] = maxabs

    if do_backward:
        g = torch.Generator().manual_seed(7)
        dL_dpix = (torch.randn(F, 3, H, W, generator=g) / (3 * H * W)).cuda()
        dL_ddepth = (torch.randn(F, 1, H, W, generator=g) / (H * W) * 0.1).cuda()
        mine = pu.ours_backward(cam, scene, bg, view, proj, campos, fw, dL_dpix, dL_ddepth, use_sigmoid=use_sigmoid)
        mine2 = pu.ours_backward(cam, scene, bg, view, proj, campos, fw, dL_dpix, dL_ddepth, use_sigmoid=use_sigmoid)
        sums = [None, None]
        per_s = []
        for rep in range(2):
            acc = None
            for s in range(F):
                b = ref_cuda.backward(refs[s], scene.means3D, scene.shs, None, scene.scales, scene.rotations, None,
                                      view[s].contiguous(), proj[s].contiguous(), campos[s].contiguous(), bg, W, H,
                                      cam.tanfovx, cam.tanfovy, sh_degree, dL_dpix[s].contiguous(),
                                      dL_ddepth[s].contiguous(), use_sigmoid=use_sigmoid)
                if rep == 0:
                    per_s.append(b)
                if acc is None:
                    acc = {k: v.double().clone() for k, v in b.items()}
                else:
                    for k, v in b.items():
                        acc[k] += v.double()
            sums[rep] = acc
        ref, ref2 = sums
        errs, noise, self_noise, errs_max = {}, {}, {}, {}

        def relmax(a, b):   # the bar of tests/test_gpu_parity.py: max-abs error over the tensor's max-abs
            a, b = a.double(), b.double()
            return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
        pairs = {"dL_dmeans3D": "dL_dmeans3D", "dL_dsh": "dL_dsh", "dL_dopacity": "dL_dopacity",
                 "dL_dscales": "dL_dscales", "dL_drotations": "dL_drotations"}
        for k, rk in pairs.items():
            errs[k] = pu.rel_err(mine[k], ref[rk])
            errs_max[k] = relmax(mine[k], ref[rk])
            noise[k] = pu.rel_err(ref2[rk], ref[rk])
            self_noise[k] = pu.rel_err(mine2[k], mine[k])
        rv = torch.stack([b["dL_dviewmatrix"] for b in per_s])
        rp = torch.stack([b["dL_dprojmatrix"] for b in per_s])
        rm2 = torch.stack([b["dL_dmeans2D"] for b in per_s])
        errs["dL_dviewmatrix"] = pu.rel_err(mine["dL_dviewmatrix"], rv)
        errs["dL_dprojmatrix"] = pu.rel_err(mine["dL_dprojmatrix"], rp)
        errs["dL_dmeans2D"] = pu.rel_err(mine["dL_dmeans2D"], rm2)
        errs_max["dL_dviewmatrix"] = relmax(mine["dL_dviewmatrix"], rv)
        errs_max["dL_dprojmatrix"] = relmax(mine["dL_dprojmatrix"], rp)
        errs_max["dL_dmeans2D"] = relmax(mine["dL_dmeans2D"], rm2)
        res["grad_rel_err"] = errs
        res["grad_maxabs_rel_err"] = errs_max
        res["ref_vs_ref_noise"] = noise
        res["ours_vs_ours_noise"] = self_noise
    return res


if __name__ == "__main__":
    names = sys.argv[1:] or ["tiny", "c1"]
    for n in names:
        t0 = time.time()
        kw = {}
        if ":" in n:
            n, opt = n.split(":")
            if opt == "sigmoid":
                kw["use_sigmoid"] = True
            elif opt.startswith("deg"):
                kw["sh_degree"] = int(opt[3:])
        r = report(n, **kw)
        r["seconds"] = round(time.time() - t0, 2)
        print(json.dumps(r), flush=True)
