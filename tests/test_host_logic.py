"""CPU: host-side mirror of the reference interface (argument checking, settings layout, nu sampling,
signature tolerance) -- no device work."""
import inspect

import pytest
import torch

import deblurgs_b200 as dg
from deblurgs_b200 import motion, renderer
from oracle import pose_torch as pt


def test_settings_fields_match_reference_order():
    # DGR/diff_gaussian_rasterization/__init__.py:172-187
    assert dg.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "z_near", "z_far",
        "use_sigmoid", "sh_degree", "campos", "prefiltered", "debug")


def test_rasterizer_forward_signature_and_errors():
    sig = inspect.signature(dg.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp", "viewmatrix", "projmatrix"]
    rs = dg.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, 0.2, 100.0, False, 0, torch.zeros(3),
                                          False, False)
    r = dg.GaussianRasterizer(rs)
    z = torch.zeros(2, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z, means2D=z, opacities=z[:, :1], scales=z, rotations=torch.zeros(2, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(2, 1, 3), colors_precomp=z, scales=z,
          rotations=torch.zeros(2, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(2, 1, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z, means2D=z, opacities=z[:, :1], shs=torch.zeros(2, 1, 3), scales=z, rotations=torch.zeros(2, 4),
          cov3D_precomp=torch.zeros(2, 6))


def test_render_signature_is_the_reference_one():
    # gaussian_renderer/__init__.py:18
    names = list(inspect.signature(renderer.render).parameters)
    assert names[:5] == ["viewpoint_camera", "pc", "bg_color", "scaling_modifier", "override_color"]


def test_nu_sampling_matches_oracle():
    class Cam:
        image_width, image_height, FoVx, FoVy, znear, zfar = 8, 8, 1.0, 1.0, 0.01, 100.0
        projection_matrix = torch.eye(4)
    for F in (3, 4, 16, 21):
        cmm = motion.CameraMotionModule([Cam], torch.zeros(2, 6), curve_order=3, num_subframes=F)
        assert cmm._nu.shape == (2, F - 2)
        nu = cmm._sample_nu_from_alignment(1)
        assert torch.equal(nu.detach(), pt.sample_nu(cmm._nu[1].detach(), F))
        assert nu[0] == 0 and nu[-1] == 1 and torch.all(nu[1:] >= nu[:-1])
        # reference init: linspace(0,1,F) (scene/motion.py:55)
        assert torch.allclose(nu.detach(), torch.linspace(0, 1, F), atol=1e-6)


def test_gaussian_params_activations():
    from deblurgs_b200 import synthetic
    cam = synthetic.make_camera(64, 48)
    sc = synthetic.make_scene(50, cam)
    g = motion.GaussianParams.from_scene(sc)
    assert torch.allclose(g.get_scaling, sc.scales, rtol=1e-6)
    assert torch.allclose(g.get_rotation.norm(dim=1), torch.ones(50), atol=1e-6)
    assert g.get_features.shape == (50, 16, 3) and torch.equal(g.get_features, sc.shs)
    g._opacity.data[0] = 1.7
    g._opacity.data[1] = -0.2
    assert g.get_opacity[0] == 1.0 and g.get_opacity[1] == 0.0     # Clamp, not sigmoid


def test_synthetic_scene_is_deterministic_and_mostly_visible():
    from deblurgs_b200 import synthetic
    cam = synthetic.make_camera(600, 400)
    a, b = synthetic.make_scene(1000, cam), synthetic.make_scene(1000, cam)
    assert torch.equal(a.means3D, b.means3D) and torch.equal(a.shs, b.shs)
    assert abs(cam.tanfovx - 600 / (2 * 1.2 * 600)) < 1e-12
    t = synthetic.make_trajectory(16, 9)
    assert t.ctrl_trans.shape == (10, 3) and torch.equal(t.nu, torch.linspace(0, 1, 16))


def test_compat_packages_resolve_to_the_library(monkeypatch):
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "compat"))
    for name in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C", "gaussian_renderer"):
        sys.modules.pop(name, None)
    dgr = importlib.import_module("diff_gaussian_rasterization")
    knn = importlib.import_module("simple_knn._C")
    gr = importlib.import_module("gaussian_renderer")
    assert dgr.GaussianRasterizer is dg.GaussianRasterizer and knn.distCUDA2 is dg.distCUDA2 and gr.render is dg.render
    for name in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C", "gaussian_renderer"):
        sys.modules.pop(name, None)


def test_integer_claims_the_binning_kernels_rely_on():
    """Two exactness arguments made in csrc/dgs_forward.cu, checked exhaustively / on random data in numpy:
    (1) k_duplicate recovers (row, column) of duplicate j in a w-wide tile rectangle from a float quotient:
        trunc((j + 0.5f) / w) == j // w even if the quotient is off by a few ulp (it uses __fdividef);
    (2) the compact depth key bits(depth) - bits(0.2f) is an exact, order-preserving code of every visible depth."""
    import numpy as np
    for w in list(range(1, 130)) + [240, 480, 512, 1023]:
        j = np.arange(0, w * 1024, dtype=np.int64)            # up to 1024 rows: a 16K-pixel-high image
        q = ((j.astype(np.float32) + np.float32(0.5)) / np.float32(w)).astype(np.float32)
        for ulps in (-2, 0, 2):   # __fdividef: <= 2 ulp
            qq = q if ulps == 0 else (q.view(np.int32) + ulps).view(np.float32)
            assert np.array_equal(qq.astype(np.int64), j // w), (w, ulps)
    rng = np.random.default_rng(0)
    near = np.float32(0.2)
    d = np.concatenate([np.nextafter(near, np.float32(1.0), dtype=np.float32)[None],
                        np.exp(rng.uniform(np.log(0.2001), np.log(8.0e8), 200000)).astype(np.float32)])
    d = d[d > near]
    code = d.view(np.uint32).astype(np.int64) - int(near.view(np.uint32))
    assert code.min() >= 1 and code.max() < (1 << 28) - 1          # fits the 28-bit field (F <= 16) up to 8.6e8
    order_v = np.argsort(d, kind="stable")
    order_c = np.argsort(code, kind="stable")
    assert np.array_equal(order_v, order_c)
    assert np.float32(1.3e4).view(np.uint32) - near.view(np.uint32) < (1 << 27) - 1   # 27-bit field (F <= 32)


def test_reference_python_binds_to_the_library_through_the_compat_shims(tmp_path):
    """Drop-in at the import level: with compat/ on the path, the REFERENCE's own, unmodified gaussian_renderer and
    scene.gaussian_model modules (imported from /root/reference where it is mounted; skipped on the GPU box) resolve
    `diff_gaussian_rasterization` / `simple_knn._C` to this library's classes.  Third-party modules the hot path never
    calls (plyfile, open3d, roma, ...) are stubbed when missing.  Runs in a subprocess to keep sys.modules clean."""
    import os
    import subprocess
    import sys
    ref = os.environ.get("DEBLURGS_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "gaussian_renderer")):
        pytest.skip("reference tree not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "chain.py"
    script.write_text('''
import os, sys, types
class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m
for name in ["roma", "open3d", "plyfile", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "imageio", "lpipsPyTorch"]:
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = _Stub(name)
sys.path[:0] = [%r, os.path.join(%r, "compat"), %r]
import gaussian_renderer, deblurgs_b200
import scene.gaussian_model as gm
assert os.path.samefile(os.path.dirname(gaussian_renderer.__file__), os.path.join(%r, "gaussian_renderer"))
assert gaussian_renderer.GaussianRasterizer is deblurgs_b200.GaussianRasterizer
assert gaussian_renderer.GaussianRasterizationSettings is deblurgs_b200.GaussianRasterizationSettings
assert gm.distCUDA2 is deblurgs_b200.distCUDA2
import inspect
ours = inspect.signature(deblurgs_b200.render).parameters
theirs = inspect.signature(gaussian_renderer.render).parameters
assert list(theirs)[:5] == list(ours)[:5] == ["viewpoint_camera", "pc", "bg_color", "scaling_modifier", "override_color"]
print("ok")
''' % (ref, root, root, ref))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
