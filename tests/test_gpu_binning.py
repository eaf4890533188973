"""GPU tests of the library's own binning machinery (dgs_binning.cu): the segmented LSD radix sort against
torch.sort(stable=True), and the speculative (no host synchronisation) forward against the exact one, including the
capacity-overflow retry and the fully asynchronous variant."""
import ctypes as C

import pytest
import torch

from tests import parity_utils as pu
from deblurgs_b200 import _lib
from deblurgs_b200 import rasterizer as rz

pytestmark = pytest.mark.gpu


def _sort(keys, key_bits):
    lib = _lib.load()
    nseg, n = keys.shape
    ks, idx = torch.empty_like(keys), torch.empty_like(keys)
    scratch = torch.empty(int(lib.dgs_debug_sort_scratch_bytes(nseg, n)), dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dgs_debug_sort(nseg, n, key_bits, _lib.ptr(keys), _lib.ptr(ks), _lib.ptr(idx), _lib.ptr(scratch), st),
               "dgs_debug_sort")
    torch.cuda.synchronize()
    return ks, idx


@pytest.mark.parametrize("nseg,n,bits", [(1, 1, 32), (1, 5, 32), (3, 4095, 32), (2, 4096, 32), (5, 4097, 32),
                                         (16, 300_000, 32), (1, 1_000_003, 30), (4, 70_001, 8), (2, 12_345, 13),
                                         (32, 50_000, 32)])
def test_segmented_radix_sort_matches_torch_stable_sort(nseg, n, bits):
    g = torch.Generator().manual_seed(nseg * 1000 + n)
    hi = (1 << bits) - 1
    keys = torch.randint(0, hi + 1 if bits < 32 else 1 << 31, (nseg, n), generator=g, dtype=torch.int64)
    if bits == 32:
        keys = keys * 2 + torch.randint(0, 2, (nseg, n), generator=g)          # full 32-bit range
    keys[:, ::7] = keys[:, :1].clone()                                          # many ties: stability matters
    if n > 10:
        keys[0, n // 2:] = hi                                                   # a long run of the maximum (culled entries)
    dev = keys.cuda()
    as_i32 = (dev & 0xFFFFFFFF).to(torch.int64)
    packed = torch.where(as_i32 >= 2 ** 31, as_i32 - 2 ** 32, as_i32).to(torch.int32).contiguous()
    ks, idx = _sort(packed, bits)
    ref = torch.sort(dev, dim=1, stable=True)
    assert torch.equal(idx.to(torch.int64), ref.indices)
    assert torch.equal(ks.to(torch.int64) & 0xFFFFFFFF, ref.values)


def _fw_equal(a, b):
    assert a["num_rendered"] == b["num_rendered"]
    for k in ("color", "depth", "radii", "blur"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_speculative_forward_equals_exact_forward(name):
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs(name)
    P, F, W, H = scene.means3D.shape[0], view.shape[0], cam.width, cam.height
    exact = pu.ours_forward(cam, scene, bg, view, proj, campos, exact=True)
    key = (torch.cuda.current_device(), P, F, H, W)
    assert rz._CAPACITY_HINT[key] >= exact["num_rendered"]
    spec = pu.ours_forward(cam, scene, bg, view, proj, campos)                  # sized from the hint, no early sync
    _fw_equal(exact, spec)
    d0, d1 = pu.ours_decode(exact, P, F, W, H), pu.ours_decode(spec, P, F, W, H)
    for k in ("keys", "point_list", "ranges", "n_contrib", "final_T", "point_offsets"):
        assert torch.equal(d0[k], d1[k]), k
    # a hint that is far too small: the library detects the overflow on the device and re-runs with the exact size
    rz._CAPACITY_HINT[key] = 1
    retry = pu.ours_forward(cam, scene, bg, view, proj, campos)
    _fw_equal(exact, retry)
    d2 = pu.ours_decode(retry, P, F, W, H)
    for k in ("keys", "point_list", "ranges", "n_contrib"):
        assert torch.equal(d0[k], d2[k]), k
    # gradients from the speculative state = gradients from the exact state
    g = torch.Generator().manual_seed(1)
    dpix = (torch.randn(F, 3, H, W, generator=g) / (H * W)).cuda()
    ddep = torch.zeros(F, 1, H, W).cuda()
    b0 = pu.ours_backward(cam, scene, bg, view, proj, campos, exact, dpix, ddep)
    b1 = pu.ours_backward(cam, scene, bg, view, proj, campos, retry, dpix, ddep)
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dviewmatrix"):
        assert (b0[k] - b1[k]).abs().max() <= 1e-6 * b0[k].abs().max()


def test_async_forward_reports_status_on_the_device():
    """num_rendered == NULL: nothing is waited for; D and the overflow flag are read back on request."""
    lib = _lib.load()
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs("small")
    P, F, W, H = scene.means3D.shape[0], view.shape[0], cam.width, cam.height
    exact = pu.ours_forward(cam, scene, bg, view, proj, campos, exact=True)
    D = exact["num_rendered"]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for cap, want_overflow in ((D + 10, 0), (max(D // 2, 1), 1)):
        geom, binning, img = rz._Buffer(scene.means3D.device), rz._Buffer(scene.means3D.device), rz._Buffer(scene.means3D.device)
        color = torch.empty((F, 3, H, W), device="cuda"); depth = torch.empty((F, 1, H, W), device="cuda")
        radii = torch.empty((F, P), dtype=torch.int32, device="cuda"); blur = torch.empty((3, H, W), device="cuda")
        p = _lib.ptr
        rc = lib.dgs_blur_forward_hint(geom.cb, None, binning.cb, None, img.cb, None, P, F, 3, 16, p(bg), W, H,
                                       p(scene.means3D), p(scene.shs), None, p(scene.opacities), p(scene.scales), 1.0,
                                       p(scene.rotations), None, p(view), p(proj), p(campos), cam.tanfovx, cam.tanfovy,
                                       0.2, 100.0, 0, 0, p(color), p(depth), p(radii), p(blur), float(F), cap, None, st)
        _lib.check(rc, "dgs_blur_forward_hint")
        n, ov = C.c_int64(0), C.c_int(0)
        _lib.check(lib.dgs_blur_forward_status(p(geom.t), P, F, C.byref(n), C.byref(ov), st), "status")
        assert n.value == D and ov.value == want_overflow
        if not want_overflow:
            assert torch.equal(color, exact["color"]) and torch.equal(blur, exact["blur"])
        else:
            assert torch.allclose(color, bg.view(1, 3, 1, 1).expand_as(color))    # lists empty: background only
