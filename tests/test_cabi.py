"""CPU: the C-ABI library loads, exports every symbol include/dgs_b200.h declares, validates arguments
before touching the device, and the Python product path fails loudly without it / without CUDA tensors."""
import ctypes as C
import os
import re

import pytest
import torch

from deblurgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dgs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgs_[a-z0-9_]+)\s*\(", src)) - {"dgs_alloc_fn"})


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_binding_covers_every_declared_symbol():
    assert set(_declared_symbols()) == set(_lib.SIGNATURES.keys())


def test_introspection_and_key_bits():
    lib = _lib.load()
    assert lib.dgs_version() >= 100 and lib.dgs_compiled_arch() == 1000
    tb, sb = C.c_int(), C.c_int()
    # reference: bit = getHigherMsb(tiles) (rasterizer_impl.cu:35-50): 256 tiles -> 9, 950 -> 10, 8160 -> 13
    for (w, h, f, etb, esb) in [(256, 256, 4, 9, 2), (600, 400, 16, 10, 4), (1920, 1080, 16, 13, 4),
                                (1280, 720, 32, 12, 5), (96, 64, 1, 5, 0)]:
        assert lib.dgs_key_bits(w, h, f, C.byref(tb), C.byref(sb)) == 0
        assert (tb.value, sb.value) == (etb, esb)
    assert lib.dgs_blur_backward_scratch_bytes(1000, 4) >= 1000 * 4 * 48
    assert lib.dgs_knn_scratch_bytes(1000) > 0
    assert lib.dgs_profile_num_stages() == 12


def test_argument_validation_happens_before_any_device_work():
    lib = _lib.load()
    n = C.c_int64(0)
    cb = _lib.ALLOC_FN(lambda ctx, nbytes: 0)
    one = C.c_void_p(1)   # never dereferenced: validation fails first
    # both SHs and precomputed colours -> the reference raises in Python (diff_gaussian_rasterization/__init__.py:210)
    rc = lib.dgs_blur_forward(cb, None, cb, None, cb, None, 10, 1, 3, 16, one, 32, 32, one, one, one, one, one, 1.0,
                              one, None, one, one, one, 1.0, 1.0, 0.2, 100.0, 0, 0, one, one, one, None, 1.0,
                              C.byref(n), None)
    assert rc == -1 and b"exactly one" in lib.dgs_last_error()
    # neither scale/rotation nor covariance
    rc = lib.dgs_blur_forward(cb, None, cb, None, cb, None, 10, 1, 3, 16, one, 32, 32, one, one, None, one, None, 1.0,
                              None, None, one, one, one, 1.0, 1.0, 0.2, 100.0, 0, 0, one, one, one, None, 1.0,
                              C.byref(n), None)
    assert rc == -1
    # bad SH degree
    rc = lib.dgs_blur_forward(cb, None, cb, None, cb, None, 10, 1, 4, 16, one, 32, 32, one, one, None, one, one, 1.0,
                              one, None, one, one, one, 1.0, 1.0, 0.2, 100.0, 0, 0, one, one, one, None, 1.0,
                              C.byref(n), None)
    assert rc == -1 and b"SH" in lib.dgs_last_error()
    # allocation failure is reported, not dereferenced
    rc = lib.dgs_blur_forward(cb, None, cb, None, cb, None, 10, 1, 3, 16, one, 32, 32, one, one, None, one, one, 1.0,
                              one, None, one, one, one, 1.0, 1.0, 0.2, 100.0, 0, 0, one, one, one, None, 1.0,
                              C.byref(n), None)
    assert rc == -3


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdgs_b200.so")
    with pytest.raises(_lib.DgsError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected_not_emulated():
    from deblurgs_b200 import GaussianRasterizationSettings, GaussianRasterizer, distCUDA2, bezier_se3_poses
    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, 0.2, 100.0, False, 0, torch.zeros(3),
                                       False, False)
    r = GaussianRasterizer(rs)
    P = 4
    with pytest.raises(_lib.DgsError, match="no CPU path"):
        r(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), opacities=torch.ones(P, 1), shs=torch.zeros(P, 1, 3),
          scales=torch.ones(P, 3), rotations=torch.ones(P, 4), viewmatrix=torch.eye(4), projmatrix=torch.eye(4))
    with pytest.raises(_lib.DgsError, match="no CPU path"):
        distCUDA2(torch.zeros(10, 3))
    with pytest.raises(_lib.DgsError, match="no CPU path"):
        bezier_se3_poses(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(3), torch.eye(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deblurgs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test infrastructure", ""), f
