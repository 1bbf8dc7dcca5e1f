"""CPU checks of the bloom restatement (oracle/orc_output.c: orc_bloom_apply; reference device/device_post.c:62-140,
cuda/post_common.cuh:71-143): structural properties of the mip chain that do not need a GPU."""
import ctypes as C

import numpy as np

import orc


def _bloom(rgb, w, h, blend):
    L = orc.lib()
    L.orc_bloom_apply.restype = None
    a = np.ascontiguousarray(rgb, np.float32).reshape(-1).copy()
    L.orc_bloom_apply(a.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(w), C.c_uint32(h), C.c_float(blend))
    return a.reshape(3, h, w)


def test_bloom_properties():
    w, h = 96, 54
    rng = np.random.default_rng(2)
    img = rng.random((3, h, w)).astype(np.float32)
    # blend 0 (or a chain of one mip) leaves the image untouched
    assert np.array_equal(_bloom(img, w, h, 0.0), img)
    assert np.array_equal(_bloom(img[:, :3, :3], 3, 3, 0.5), img[:, :3, :3])
    # linear in the image: scaling by a power of two commutes exactly
    a = _bloom(img, w, h, 0.3)
    assert np.array_equal(_bloom(img * np.float32(4.0), w, h, 0.3), a * np.float32(4.0))
    # channels are independent
    solo = img.copy()
    solo[1:] = 0.0
    b = _bloom(solo, w, h, 0.3)
    assert np.array_equal(b[0], a[0]) and not b[1:].any()
    # one bright pixel: the base keeps (1 - blend) of it, the rest spreads over a wide, monotonically fading halo
    spot = np.zeros((3, h, w), np.float32)
    spot[:, 27, 48] = 100.0
    s = _bloom(spot, w, h, 0.2)
    assert abs(s[0, 27, 48] - 80.0) < 2.0
    assert s[0, 27, 49] > s[0, 27, 56] > s[0, 27, 70] > 0.0
    assert (s >= 0).all() and np.isfinite(s).all()
    # the halo carries roughly blend x the energy (border taps and the 1/mip_count weights lose some at the edges)
    assert 0.05 * 100.0 < s[0].sum() - 80.0 < 0.25 * 100.0
    # odd sizes and the degenerate last mips (height 1: 1 / (th - 1) = inf in the reference's arithmetic) stay finite
    odd = rng.random((3, 37, 101)).astype(np.float32)
    assert np.isfinite(_bloom(odd, 101, 37, 0.5)).all()
