"""K1-K3 parity: CUDA separable convolution vs the oracle (bit-exact, including the float variants: same fma chain)."""
import numpy as np
import pytest

import oracle
from frames import frame_uniform

pytestmark = pytest.mark.gpu

SIZES = [(64, 48, 64), (100, 37, 128), (257, 65, 320), (640, 480, 640), (1285, 720, 1344)]


def _kernels(tk, ks, rng):
    if tk == np.int16:
        return rng.integers(-9, 10, ks).astype(np.int16), rng.integers(-9, 10, ks).astype(np.int16)
    if tk == np.float32:
        return oracle.gauss_kernel("orc", ks, 0.8 + ks / 4.0), (rng.random(ks).astype(np.float32) - np.float32(0.3))
    k = oracle.gauss_kernel("orc", ks, 0.8 + ks / 4.0, fixed_point=True)
    return k, k[::-1].copy()


def _input(tin, w, h, stride, rng):
    if tin == np.uint8:
        return frame_uniform(w, h, int(rng.integers(1 << 30)), stride)
    if tin == np.int16:
        return rng.integers(-4000, 4000, (h, stride)).astype(np.int16)
    return (rng.random((h, stride)) * 400 - 100).astype(np.float32)


def _same(a, b, w):
    if a.dtype == np.float32:
        np.testing.assert_array_equal(a[:, :w].view(np.uint32), b[:, :w].view(np.uint32))
    else:
        np.testing.assert_array_equal(a[:, :w], b[:, :w])


@pytest.mark.parametrize("w,h,stride", SIZES)
@pytest.mark.parametrize("name", list(oracle.CONV_TYPES))
@pytest.mark.parametrize("ks", [3, 5, 7, 15])
def test_convlt1_zero_border(cvb, name, ks, w, h, stride):
    tin, tk, tout = oracle.CONV_TYPES[name]
    rng = np.random.default_rng(ks * 1000 + w)
    img = _input(tin, w, h, stride, rng)
    vt, hz = _kernels(tk, ks, rng)
    _same(cvb.convlt1(name, img, vt, hz, width=w), oracle.convlt1("orc", name, img, vt, hz, width=w), w)


@pytest.mark.parametrize("name", ["8u16s16s", "8u32f32f", "8u32f8u", "fxp_8u16u8u"])
@pytest.mark.parametrize("border", [1, 2])
def test_convlt1_other_borders(cvb, name, border):
    tin, tk, tout = oracle.CONV_TYPES[name]
    w, h, stride = 333, 77, 384
    rng = np.random.default_rng(border)
    img = _input(tin, w, h, stride, rng)
    vt, hz = _kernels(tk, 5, rng)
    base = rng.integers(0, 100, (h, stride)).astype(tout)  # "ignore" must leave these values in the border ring
    a = cvb.convlt1(name, img, vt, hz, width=w, border=border, out=base.copy())
    b = oracle.convlt1("orc", name, img, vt, hz, width=w, border=border, out=base.copy())
    _same(a, b, w)


def test_convlt1_minimum_size_and_errors(cvb):
    from compv_b200 import _ffi
    img = frame_uniform(7, 7, 3, 16)
    k = np.array([1, -2, 3, 4, 3, -2, 1], np.int16)
    _same(cvb.convlt1("8u16s16s", img, k, k, width=7), oracle.convlt1("orc", "8u16s16s", img, k, k, width=7), 7)
    with pytest.raises(_ffi.CvbError) as e:   # kernel larger than the image (compv_math_convlt.h:101)
        cvb.convlt1("8u16s16s", np.ascontiguousarray(img[:5]), k, k, width=7)
    assert e.value.code == _ffi.E_INVALID_PARAMETER
    with pytest.raises(_ffi.CvbError) as e:   # even kernel size
        cvb.convlt1("8u16s16s", img, k[:4], k[:4], width=7)
    assert e.value.code == _ffi.E_INVALID_PARAMETER


def test_gauss_kernels(cvb):
    for size, sigma in [(3, 0.8), (5, 1.0), (7, 2.0), (9, 1.7)]:
        np.testing.assert_array_equal(cvb.gauss_kernel(size, sigma).view(np.uint32), oracle.gauss_kernel("orc", size, sigma).view(np.uint32))
        np.testing.assert_array_equal(cvb.gauss_kernel(size, sigma, True), oracle.gauss_kernel("orc", size, sigma, True))


@pytest.mark.gpu
@pytest.mark.parametrize("ks", [3, 5])
@pytest.mark.parametrize("border", [0, 1, 2])
@pytest.mark.parametrize("w,h,stride", [(120, 60, 128), (121, 61, 128), (240, 119, 240), (250, 125, 256), (7, 9, 16), (333, 77, 384)])
@pytest.mark.parametrize("packed", [False, True])
def test_convlt1_8u32f8u_tma_path(cvb, ks, border, w, h, stride, packed):
    """16-byte aligned strides take the TMA-staged 4-px-per-lane kernel: every border type, tile edges, tiny frames.  Two instances: negative taps keep the scalar
    chain with the clamp; two normalised Gaussians (clamp is a no-op) take the packed FFMA2 / FADD2 chain."""
    rng = np.random.default_rng(ks * 100 + border * 10 + w)
    img = frame_uniform(w, h, int(rng.integers(1 << 30)), stride)
    vt = oracle.gauss_kernel("orc", ks, 0.7) if packed else (rng.random(ks).astype(np.float32) - np.float32(0.3))
    hz = oracle.gauss_kernel("orc", ks, 1.0)
    base = rng.integers(0, 100, (h, stride)).astype(np.uint8)
    a = cvb.convlt1("8u32f8u", img, vt, hz, width=w, border=border, out=base.copy())
    b = oracle.convlt1("orc", "8u32f8u", img, vt, hz, width=w, border=border, out=base.copy())
    np.testing.assert_array_equal(a[:, :w], b[:, :w])
