"""a3/a5 parity: Sobel family + Canny on the GPU vs the oracle (bit-exact) and, where present, the compiled reference."""
import ctypes

import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform, frame_smooth, frame_const

pytestmark = pytest.mark.gpu

SIZES = [(64, 48, 64), (100, 37, 128), (257, 65, 320), (640, 480, 640), (1282, 720, 1344)]
KIND_ID = {"sobel": 27, "scharr": 28, "prewitt": 29, "canny": 20}


def _frames(w, h, stride):
    return [frame_g(w, h, 12345, stride), frame_uniform(w, h, 1, stride), frame_smooth(w, h, 3, stride), frame_const(w, h, 200, stride)]


@pytest.mark.parametrize("w,h,stride", SIZES)
@pytest.mark.parametrize("kind,ks", [("sobel", 3), ("sobel", 5), ("scharr", 3), ("prewitt", 3)])
def test_sobel_g(cvb, kind, ks, w, h, stride):
    for img in _frames(w, h, stride):
        a = cvb.sobel_g(img, KIND_ID[kind], ks, width=w)
        b = oracle.sobel_g("orc", img, kind, ks, width=w)
        for pa, pb in zip(a, b):
            np.testing.assert_array_equal(pa[:, :w], pb[:, :w])


@pytest.mark.parametrize("w,h,stride", SIZES)
@pytest.mark.parametrize("kind", ["sobel", "scharr", "prewitt"])
def test_edge_normalized(cvb, kind, w, h, stride):
    d = cvb.CompVEdgeDete.newObj(KIND_ID[kind])
    q = cvb.CompVEdgeDete.newObj(KIND_ID[kind])
    q.setBool(cvb.CompVEdgeDete.EDGE_SET_BOOL_X86_SSE41_GMAX_LANES, True)
    for img in _frames(w, h, stride):
        np.testing.assert_array_equal(d.process(img, width=w)[:, :w], oracle.edge_dete("orc", img, kind, width=w)[:, :w])
        np.testing.assert_array_equal(q.process(img, width=w)[:, :w], oracle.edge_dete("orc", img, kind, width=w, sse41_gmax_lanes=True)[:, :w])
        if oracle.have_ref():  # the compiled reference's default (SIMD) path, straight
            np.testing.assert_array_equal(q.process(img, width=w)[:, :w], oracle.edge_dete("ref", img, kind, width=w)[:, :w])


@pytest.mark.parametrize("w,h,stride", SIZES + [(1920, 1080, 1920)])
@pytest.mark.parametrize("ks,tlow,thigh", [(3, 59.0, 119.0), (3, 20.0, 300.0), (5, 300.0, 900.0)])
def test_canny(cvb, ks, tlow, thigh, w, h, stride):
    d = cvb.CompVEdgeDete.newObj(KIND_ID["canny"], tlow, thigh, ks)
    dg = cvb.CompVEdgeDete.newObj(KIND_ID["canny"], tlow, thigh, ks)
    dg.setBool(cvb.CompVEdgeDete.EDGE_SET_BOOL_GENERIC_KERNEL, True)  # the non-TMA generic front kernel must agree too
    for img in _frames(w, h, stride):
        a = d.process(img, width=w)
        b = oracle.edge_dete("orc", img, "canny", tlow, thigh, ks, width=w)
        np.testing.assert_array_equal(a[:, :w], b[:, :w])
        np.testing.assert_array_equal(dg.process(img, width=w)[:, :w], b[:, :w])
        assert set(np.unique(a[:, :w])) <= {0, 255}
        if oracle.have_ref() and (w - 1) % 16:
            # single-threaded reference: its multi-threaded hysteresis is racy (see test_oracle_vs_ref.test_canny_reference_mt_is_a_subset)
            np.testing.assert_array_equal(a[:, :w], oracle.edge_dete("ref", img, "canny", tlow, thigh, ks, width=w, threads=1)[:, :w])


def test_canny_in_place_and_caps(cvb):
    from compv_b200 import _ffi
    w, h, stride = 320, 200, 320
    img = frame_g(w, h, 5, stride)
    d = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
    expect = oracle.edge_dete("orc", img, "canny", 59.0, 119.0, 3, width=w)
    buf = img.copy()
    d.process(buf, width=w, edges=buf)  # image == edges (canny_dete.cxx:122)
    np.testing.assert_array_equal(buf, expect)
    # caps: same size checks / error codes as canny_dete.cxx:77-117
    assert d.set(_ffi.CANNY_SET_FLT32_THRESHOLD_LOW, -1.0, ctypes.c_float) == _ffi.E_INVALID_PARAMETER
    assert d.set(_ffi.CANNY_SET_INT_KERNEL_SIZE, 4, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER
    d.setInt(_ffi.CANNY_SET_INT_KERNEL_SIZE, 5)
    d.setFloat32(_ffi.CANNY_SET_FLT32_THRESHOLD_LOW, 300.0)
    d.setFloat32(_ffi.CANNY_SET_FLT32_THRESHOLD_HIGH, 900.0)
    np.testing.assert_array_equal(d.process(img, width=w), oracle.edge_dete("orc", img, "canny", 300.0, 900.0, 5, width=w))
    # tLow >= tHigh -> E_INVALID_STATE (canny_dete.cxx:126)
    d.setFloat32(_ffi.CANNY_SET_FLT32_THRESHOLD_HIGH, 100.0)
    with pytest.raises(_ffi.CvbError) as e:
        d.process(img, width=w)
    assert e.value.code == _ffi.E_INVALID_STATE


def test_canny_percent_of_mean(cvb):
    from compv_b200 import _ffi
    w, h, stride = 320, 200, 320
    d = cvb.CompVEdgeDete.newObj(20, 0.8, 1.6, 3)
    d.setInt(_ffi.CANNY_SET_INT_THRESHOLD_TYPE, _ffi.CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN)
    for img in _frames(w, h, stride)[:3]:
        np.testing.assert_array_equal(d.process(img, width=w), oracle.edge_dete("orc", img, "canny", 0.8, 1.6, 3, width=w, threshold_type=1))


@pytest.mark.parametrize("size,sigma", [(3, 0.8), (5, 1.0), (7, 2.0)])
def test_canny_fused_preblur_equals_blur_then_canny(cvb, size, sigma):
    """BASELINE config 2: Gaussian (convlt1<u8,f32,u8>) + Canny; the fused kernel must equal the two-step pipeline bit for bit."""
    for (w, h, stride) in [(257, 130, 320), (640, 480, 640), (1920, 1080, 1920)]:
        d = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
        d.set_preblur(size, sigma)
        dg = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
        dg.set_preblur(size, sigma)
        dg.setBool(cvb.CompVEdgeDete.EDGE_SET_BOOL_GENERIC_KERNEL, True)
        for img in _frames(w, h, stride)[:3]:
            k = oracle.gauss_kernel("orc", size, sigma)
            blurred = oracle.convlt1("orc", "8u32f8u", img, k, k, width=w)
            expect = oracle.edge_dete("orc", blurred, "canny", 59.0, 119.0, 3, width=w)
            np.testing.assert_array_equal(d.process(img, width=w)[:, :w], expect[:, :w])
            np.testing.assert_array_equal(dg.process(img, width=w)[:, :w], expect[:, :w])


def test_canny_batched_device_api(cvb):
    """_dev entry point: a batch of different frames in one call equals frame-by-frame results."""
    import torch
    w, h, stride, batch = 640, 480, 640, 5
    frames = np.stack([frame_g(w, h, 100 + k, stride) for k in range(batch - 1)] + [frame_smooth(w, h, 9, stride)])
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty_like(d_in)
    d = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
    d.process_dev(d_in, w, h, stride, d_out, batch=batch, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for k in range(batch):
        np.testing.assert_array_equal(out[k], oracle.edge_dete("orc", frames[k], "canny", 59.0, 119.0, 3, width=w))


def test_canny_long_edges_cross_tiles(cvb):
    """A spiral of weak pixels seeded by one strong spot: the closure must cross many 64x64 hysteresis tiles."""
    w = h = stride = 512
    img = np.full((h, stride), 100, np.uint8)
    x0, y0, x1, y1 = 20, 20, w - 20, h - 20
    while x1 - x0 > 40:
        img[y0:y0 + 3, x0:x1] = 122
        img[y0:y1, x1 - 3:x1] = 122
        img[y1 - 3:y1, x0 + 20:x1] = 122
        img[y0 + 20:y1, x0 + 20:x0 + 23] = 122
        x0 += 20; y0 += 20; x1 -= 20; y1 -= 20
    img[20:23, 30:40] = 200
    d = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
    a = d.process(img)
    b = oracle.edge_dete("orc", img, "canny", 59.0, 119.0, 3)
    np.testing.assert_array_equal(a, b)
    assert (a == 255).sum() > 2000
