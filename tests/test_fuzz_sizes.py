"""Small, odd and awkward frame sizes through every detector (CUDA vs oracle): tile edges, widths below a warp, strides wider than the width."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform


def sizes():
    rng = np.random.default_rng(99)
    out = [(16, 16), (17, 9), (9, 33), (31, 31), (33, 8), (120, 60), (121, 61), (127, 59), (129, 7), (255, 16)]
    out += [(int(rng.integers(8, 200)), int(rng.integers(8, 120))) for _ in range(10)]
    return out


def padded(img, pad):
    """Same pixels in a buffer whose stride is `pad` samples wider than the width (padding filled with a value no detector may read)."""
    h, w = img.shape
    out = np.full((h, w + pad), 201, np.uint8)
    out[:, :w] = img
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", sizes())
def test_cuda_small_sizes_every_detector(cvb, w, h):
    from compv_b200 import _ffi
    img = frame_g(w, h, w * 131 + h) if (w + h) % 2 else frame_uniform(w, h, w + h)
    for pad in (0, 3):
        buf = padded(img, pad) if pad else img
        kw = dict(width=w)
        # Canny (plain and with the fused 5-tap blur when the frame is large enough for the kernel), Sobel
        got = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3).process(buf, **kw)
        want = oracle.edge_dete("orc", buf, "canny", 59.0, 119.0, 3, **kw)
        np.testing.assert_array_equal(got[:, :w], want[:, :w])
        np.testing.assert_array_equal(cvb.CompVEdgeDete.newObj(_ffi.SOBEL_ID).process(buf, **kw)[:, :w], oracle.edge_dete("orc", buf, "sobel", 0, 0, 3, **kw)[:, :w])
        edges = np.ascontiguousarray(want)
        for hid, fn, thr in ((_ffi.HOUGHKHT_ID, oracle.hough_kht, 5), (_ffi.HOUGHSHT_ID, oracle.hough_sht, 8)):
            a = cvb.CompVHough.newObj(hid, 1.0, 1.0, thr).process(edges, capacity=1 << 18, **kw)
            o = fn("orc", edges, 1.0, 1.0, thr, **kw)[0] if hid == _ffi.HOUGHKHT_ID else fn("orc", edges, 1.0, 1.0, thr, cap=1 << 18, **kw)[0]
            assert len(a) == len(o)
            for key in ("rho", "theta", "strength"):
                np.testing.assert_array_equal(a[key], o[key])
        fast = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
        fast.setInt(_ffi.FAST_SET_INT_MAX_FEATURES, -1)
        a, o = fast.process(buf, **kw), oracle.fast_detect("orc", buf, 9, 20, True, **kw)
        assert len(a) == len(o)
        for key in ("x", "y", "strength"):
            np.testing.assert_array_equal(a[key], o[key])
        out, thr = cvb.threshold_otsu(buf, **kw)
        wo, wt = oracle.threshold("orc", "otsu", buf, **kw)
        assert thr == wt
        np.testing.assert_array_equal(out[:, :w], wo[:, :w])
        np.testing.assert_array_equal(cvb.threshold_adaptive(buf, **kw)[:, :w], oracle.threshold("orc", "adaptive", buf, **kw)[0][:, :w])
        binar = np.ascontiguousarray(wo)
        r = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID).process(binar, **kw)
        wl = oracle.ccl_lsl("orc", binar, **kw)
        assert r.labelsCount() == wl["na"]
        np.testing.assert_array_equal(r.debugFlatten(), wl["labels"])
        if w >= 3 and h >= 3:
            se = cvb.morph_strel((3, 3), 2)
            np.testing.assert_array_equal(cvb.morph(binar, se, 3, **kw)[:, :w], oracle.morph("orc", binar, se, 3, **kw)[:, :w])
        if w >= 16 and h >= 16:
            np.testing.assert_array_equal(cvb.CompVHOG.newObj().process(buf, **kw), oracle.hog("orc", buf, **kw))
    # MSER: the stride is part of the neighbourhood rule, so only the tight buffer is compared
    mser = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, delta=2, min_area=0.001, max_area=0.5, max_variation=0.5, min_diversity=0.2)
    got = mser.process(img).regions()
    want = oracle.ccl_lmser("orc", img, delta=2, min_area=0.001, max_area=0.5, max_variation=0.5, min_diversity=0.2)
    assert sorted(got["sizes"].tolist()) == sorted(want["sizes"].tolist())
