"""a10 thresholding: oracle pinned on the compiled reference (CPU); CUDA vs oracle (GPU, bit-exact)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform, frame_smooth, frame_text, frame_const

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
SIZES = [(64, 48, 64), (100, 37, 128), (257, 65, 320), (640, 480, 640)]


def _frames(w, h, stride):
    return [frame_g(w, h, 12345, stride), frame_uniform(w, h, 1, stride), frame_smooth(w, h, 3, stride), frame_text(w, h, 7, stride), frame_const(w, h, 77, stride)]


@needs_ref
@pytest.mark.parametrize("w,h,stride", SIZES)
def test_oracle_thresholds_vs_reference(w, h, stride):
    for img in _frames(w, h, stride):
        for t in (0.0, 99.6, 128.0, 300.0):
            np.testing.assert_array_equal(oracle.threshold("orc", "global", img, threshold=t, width=w)[0][:, :w], oracle.threshold("ref", "global", img, threshold=t, width=w)[0][:, :w])
        a, ta = oracle.threshold("orc", "otsu", img, width=w)
        b, tb = oracle.threshold("ref", "otsu", img, width=w)
        assert ta == tb
        np.testing.assert_array_equal(a[:, :w], b[:, :w])
        for (bs, delta, mv, inv) in [(5, 8.0, 255.0, False), (5, 8.0, 255.0, True), (11, 3.4, 200.0, False), (3, 0.0, 255.0, False)]:
            if bs > min(w, h):
                continue
            np.testing.assert_array_equal(oracle.threshold("orc", "adaptive", img, block_size=bs, delta=delta, max_val=mv, invert=inv, width=w)[0][:, :w],
                                          oracle.threshold("ref", "adaptive", img, block_size=bs, delta=delta, max_val=mv, invert=inv, width=w)[0][:, :w])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,stride", SIZES + [(1920, 1080, 1920), (3840, 2160, 3840)])
def test_cuda_thresholds(cvb, w, h, stride):
    for img in _frames(w, h, stride):
        np.testing.assert_array_equal(cvb.histogram(img, width=w), oracle.histogram(img, width=w))
        for t in (0.0, 99.6, 300.0):
            np.testing.assert_array_equal(cvb.threshold_global(img, t, width=w)[:, :w], oracle.threshold("orc", "global", img, threshold=t, width=w)[0][:, :w])
        a, ta = cvb.threshold_otsu(img, width=w)
        b, tb = oracle.threshold("orc", "otsu", img, width=w)
        assert ta == tb
        np.testing.assert_array_equal(a[:, :w], b[:, :w])
        assert cvb.threshold_otsu(img, width=w, want_output=False)[1] == tb
        for (bs, delta, mv, inv) in [(5, 8.0, 255.0, False), (5, 8.0, 255.0, True), (11, 3.4, 200.0, False), (31, 10.0, 255.0, False)]:
            if bs > min(w, h):
                continue
            np.testing.assert_array_equal(cvb.threshold_adaptive(img, bs, delta, mv, inv, width=w)[:, :w],
                                          oracle.threshold("orc", "adaptive", img, block_size=bs, delta=delta, max_val=mv, invert=inv, width=w)[0][:, :w])


@pytest.mark.gpu
def test_cuda_threshold_errors(cvb):
    from compv_b200 import _ffi
    img = frame_uniform(64, 48, 1)
    with pytest.raises(_ffi.CvbError) as e:
        cvb.threshold_global(img, -1.0)                      # compv_image_threshold.cxx:120
    assert e.value.code == _ffi.E_INVALID_PARAMETER
    with pytest.raises(_ffi.CvbError) as e:
        cvb.threshold_adaptive(img, block_size=4)            # even block size (compv_image_threshold.cxx:185)
    assert e.value.code == _ffi.E_INVALID_PARAMETER


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [3, 5, 7])
@pytest.mark.parametrize("w,h,stride", [(120, 60, 128), (121, 61, 128), (240, 119, 240), (250, 125, 256), (9, 11, 16), (640, 360, 640)])
def test_cuda_adaptive_tma_path(cvb, bs, w, h, stride):
    """16-byte aligned strides take the TMA-staged kernel: block sizes 3/5/7, tile edges, tiny frames, both polarities, padding bytes ignored."""
    from frames import frame_uniform, frame_text
    img = frame_uniform(w, h, bs * 7 + w, stride) if w % 2 else np.ascontiguousarray(np.pad(frame_text(w, h, bs), ((0, 0), (0, stride - w)), constant_values=33))
    for delta, mv, inv in [(8.0, 255.0, False), (0.0, 200.0, True), (40.5, 255.0, False)]:
        np.testing.assert_array_equal(cvb.threshold_adaptive(img, bs, delta, mv, inv, width=w)[:, :w],
                                      oracle.threshold("orc", "adaptive", img, block_size=bs, delta=delta, max_val=mv, invert=inv, width=w)[0][:, :w])
