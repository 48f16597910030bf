"""a8 FAST: oracle pinned on the compiled reference (CPU); CUDA vs oracle and vs reference (GPU)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform, frame_smooth, frame_const, frame_text

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
SIZES = [(64, 48, 64), (100, 37, 128), (257, 65, 320), (640, 480, 640)]


def _frames(w, h, stride):
    return [frame_g(w, h, 12345, stride), frame_uniform(w, h, 1, stride), frame_smooth(w, h, 3, stride), frame_text(w, h, 7, stride), frame_const(w, h, 90, stride)]


def _same_points(a, b):
    assert len(a) == len(b)
    for f in ("x", "y", "strength", "orient", "level", "size"):
        np.testing.assert_array_equal(a[f], b[f])


@needs_ref
@pytest.mark.parametrize("w,h,stride", SIZES)
@pytest.mark.parametrize("N,threshold,nms", [(9, 20, True), (9, 20, False), (12, 10, True), (12, 30, False), (9, 1, True)])
def test_oracle_fast_vs_reference(N, threshold, nms, w, h, stride):
    for img in _frames(w, h, stride):
        a = oracle.fast_detect("orc", img, N, threshold, nms, width=w)
        for threads in (1, -1):  # raster order must hold for any thread count (fast_dete.cxx:404-408)
            b = oracle.fast_detect("ref", img, N, threshold, nms, max_features=-1, width=w, threads=threads)
            _same_points(a, b)


def test_oracle_scores_are_the_unsuppressed_points():
    img = frame_uniform(100, 60, 5, 128)
    s = oracle.fast_scores(img, 9, 20, width=100)
    pts = oracle.fast_detect("orc", img, 9, 20, False, width=100)
    assert len(pts) == int((s != 0).sum())
    for p in pts[:50]:
        assert s[int(p["y"]), int(p["x"])] + 19 == p["strength"]


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("w,h,stride", SIZES + [(1282, 720, 1344), (1920, 1080, 1920)])
@pytest.mark.parametrize("N,threshold", [(9, 20), (12, 10), (9, 1), (12, 40)])
def test_cuda_fast_scores(cvb, N, threshold, w, h, stride):
    for img in _frames(w, h, stride):
        np.testing.assert_array_equal(cvb.fast_scores(img, N, threshold, width=w)[:, :w], oracle.fast_scores(img, N, threshold, width=w)[:, :w])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,stride", SIZES + [(1282, 720, 1344), (1920, 1080, 1920)])
@pytest.mark.parametrize("N,threshold,nms", [(9, 20, True), (9, 20, False), (12, 10, True), (9, 1, True)])
def test_cuda_fast_points(cvb, N, threshold, nms, w, h, stride):
    from compv_b200 import _ffi
    d = cvb.CompVCornerDete.newObj()
    d.setInt(_ffi.FAST_SET_INT_THRESHOLD, threshold)
    d.setInt(_ffi.FAST_SET_INT_FAST_TYPE, _ffi.FAST_TYPE_12 if N == 12 else _ffi.FAST_TYPE_9)
    d.setBool(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, nms)
    d.setInt(_ffi.FAST_SET_INT_MAX_FEATURES, -1)
    for img in _frames(w, h, stride):
        a = d.process(img, width=w)
        _same_points(a, oracle.fast_detect("orc", img, N, threshold, nms, width=w))
        if oracle.have_ref():
            _same_points(a, oracle.fast_detect("ref", img, N, threshold, nms, max_features=-1, width=w, threads=1))


@pytest.mark.gpu
def test_cuda_fast_select_best_matches_reference(cvb):
    """Default object: maxFeatures=2000 -> CompVInterestPoint::selectBest (nth_element + partition).  Same libstdc++ => same order."""
    if not oracle.have_ref():
        pytest.skip("needs the compiled reference")
    d = cvb.CompVCornerDete.newObj()
    for img in [frame_uniform(640, 480, 1), frame_g(1920, 1080, 12345), frame_text(640, 480, 7)]:
        a = d.process(img)
        b = oracle.fast_detect("ref", img, 9, 20, True, max_features=2000, threads=1)
        _same_points(a, b)


@pytest.mark.gpu
def test_cuda_fast_batched_device_api(cvb):
    import torch
    w, h, stride, batch, cap = 640, 480, 640, 4, 40000
    frames = np.stack([frame_g(w, h, 100 + k, stride) for k in range(batch - 1)] + [frame_uniform(w, h, 9, stride)])
    d_in = torch.from_numpy(frames).cuda()
    d_pts = torch.zeros((batch, cap, 6), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(batch, dtype=torch.int32, device="cuda")
    d = cvb.CompVCornerDete.newObj()
    d.process_dev(d_in, w, h, stride, d_pts, cap, d_cnt, batch=batch, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    pts = d_pts.cpu().numpy()
    for k in range(batch):
        want = oracle.fast_detect("orc", frames[k], 9, 20, True, width=w)
        assert cnt[k] == len(want)
        got = pts[k, :cnt[k]].copy().view(cvb.POINT_DTYPE).reshape(-1)
        _same_points(got, want)


@pytest.mark.gpu
def test_cuda_fast_errors_and_empty(cvb):
    from compv_b200 import _ffi
    d = cvb.CompVCornerDete.newObj()
    assert len(d.process(frame_const(64, 48, 0))) == 0
    assert len(d.process(frame_const(64, 48, 255))) == 0
    with pytest.raises(_ffi.CvbError) as e:   # width < 4 (fast_dete.cxx:181)
        d.process(np.zeros((10, 3), np.uint8))
    assert e.value.code == _ffi.E_INVALID_PARAMETER
    import ctypes
    assert d.set(_ffi.FAST_SET_INT_FAST_TYPE, 3, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER
    assert d.set(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, 1, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER  # wrong size
    assert d.set(9999, 1, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED
