"""a6 HoughSHT: oracle pinned on the compiled reference (CPU, bit-exact lines in the reference's order); CUDA vs oracle / reference (GPU)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_smooth, frame_text
from test_kht import canny_edges, edge_maps, same_lines

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("w,h", [(64, 48), (320, 200), (641, 479)])
@pytest.mark.parametrize("threshold", [1, 30, 100])
@pytest.mark.parametrize("simd", [True, False])
def test_oracle_sht_vs_reference(w, h, threshold, simd):
    for e in edge_maps(w, h):
        a, na = oracle.hough_sht("orc", e, 1.0, 1.0, threshold, x86_simd=simd, cap=1 << 20)
        r, nr = oracle.hough_sht("ref", e, 1.0, 1.0, threshold, threads=1, x86_simd=simd, cap=1 << 20)
        assert na == nr
        same_lines(a, r)


@needs_ref
def test_oracle_sht_1080p_vs_reference():
    e = canny_edges(frame_g(1920, 1080, 4242))
    a, na = oracle.hough_sht("orc", e, 1.0, 1.0, 150)
    r, nr = oracle.hough_sht("ref", e, 1.0, 1.0, 150, threads=1)
    assert na == nr and na > 0
    same_lines(a, r)


@needs_ref
@pytest.mark.parametrize("kw", [dict(theta=0.5), dict(theta=2.0, max_lines=7), dict(theta=3.0), dict(theta=0.7), dict(max_lines=1)])
def test_oracle_sht_parameters_vs_reference(kw):
    e = canny_edges(frame_text(400, 300, 5))
    a, na = oracle.hough_sht("orc", e, threshold=40, **kw)
    r, nr = oracle.hough_sht("ref", e, threshold=40, threads=1, **kw)
    assert na == nr
    same_lines(a, r)


@needs_ref
def test_oracle_sht_multithreaded_reference_agrees():
    # the reference's MT path sums per-thread accumulators and concatenates per-thread line lists in row order: same result as ST
    e = canny_edges(frame_smooth(640, 480, 9))
    a, na = oracle.hough_sht("orc", e, threshold=60)
    r, nr = oracle.hough_sht("ref", e, threshold=60, threads=-1)
    assert na == nr
    same_lines(a, r)


@needs_ref
def test_oracle_sht_empty_and_full():
    z = np.zeros((40, 64), np.uint8)
    a, na = oracle.hough_sht("orc", z, threshold=1)
    r, nr = oracle.hough_sht("ref", z, threshold=1, threads=1)
    assert na == nr == 0
    f = np.full((24, 32), 255, np.uint8)
    a, na = oracle.hough_sht("orc", f, threshold=5, cap=1 << 20)
    r, nr = oracle.hough_sht("ref", f, threshold=5, threads=1, cap=1 << 20)
    assert na == nr
    same_lines(a, r)


def test_oracle_sht_rejects_fractional_rho():
    z = np.zeros((16, 16), np.uint8)
    with pytest.raises(Exception):
        oracle.hough_sht("orc", z, rho=0.5)


# ---------------------------------------------------------------- CUDA (C ABI) vs oracle / reference
@pytest.mark.gpu
# threshold 1 at 1080p would be ~1M lines per frame: that threshold is covered at the three smaller sizes
@pytest.mark.parametrize("w,h,threshold", [(w, h, t) for (w, h) in [(64, 48), (320, 200), (641, 479), (1920, 1080)] for t in [1, 30, 100] if not ((w, h) == (1920, 1080) and t == 1)])
@pytest.mark.parametrize("simd", [True, False])
def test_cuda_sht(cvb, w, h, threshold, simd):
    from compv_b200 import _ffi
    d = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, 1.0, threshold)
    d.setBool(cvb.CompVHough.HOUGH_SET_BOOL_X86_SIMD_SCAN, simd)
    for e in edge_maps(w, h):
        a = d.process(e, capacity=1 << 20)
        o, n = oracle.hough_sht("orc", e, 1.0, 1.0, threshold, x86_simd=simd, cap=1 << 20)
        assert n == len(o)
        same_lines(a, o)
        if oracle.have_ref() and (w, h) != (1920, 1080):
            r, _ = oracle.hough_sht("ref", e, 1.0, 1.0, threshold, threads=1, x86_simd=simd, cap=1 << 20)
            same_lines(a, r)


@pytest.mark.gpu
def test_cuda_sht_parameters_and_caps(cvb):
    import ctypes
    from compv_b200 import _ffi
    e = canny_edges(frame_text(400, 300, 5))
    for kw in [dict(theta=0.5), dict(theta=2.0, max_lines=7), dict(theta=3.0), dict(theta=0.7), dict(max_lines=1)]:
        d = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, kw.get("theta", 1.0), 40)
        if "max_lines" in kw:
            d.setInt(_ffi.HOUGH_SET_INT_MAXLINES, kw["max_lines"])
        o, _ = oracle.hough_sht("orc", e, threshold=40, **kw)
        same_lines(d.process(e), o)
    # houghsht.cxx:310-314 / :69-72: the SHT only takes rho == 1
    h = ctypes.c_void_p()
    assert cvb.lib().cvb200_hough_new(ctypes.byref(h), _ffi.HOUGHSHT_ID, ctypes.c_float(0.5), ctypes.c_float(1.0), ctypes.c_size_t(1)) == _ffi.E_INVALID_PARAMETER
    d = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID)
    assert d.set(_ffi.HOUGH_SET_FLT32_RHO, 0.5, ctypes.c_float) == _ffi.E_INVALID_PARAMETER
    assert d.set(_ffi.HOUGH_SET_FLT32_RHO, 1.0, ctypes.c_float) == _ffi.S_OK
    assert d.set(_ffi.HOUGH_SET_INT_THRESHOLD, 0, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER     # :81
    assert d.set(_ffi.HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE, 5, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED  # :88-91: unknown id for the SHT
    # empty / all-set frames, capacity smaller than the number of lines
    z = np.zeros((40, 64), np.uint8)
    assert len(d.process(z)) == 0
    f = np.full((24, 32), 255, np.uint8)
    d.setInt(_ffi.HOUGH_SET_INT_THRESHOLD, 5)
    o, n = oracle.hough_sht("orc", f, threshold=5, cap=1 << 20)
    same_lines(d.process(f, capacity=1 << 20), o)
    same_lines(d.process(f, capacity=10), o[:10])


@pytest.mark.gpu
def test_cuda_canny_then_sht_batched_on_device(cvb):
    """Fused Gaussian+Canny on the device, SHT on the device edge maps; the batch spans several L2-sized chunks and an unaligned pitch."""
    import torch
    from compv_b200 import _ffi
    w, h, batch = 1280, 720, 27
    frames = np.stack([frame_g(w, h, 999 + k) if k % 3 else frame_text(w, h, k) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    d_edges = torch.empty_like(d_in)
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    stream = torch.cuda.current_stream().cuda_stream
    canny.process_dev(d_in, w, h, w, d_edges, batch=batch, stream=stream)
    sht = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, 1.0, 120)
    got = sht.process_dev(d_edges, w, h, w, batch=batch, stream=stream, capacity=1 << 16)
    edges = d_edges.cpu().numpy()
    for k in range(batch):
        want, _ = oracle.hough_sht("orc", edges[k], 1.0, 1.0, 120)
        same_lines(got[k], want)
    # odd width + stride not a multiple of 4 -> the byte-load variant of the list kernel
    e = np.zeros((57, 131), np.uint8)
    e[:, :127] = (canny_edges(frame_g(127, 57, 5)) != 0) * 255
    sht2 = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, 1.0, 10)
    want, _ = oracle.hough_sht("orc", e, 1.0, 1.0, 10, width=127)
    same_lines(sht2.process(e, width=127), want)
