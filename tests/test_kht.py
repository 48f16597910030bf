"""a7 HoughKHT: oracle pinned on the compiled reference (CPU, bit-exact lines and Gs); CUDA vs oracle / reference (GPU)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform, frame_smooth, frame_text, frame_const

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


def canny_edges(img, blur=True):
    if blur:
        k = oracle.gauss_kernel("orc", 5, 1.0)
        img = oracle.convlt1("orc", "8u32f8u", img, k, k)
    return oracle.edge_dete("orc", img, "canny", 59.0, 119.0, 3)


def edge_maps(w, h):
    maps = [canny_edges(frame_g(w, h, 12345)), canny_edges(frame_smooth(w, h, 3)), canny_edges(frame_text(w, h, 7)), canny_edges(frame_uniform(w, h, 1), blur=False)]
    lines = np.zeros((h, w), np.uint8)          # a few exact straight lines, one touching the border, plus an isolated short one (< min size)
    lines[h // 3, 5:w - 5] = 255
    lines[10:h - 10, w // 4] = 255
    for i in range(min(w, h) - 20):
        lines[10 + i, 10 + i] = 255
    lines[0, 3:40] = 255
    lines[h - 5, 50:56] = 255
    maps.append(lines)
    return maps


def same_lines(a, b):
    assert len(a) == len(b)
    np.testing.assert_array_equal(a["rho"], b["rho"])
    np.testing.assert_array_equal(a["theta"], b["theta"])
    np.testing.assert_array_equal(a["strength"], b["strength"])


@needs_ref
@pytest.mark.parametrize("w,h", [(64, 48), (320, 200), (640, 480), (1920, 1080)])
@pytest.mark.parametrize("threshold", [1, 100])
def test_oracle_kht_vs_reference(w, h, threshold):
    for e in edge_maps(w, h):
        a, gsa = oracle.hough_kht("orc", e, 1.0, 1.0, threshold)
        r, gsr = oracle.hough_kht("ref", e, 1.0, 1.0, threshold, threads=1)
        same_lines(a, r)
        assert gsa == gsr


@needs_ref
def test_oracle_kht_parameters_vs_reference():
    e = canny_edges(frame_g(640, 480, 777))
    for kw in [dict(rho=0.5, theta=0.5), dict(rho=1.0, theta=2.0, max_lines=5), dict(cluster_min_deviation=1.0, cluster_min_size=6), dict(kernel_min_height=0.05)]:
        a, gsa = oracle.hough_kht("orc", e, threshold=10, **kw)
        r, gsr = oracle.hough_kht("ref", e, threshold=10, threads=1, **kw)
        same_lines(a, r)
        assert gsa == gsr


@needs_ref
def test_oracle_kht_empty_inputs():
    for e in [frame_const(64, 48, 0), frame_const(64, 48, 255)]:
        a, _ = oracle.hough_kht("orc", e)
        r, _ = oracle.hough_kht("ref", e, threads=1)
        same_lines(a, r)


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 48), (320, 200), (640, 480), (1920, 1080)])
@pytest.mark.parametrize("threshold", [1, 100])
def test_cuda_kht(cvb, w, h, threshold):
    from compv_b200 import _ffi
    d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, threshold)
    for e in edge_maps(w, h):
        a = d.process(e)
        o, gs = oracle.hough_kht("orc", e, 1.0, 1.0, threshold)
        same_lines(a, o)
        if len(o):
            assert d.getFloat64(_ffi.HOUGHKHT_GET_FLT64_GS) == gs
        if oracle.have_ref():
            r, _ = oracle.hough_kht("ref", e, 1.0, 1.0, threshold, threads=1)
            same_lines(a, r)


@pytest.mark.gpu
def test_cuda_kht_parameters_and_caps(cvb):
    import ctypes
    from compv_b200 import _ffi
    e = canny_edges(frame_g(640, 480, 777))
    for kw in [dict(rho=0.5, theta=0.5), dict(rho=1.0, theta=2.0, max_lines=5), dict(cluster_min_deviation=1.0, cluster_min_size=6), dict(kernel_min_height=0.05)]:
        d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, kw.get("rho", 1.0), kw.get("theta", 1.0), 10)
        if "max_lines" in kw:
            d.setInt(_ffi.HOUGH_SET_INT_MAXLINES, kw["max_lines"])
        if "cluster_min_deviation" in kw:
            d.setFloat32(_ffi.HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION, kw["cluster_min_deviation"])
            d.setInt(_ffi.HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE, kw["cluster_min_size"])
        if "kernel_min_height" in kw:
            d.setFloat32(_ffi.HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH, kw["kernel_min_height"])
        o, _ = oracle.hough_kht("orc", e, threshold=10, **kw)
        same_lines(d.process(e), o)
    d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID)
    assert d.set(_ffi.HOUGH_SET_FLT32_RHO, 2.0, ctypes.c_float) == _ffi.E_INVALID_PARAMETER      # houghkht.cxx:144
    assert d.set(_ffi.HOUGH_SET_INT_THRESHOLD, 0, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER    # houghkht.cxx:156
    assert d.set(_ffi.HOUGH_SET_INT_MAXLINES, 1, ctypes.c_float) == _ffi.S_OK                    # only the size is checked (houghkht.cxx:162)
    assert d.set(9999, 1, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED
    for e0 in [frame_const(64, 48, 0), frame_const(64, 48, 255)]:
        same_lines(d.process(e0), oracle.hough_kht("orc", e0)[0])


@pytest.mark.gpu
def test_cuda_canny_then_kht_batched_on_device(cvb):
    """The headline pipeline: fused Gaussian+Canny on the device, KHT on the device edge maps, whole batch per call."""
    import torch
    from compv_b200 import _ffi
    w, h, batch = 1920, 1080, 3
    frames = np.stack([frame_g(w, h, 12345 + k) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    d_edges = torch.empty_like(d_in)
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    stream = torch.cuda.current_stream().cuda_stream
    canny.process_dev(d_in, w, h, w, d_edges, batch=batch, stream=stream)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 100)
    got = kht.process_dev(d_edges, w, h, w, batch=batch, stream=stream)
    for k in range(batch):
        want, _ = oracle.hough_kht("orc", canny_edges(frames[k]), 1.0, 1.0, 100)
        same_lines(got[k], want)


@pytest.mark.gpu
@pytest.mark.parametrize("lanes_min", [1, 1 << 30])
def test_cuda_kht_both_linking_kernels(cvb, lanes_min, monkeypatch):
    """The linking stage has two kernels: one WARP per frame (few frames: the lanes share the seed scan) and one LANE per frame (kht_link_lanes_kernel, 32 frames per warp,
    taken from CVB200_KHT_LANES_MIN frames per launch on).  Both must give the oracle's lines for every frame of a batch whose frames differ widely (empty, full,
    text, long strings) and whose size is not a multiple of 32."""
    import torch
    from compv_b200 import _ffi
    monkeypatch.setenv("CVB200_KHT_LANES_MIN", str(lanes_min))
    w, h = 320, 200
    maps = edge_maps(w, h) + [serpentine(w, h), frame_const(w, h, 0), frame_const(w, h, 255)]
    maps += [canny_edges(frame_g(w, h, 4000 + k)) for k in range(70 - len(maps))]
    maps = [maps[(7 * k) % len(maps)] for k in range(len(maps))]   # neighbours in a warp differ
    batch = len(maps)
    d_edges = torch.from_numpy(np.stack(maps)).cuda()
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    for _ in range(2):
        got = kht.process_dev(d_edges, w, h, w, batch=batch, stream=torch.cuda.current_stream().cuda_stream)
        for k in range(batch):
            want, _ = oracle.hough_kht("orc", maps[k], 1.0, 1.0, 30)
            same_lines(got[k], want)
    # the per-frame host API with the lane kernel forced (a batch of one)
    d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 100)
    for e in edge_maps(640, 480):
        same_lines(d.process(e), oracle.hough_kht("orc", e, 1.0, 1.0, 100)[0])


def serpentine(w, h):
    """One 8-connected string several thousand pixels long (longer than the linking kernel's shared-memory stage)."""
    e = np.zeros((h, w), np.uint8)
    left = True
    for y in range(10, h - 10, 4):
        e[y, 20:w - 20] = 255
        x = w - 21 if left else 20
        if y + 4 < h - 10:
            e[y:y + 4, x] = 255
        left = not left
    return e


@needs_ref
def test_oracle_kht_long_string_vs_reference():
    e = serpentine(320, 200)
    assert (e != 0).sum() > 4096
    a, gsa = oracle.hough_kht("orc", e, 1.0, 1.0, 10)
    r, gsr = oracle.hough_kht("ref", e, 1.0, 1.0, 10, threads=1)
    same_lines(a, r)
    assert gsa == gsr


@pytest.mark.gpu
def test_cuda_kht_long_string(cvb):
    from compv_b200 import _ffi
    d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 10)
    for (w, h) in [(320, 200), (640, 480)]:
        e = serpentine(w, h)
        o, _ = oracle.hough_kht("orc", e, 1.0, 1.0, 10)
        same_lines(d.process(e), o)


@pytest.mark.gpu
def test_cuda_canny_kht_host_batch_many_chunks(cvb):
    """cvb200_canny_kht_process_batch on a batch that spans many upload chunks (and is not a multiple of the chunk size); called twice: cached streams and buffers are reused."""
    from compv_b200 import _ffi
    w, h, batch = 320, 200, 150
    frames = np.stack([frame_g(w, h, 500 + k) if k % 4 else frame_text(w, h, k) for k in range(batch)])
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    for _ in range(2):
        got = cvb.canny_kht_process_batch(canny, kht, frames, width=w)
        gs_last = None
        for k in range(batch):
            want, gs_last = oracle.hough_kht("orc", canny_edges(frames[k]), 1.0, 1.0, 30)
            same_lines(got[k], want)
        assert kht.getFloat64(_ffi.HOUGHKHT_GET_FLT64_GS) == gs_last


@pytest.mark.gpu
@pytest.mark.parametrize("sub,slots", [(4, 2), (7, 3), (256, 6)])
def test_cuda_canny_kht_device_pipeline_ring(cvb, sub, slots, monkeypatch):
    """cvb200_canny_kht_process_batch_dev: sub-batches on a ring of slots (each slot a private copy of the two detectors), slots reused several times,
    last sub-batch ragged; every frame's lines must be the oracle's, whatever the cut."""
    import torch
    from compv_b200 import _ffi
    monkeypatch.setenv("CVB200_PIPE_SUB", str(sub))
    monkeypatch.setenv("CVB200_PIPE_SLOTS", str(slots))
    w, h, batch = 320, 200, 23
    frames = np.stack([frame_g(w, h, 900 + k) if k % 3 else frame_smooth(w, h, k) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        got = cvb.canny_kht_process_batch_dev(canny, kht, d_in, w, h, w, batch, stream=stream)
        gs_last = None
        for k in range(batch):
            want, gs_last = oracle.hough_kht("orc", canny_edges(frames[k]), 1.0, 1.0, 30)
            same_lines(got[k], want)
        assert kht.getFloat64(_ffi.HOUGHKHT_GET_FLT64_GS) == gs_last


@pytest.mark.gpu
@pytest.mark.parametrize("sub,slots", [(4, 4), (7, 3), (5, 8)])
def test_cuda_canny_kht_host_pipeline_ring_rotation(cvb, sub, slots, monkeypatch):
    """cvb200_canny_kht_process_batch (host frames): the ring is rotated so that the LAST sub-batch runs on the last slot (the highest-priority stream), whatever the
    number of sub-batches; slots are reused while earlier sub-batches are still in flight.  Every frame's lines must be the oracle's, in frame order."""
    from compv_b200 import _ffi
    monkeypatch.setenv("CVB200_PIPE_SUB", str(sub))
    monkeypatch.setenv("CVB200_PIPE_SLOTS", str(slots))
    w, h, batch = 320, 200, 23
    frames = np.stack([frame_g(w, h, 1300 + k) if k % 3 else frame_text(w, h, k) for k in range(batch)])
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    for _ in range(2):
        got = cvb.canny_kht_process_batch(canny, kht, frames, width=w)
        gs_last = None
        for k in range(batch):
            want, gs_last = oracle.hough_kht("orc", canny_edges(frames[k]), 1.0, 1.0, 30)
            same_lines(got[k], want)
        assert kht.getFloat64(_ffi.HOUGHKHT_GET_FLT64_GS) == gs_last


@pytest.mark.gpu
def test_cuda_kht_pools_grow_on_overflow(cvb):
    """The pools are sized from earlier calls: a first small call followed by a dense frame (every pool too small) must still give the oracle's lines."""
    from compv_b200 import _ffi
    d = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 1)
    small = np.zeros((48, 64), np.uint8); small[20, 5:40] = 255
    same_lines(d.process(small), oracle.hough_kht("orc", small, 1.0, 1.0, 1)[0])
    dense = canny_edges(frame_uniform(640, 480, 5), blur=False)
    same_lines(d.process(dense), oracle.hough_kht("orc", dense, 1.0, 1.0, 1)[0])


@pytest.mark.gpu
def test_cuda_canny_hysteresis_needs_many_rounds(cvb):
    """A weak edge snaking through many 64x64 tiles with a single strong pixel at one end: the closure needs more list-driven rounds than one call issues,
    the detector must notice, raise its round count and still return the reference's edge map."""
    from compv_b200 import _ffi
    w, h = 1024, 320
    img = np.full((h, w), 20, np.uint8)
    # a low-contrast ridge (weak after NMS) across the whole width, alternating rows joined at the ends ...
    rows = list(range(20, h - 20, 24))
    for i, r in enumerate(rows):
        img[r, 10:w - 10] = 60
        x = w - 11 if i % 2 == 0 else 10
        if i + 1 < len(rows):
            img[r:rows[i + 1] + 1, x] = 60
    img[rows[0], 10:14] = 255              # ... and one strong seed at its beginning
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 30.0, 400.0, 3)
    got = canny.process(img)
    want = oracle.edge_dete("orc", img, "canny", 30.0, 400.0, 3)
    assert np.array_equal(got, want)
    assert (want == 255).sum() > 2000       # the closure really travelled


@pytest.mark.gpu
def test_cuda_canny_kht_multi_device_in_process(cvb):
    """cvb200_init_devices + cvb200_canny_kht_process_batch_multi: the batch is cut into one shard per device, each shard runs on its device from a worker thread
    of THIS process (no torchrun).  With one GPU on the box this is one shard through the same worker path; with more (gpurun --gpus N) every device takes part."""
    from compv_b200 import _ffi
    n = cvb.init_devices(0)
    assert n >= 1
    w, h, batch = 320, 200, 4 * n + 3
    frames = np.stack([frame_g(w, h, 1200 + k) if k % 2 else frame_smooth(w, h, 50 + k) for k in range(batch)])
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30)
    for _ in range(2):
        got = cvb.canny_kht_process_batch_multi(canny, kht, frames, width=w)
        for k in range(batch):
            want, gs = oracle.hough_kht("orc", canny_edges(frames[k]), 1.0, 1.0, 30)
            same_lines(got[k], want)
        assert kht.getFloat64(_ffi.HOUGHKHT_GET_FLT64_GS) == gs
    # the single-device entry points still work afterwards (device 0)
    same_lines(kht.process(canny_edges(frames[0])), oracle.hough_kht("orc", canny_edges(frames[0]), 1.0, 1.0, 30)[0])
