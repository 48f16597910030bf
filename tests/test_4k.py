"""BASELINE.json's second size: 3840x2160.  Canny (+ fused Gaussian) -> KHT and FAST9 against the oracle, bit for bit, single frames and a small device batch."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_text, frame_smooth

W, H = 3840, 2160


def want_edges(img):
    k = oracle.gauss_kernel("orc", 5, 1.0)
    return oracle.edge_dete("orc", oracle.convlt1("orc", "8u32f8u", img, k, k), "canny", 59.0, 119.0, 3)


@pytest.mark.gpu
def test_cuda_canny_kht_4k(cvb):
    import torch
    from compv_b200 import _ffi
    frames = np.stack([frame_g(W, H, 77), frame_text(W, H, 5)])
    d_in = torch.from_numpy(frames).cuda()
    d_edges = torch.empty_like(d_in)
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    stream = torch.cuda.current_stream().cuda_stream
    canny.process_dev(d_in, W, H, W, d_edges, batch=2, stream=stream)
    kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 100)
    got = kht.process_dev(d_edges, W, H, W, batch=2, stream=stream)
    edges = d_edges.cpu().numpy()
    for k in range(2):
        e = want_edges(frames[k])
        np.testing.assert_array_equal(edges[k], e)
        want, _ = oracle.hough_kht("orc", e, 1.0, 1.0, 100)
        assert len(got[k]) == len(want)
        for key in ("rho", "theta", "strength"):
            np.testing.assert_array_equal(got[k][key], want[key])
    # the host-buffer pipeline on the same frames
    lines = cvb.canny_kht_process_batch(canny, kht, frames, width=W)
    for k in range(2):
        for key in ("rho", "theta", "strength"):
            np.testing.assert_array_equal(lines[k][key], got[k][key])


@pytest.mark.gpu
def test_cuda_fast_4k(cvb):
    from compv_b200 import _ffi
    img = frame_g(W, H, 78)
    fast = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
    fast.setInt(_ffi.FAST_SET_INT_THRESHOLD, 20)
    fast.setInt(_ffi.FAST_SET_INT_MAX_FEATURES, -1)
    fast.setBool(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, True)
    got = fast.process(img)
    want = oracle.fast_detect("orc", img, 9, 20, True)
    assert len(got) == len(want) > 0
    for key in ("x", "y", "strength"):
        np.testing.assert_array_equal(got[key], want[key])


# ---- BASELINE config 5 at its real size: PLSL and LMSER at 3840x2160 (and LMSER at 1920x1080), incl. the multi-pass chunking of LMSER batches ----
def _lsl_check(res, want):
    assert res.labelsCount() == want["na"]
    np.testing.assert_array_equal(res.debugFlatten(), want["labels"])
    np.testing.assert_array_equal(res.boundingBoxes(), want["boxes"])
    ro, rg = res.segments()
    np.testing.assert_array_equal(ro, want["row_offsets"])
    for k in ("a", "start", "end"):
        np.testing.assert_array_equal(rg[k], want["ranges"][k])


def _mser_canonical(r):
    s = set()
    for n, b, p in zip(r["sizes"], r["boxes"], r["points"]):
        s.add((int(n), tuple(int(v) for v in b), np.sort(p[:, 1].astype(np.int64) * 65536 + p[:, 0]).tobytes()))
    return s


@pytest.mark.gpu
def test_cuda_plsl_4k(cvb):
    import torch
    from compv_b200 import _ffi
    frames = np.stack([((frame_text(W, H, 21) < 128) * 255).astype(np.uint8), ((frame_g(W, H, 5) > 120) * 255).astype(np.uint8), ((frame_text(W, H, 22) < 128) * 255).astype(np.uint8)])
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    want = [oracle.ccl_lsl("orc", f) for f in frames]
    _lsl_check(ccl.process(frames[0]), want[0])                     # host entry point, one frame
    if oracle.have_ref():
        r = oracle.ccl_lsl("ref", frames[1], threads=1)
        assert r["na"] == want[1]["na"]
        np.testing.assert_array_equal(r["labels"], want[1]["labels"])
    d_in = torch.from_numpy(frames).cuda()                          # device batch of differing frames, label images written on the device
    d_labels = torch.empty((3, H, W), dtype=torch.int32, device="cuda")
    na, results = ccl.process_dev(d_in, W, H, W, batch=3, d_labels=d_labels, want_results=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    labels = d_labels.cpu().numpy()
    for k in range(3):
        assert na[k] == want[k]["na"]
        np.testing.assert_array_equal(labels[k], want[k]["labels"])
        _lsl_check(results[k], want[k])


@pytest.mark.gpu
def test_cuda_lmser_4k_multi_pass(cvb):
    """Four differing 3840x2160 frames: more than one pass of the <= 3 GB per pass chunking (ccl_lmser.cu), regions compared as sets with the oracle."""
    import torch
    from compv_b200 import _ffi
    kw = dict(delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8)
    frames = np.stack([frame_g(W, H, 31), frame_smooth(W, H, 8), frame_g(W, H, 32), frame_smooth(W, H, 9)])
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, **kw)
    d_in = torch.from_numpy(frames).cuda()
    na, results = ccl.process_dev(d_in, W, H, W, batch=4, want_results=True, stream=torch.cuda.current_stream().cuda_stream)
    for k in range(4):
        want = oracle.ccl_lmser("orc", frames[k], **kw)
        assert na[k] == len(want["sizes"]) > 0
        assert _mser_canonical(results[k].regions()) == _mser_canonical(want)
    one = ccl.process(frames[1])                                    # host entry point, single frame
    assert _mser_canonical(one.regions()) == _mser_canonical(oracle.ccl_lmser("orc", frames[1], **kw))


@pytest.mark.gpu
def test_cuda_lmser_1080p_batch_over_one_pass(cvb):
    """Seventeen 1920x1080 frames (15 fit one pass): two passes, differing frames."""
    import torch
    from compv_b200 import _ffi
    w, h, batch = 1920, 1080, 17
    kw = dict(delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8)
    distinct = [frame_g(w, h, 40), frame_smooth(w, h, 11), frame_g(w, h, 41)]
    frames = np.stack([distinct[k % 3] for k in range(batch)])
    want = [oracle.ccl_lmser("orc", f, **kw) for f in distinct]
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, **kw)
    na, results = ccl.process_dev(torch.from_numpy(frames).cuda(), w, h, w, batch=batch, want_results=True, stream=torch.cuda.current_stream().cuda_stream)
    for k in range(batch):
        assert na[k] == len(want[k % 3]["sizes"])
        if k in (0, 1, 2, 14, 15, 16):
            assert _mser_canonical(results[k].regions()) == _mser_canonical(want[k % 3])
