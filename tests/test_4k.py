"""BASELINE.json's second size: 3840x2160.  Canny (+ fused Gaussian) -> KHT and FAST9 against the oracle, bit for bit, single frames and a small device batch."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_text

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
