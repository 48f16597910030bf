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
