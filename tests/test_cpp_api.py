"""The C++ host mirror of the reference interface (include/compv_b200.hpp): tests/cpp/api_check.cpp drives Canny -> KHT, FAST, Otsu -> PLSL, MSER and HOG through
the classes a CompV user writes against; its dumps are compared with the oracle here."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from frames import frame_g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "api_check")
LINE_DTYPE = np.dtype([("rho", np.float32), ("theta", np.float32), ("strength", np.uint64)])
POINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("strength", np.float32), ("orient", np.float32), ("level", np.int32), ("size", np.float32)])


def test_cpp_mirror_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    assert os.path.exists(EXE), "tests/cpp/api_check not built: run __graft_entry__.build()"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    frame_g(64, 48, 1).tofile(tmp_path / "f.u8")
    r = subprocess.run([EXE, "64", "48", str(tmp_path / "f.u8"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "api_check FAILED" in r.stderr and "20035" in r.stderr   # E_CUDA from CompVBase::init: no CPU path


@pytest.mark.gpu
def test_cpp_mirror_matches_the_oracle(tmp_path):
    w, h = 640, 480   # 640: CompVMat's aligned stride equals the width, the layout the reference would use for this frame
    img = frame_g(w, h, 4242)
    img.tofile(tmp_path / "f.u8")
    r = subprocess.run([EXE, str(w), str(h), str(tmp_path / "f.u8"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "api_check OK" in r.stdout
    edges = np.fromfile(tmp_path / "edges.u8", np.uint8).reshape(h, w)
    want_edges = oracle.edge_dete("orc", img, "canny", 59.0, 119.0, 3)
    np.testing.assert_array_equal(edges, want_edges)
    lines = np.fromfile(tmp_path / "kht_lines.bin", LINE_DTYPE)
    want_lines, _ = oracle.hough_kht("orc", want_edges, 1.0, 1.0, 50, max_lines=20)
    assert len(lines) == len(want_lines) == 20
    for k in ("rho", "theta", "strength"):
        np.testing.assert_array_equal(lines[k], want_lines[k])
    cart = np.fromfile(tmp_path / "kht_cartesian.f32", np.float32).reshape(-1, 6)
    assert len(cart) == len(lines)
    for (ax, ay, az, bx, by, bz), ln in zip(cart, lines):           # houghkht.cxx:1249-1280, restated in float32
        rho, theta = np.float32(ln["rho"]), np.float32(ln["theta"])
        assert az == 1 and bz == 1
        if theta == 0:
            assert ax == bx == rho + np.float32(w) * np.float32(0.5)
        else:
            a, b = np.cos(np.float64(theta)) * (w * 0.5), 1.0 / np.sin(np.float64(theta))
            tol = 1e-5 * (abs(float(rho)) + abs(a)) * abs(b) + 1e-2     # float32 rounding of rho +- a, amplified by 1 / sin(theta)
            assert ax == 0 and bx == w
            assert abs(ay - ((rho + a) * b + h * 0.5)) <= tol and abs(by - ((rho - a) * b + h * 0.5)) <= tol
    pts = np.fromfile(tmp_path / "fast_points.bin", POINT_DTYPE)
    want_pts = oracle.fast_detect("orc", img, 9, 20, True)
    assert len(pts) == len(want_pts) > 0
    for k in ("x", "y", "strength"):
        np.testing.assert_array_equal(pts[k], want_pts[k])
    otsu = np.fromfile(tmp_path / "otsu.u8", np.uint8).reshape(h, w)
    want_otsu, _ = oracle.threshold("orc", "otsu", img)
    np.testing.assert_array_equal(otsu, want_otsu)
    labels = np.fromfile(tmp_path / "plsl_labels.i32", np.int32).reshape(h, w)
    np.testing.assert_array_equal(labels, oracle.ccl_lsl("orc", otsu)["labels"])
    blobs = np.fromfile(tmp_path / "plsl_blobs.i32", np.int32)
    i = 0
    while i < len(blobs):                                            # extract(BLOB): every pixel of the label, rows top-down, left to right
        a, n = int(blobs[i]), int(blobs[i + 1])
        xy = blobs[i + 2:i + 2 + 2 * n].reshape(n, 2)
        ys, xs = np.nonzero(labels == a)
        np.testing.assert_array_equal(xy, np.stack([xs, ys], 1))
        i += 2 + 2 * n
    closed = np.fromfile(tmp_path / "closed.u8", np.uint8).reshape(h, w)
    np.testing.assert_array_equal(closed, oracle.morph("orc", otsu, oracle.morph_strel("orc", (3, 3), oracle.STREL_RECT), oracle.MORPH_CLOSE))
    sizes = np.fromfile(tmp_path / "mser_sizes.i32", np.int32)
    want = oracle.ccl_lmser("orc", img)
    np.testing.assert_array_equal(np.sort(sizes), np.sort(want["sizes"]))
    hog = np.fromfile(tmp_path / "hog.f32", np.float32)
    np.testing.assert_array_equal(hog, oracle.hog("orc", img))
