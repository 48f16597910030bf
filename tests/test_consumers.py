"""the consumers of the PLSL / Hough outputs -- CompVConnectedComponentLabelingResultLSL::extract / boundingBoxes(segments)
(core/ccl/compv_core_ccl_lsl_result.cxx:100-230, 308-416) and CompVHough::toCartesian (houghkht.cxx:1249-1280, houghsht.cxx:566-592).
They are host arithmetic in the product (include/compv_b200.hpp, integration/compv_b200_plugin.cxx).  Here the restatements the GPU tests use are pinned on the
REFERENCE's own output through the shim (CPU), and tests/cpp/api_check (GPU) dumps what the C++ mirror computes for the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from frames import frame_g, frame_text

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def extract_restated(labels, blob):
    """Per label: every pixel (BLOB) or the {start, y}, {end, y} end points of every run (SEGMENT); rows top-down, runs left to right."""
    na = int(labels.max())
    out = [[] for _ in range(na)]
    h, w = labels.shape
    for y in range(h):
        row = labels[y]
        x = 0
        while x < w:
            a = row[x]
            if a:
                e = x
                while e < w and row[e] == a:
                    e += 1
                if blob:
                    out[a - 1].extend((xx, y) for xx in range(x, e))
                else:
                    out[a - 1].extend([(x, y), (e, y)])
                x = e
            else:
                x += 1
    return [np.array(p, np.int16).reshape(-1, 2) for p in out]


def to_cartesian_restated(kht, w, h, rho, theta):
    """float32 arithmetic in the reference's order."""
    f = np.float32
    wf, hf = f(w), f(h)
    r = np.sqrt(wf * wf + hf * hf, dtype=f)
    ox, oy = (wf * f(0.5), hf * f(0.5)) if kht else (f(0), f(0))
    out = np.zeros((len(rho), 4), f)
    for i, (rh, th) in enumerate(zip(rho.astype(f), theta.astype(f))):
        if th == 0:
            out[i] = (rh + ox, r, rh + ox, -r)
        elif kht:
            a, b = f(np.cos(th, dtype=f) * ox), f(f(1) / np.sin(th, dtype=f))
            out[i] = (0, f(f(rh + a) * b) + oy, wf, f(f(rh - a) * b) + oy)
        else:
            a, b = np.cos(th, dtype=f), f(f(1) / np.sin(th, dtype=f))
            out[i] = (0, f(rh * b), wf, f(f(rh - f(wf * a)) * b))
    return out


@needs_ref
@pytest.mark.parametrize("blob", [True, False])
def test_extract_restatement_equals_the_reference(blob):
    for img in [((frame_text(320, 200, 3) < 128) * 255).astype(np.uint8), ((frame_g(257, 130, 9) > 120) * 255).astype(np.uint8), np.zeros((40, 64), np.uint8)]:
        labels = oracle.ccl_lsl("ref", img, threads=1)["labels"]
        got, boxes = oracle.ccl_lsl_extract_ref(img, blob=blob, threads=1)
        want = extract_restated(labels, blob)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
        if not blob and len(want):
            # boundingBoxes(segments): left/top from the run starts, right/bottom from the run ENDS (exclusive column), as the reference computes them
            for a, seg in enumerate(want):
                assert tuple(boxes[a]) == (seg[0::2, 0].min(), seg[0::2, 1].min(), seg[1::2, 0].max(), seg[1::2, 1].max())


@needs_ref
@pytest.mark.parametrize("kht", [True, False])
def test_to_cartesian_restatement_close_to_the_reference(kht):
    rng = np.random.default_rng(5)
    rho = rng.uniform(-900, 900, 200).astype(np.float32)
    theta = rng.uniform(0, np.pi, 200).astype(np.float32)
    theta[::17] = 0
    got = oracle.hough_to_cartesian_ref(kht, 1920, 1080, rho, theta)
    want = to_cartesian_restated(kht, 1920, 1080, rho, theta)
    # libm's float cos/sin (the reference calls std::cos/std::sin on floats) against numpy's: identical here up to the last ulp of the products
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-3)
    np.testing.assert_array_equal(got[::17], want[::17])


@pytest.mark.gpu
def test_cpp_mirror_extract_and_to_cartesian(tmp_path):
    """The C++ mirror (include/compv_b200.hpp) on the GPU path: extract(BLOB) / extract(SEGMENT) and toCartesian against the restatements pinned above."""
    exe = os.path.join(ROOT, "tests", "cpp", "api_check")
    w, h = 640, 480
    img = frame_g(w, h, 4242)
    img.tofile(tmp_path / "f.u8")
    r = subprocess.run([exe, str(w), str(h), str(tmp_path / "f.u8"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    labels = np.fromfile(tmp_path / "plsl_labels.i32", np.int32).reshape(h, w)
    for name, blob in (("plsl_all_blobs.i32", True), ("plsl_all_segments.i32", False)):
        flat = np.fromfile(tmp_path / name, np.int32)
        want = extract_restated(labels, blob)
        i, a = 0, 0
        while i < len(flat):
            n = int(flat[i])
            np.testing.assert_array_equal(flat[i + 1:i + 1 + 2 * n].reshape(n, 2), want[a])
            i += 1 + 2 * n
            a += 1
        assert a == len(want)
    lines = np.fromfile(tmp_path / "kht_lines.bin", np.dtype([("rho", np.float32), ("theta", np.float32), ("strength", np.uint64)]))
    cart = np.fromfile(tmp_path / "kht_cartesian.f32", np.float32).reshape(-1, 6)
    want = to_cartesian_restated(True, w, h, lines["rho"], lines["theta"])
    np.testing.assert_allclose(cart[:, [0, 1, 3, 4]], want, rtol=2e-6, atol=1e-3)
    if oracle.have_ref():
        np.testing.assert_allclose(cart[:, [0, 1, 3, 4]], oracle.hough_to_cartesian_ref(True, w, h, lines["rho"], lines["theta"]), rtol=2e-6, atol=1e-3)
