"""The seeded cases behind tests/golden/reference_md5.json: one function per row returning the bytes whose MD5 is pinned.
`impl` is a small adapter so that the same case runs on the compiled reference (the generator), on the oracle (CPU test) and on the CUDA library (GPU test)."""
import hashlib

import numpy as np

from frames import frame_g, frame_text, frame_uniform

W, H = 640, 360


def md5(*arrays):
    m = hashlib.md5()
    for a in arrays:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()


def lines_bytes(lines):
    return [np.asarray(lines["rho"], np.float32), np.asarray(lines["theta"], np.float32), np.asarray(lines["strength"], np.uint64)]


def cases(impl):
    g, t, u = frame_g(W, H, 2024), frame_text(W, H, 11), frame_uniform(W, H, 3)
    binar = ((t < 128) * 255).astype(np.uint8)
    out = {}
    k5 = impl.gauss_kernel(5, 1.0)
    blurred = impl.convlt("8u32f8u", g, k5, k5)
    out["a2_gauss5_u8"] = md5(blurred)
    out["a3_sobel"] = md5(impl.edge(g, "sobel", 0.0, 0.0))
    edges = impl.edge(blurred, "canny", 59.0, 119.0)
    out["a5_canny_blur5"] = md5(edges)
    out["a5_canny_uniform"] = md5(impl.edge(u, "canny", 59.0, 119.0))
    out["a6_sht_thr60"] = md5(*lines_bytes(impl.sht(edges, 60)))
    out["a7_kht_thr30"] = md5(*lines_bytes(impl.kht(edges, 30)))
    p = impl.fast(g, 9, 20)
    out["a8_fast9_t20"] = md5(np.asarray(p["x"], np.float32), np.asarray(p["y"], np.float32), np.asarray(p["strength"], np.float32))
    p = impl.fast(u, 12, 40)
    out["a8_fast12_t40_uniform"] = md5(np.asarray(p["x"], np.float32), np.asarray(p["y"], np.float32), np.asarray(p["strength"], np.float32))
    o, thr = impl.otsu(g)
    out["a10_otsu"] = md5(o, np.array([thr], np.float64))
    out["a10_adaptive5"] = md5(impl.adaptive(t))
    r = impl.lsl(binar)
    out["a11_plsl_text"] = md5(r["labels"], r["boxes"], np.array([r["na"]], np.int32))
    r = impl.lsl(((u > 120) * 255).astype(np.uint8))
    out["a11_plsl_noise"] = md5(r["labels"], r["boxes"], np.array([r["na"]], np.int32))
    out["a12_mser_sizes_boxes"] = md5(*impl.mser_canonical(g))
    se = np.ones((3, 3), np.uint8) * 255
    out["8f1_close3"] = md5(impl.morph(binar, se, 3))
    out["8f1_erode_cross5"] = md5(impl.morph(g, impl.strel((5, 5), 2), 0))
    return out
