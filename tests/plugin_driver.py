"""Run in a subprocess by tests/test_plugin.py (registering the B200 factories changes what the reference's factory hands out for the whole process).
Loads the compiled reference + shim, then integration/compv_b200_plugin.cxx's library, registers, and drives the REFERENCE's public API through the shim."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from frames import frame_g  # noqa: E402

out = {}
oracle.ref(1)                                   # CompVBase::init + CompVCore::init: the reference registers its own CPU factories
plugin = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcompv_b200_plugin.so"), mode=ctypes.RTLD_GLOBAL)
out["register_rc"] = int(plugin.compv_b200_register(0))
b200 = ctypes.CDLL(os.path.join(ROOT, "compv_b200", "lib", "libcompv_b200.so"))
b200.cvb200_launch_count.restype = ctypes.c_uint64
launches0 = int(b200.cvb200_launch_count())

img = frame_g(640, 480, 31)
# the calls below go: shim -> CompVEdgeDete::newObj / CompVHough::newObj / CompVCornerDete::newObj (the reference's factory) -> whatever is registered
edges = oracle.edge_dete("ref", img, "canny", 59.0, 119.0, 3, threads=1)
want_edges = oracle.edge_dete("orc", img, "canny", 59.0, 119.0, 3)
out["canny_equal"] = bool(np.array_equal(edges, want_edges))
lines, gs = oracle.hough_kht("ref", want_edges, 1.0, 1.0, 30, threads=1)
wl, wgs = oracle.hough_kht("orc", want_edges, 1.0, 1.0, 30)
out["kht_equal"] = bool(len(lines) == len(wl) and all(np.array_equal(lines[k], wl[k]) for k in ("rho", "theta", "strength")) and gs == wgs)
sl, _ = oracle.hough_sht("ref", want_edges, 1.0, 1.0, 60, threads=1)
wsl, _ = oracle.hough_sht("orc", want_edges, 1.0, 1.0, 60)
out["sht_equal"] = bool(len(sl) == len(wsl) and all(np.array_equal(sl[k], wsl[k]) for k in ("rho", "theta", "strength")))
pts = oracle.fast_detect("ref", img, 9, 20, True, threads=1)
wp = oracle.fast_detect("orc", img, 9, 20, True)
out["fast_equal"] = bool(len(pts) == len(wp) and all(np.array_equal(pts[k], wp[k]) for k in ("x", "y", "strength")))
sob = oracle.edge_dete("ref", img, "sobel", 0.0, 0.0, 3, threads=1)
out["sobel_equal"] = bool(np.array_equal(sob, oracle.edge_dete("orc", img, "sobel", 0.0, 0.0, 3)))          # B200 default = true max (C++ path)
out["sobel_equal_x86_quirk"] = bool(np.array_equal(sob, oracle.edge_dete("orc", img, "sobel", 0.0, 0.0, 3, sse41_gmax_lanes=True)))
# HOG and the two CCL algorithms go through CompVHOG::newObj / CompVConnectedComponentLabeling::newObj (its own factory map, base/compv_ccl.cxx:42-52)
hog = oracle.hog("ref", img, threads=1)
whog = oracle.hog("orc", img)
out["hog_size_equal"] = bool(hog.shape == whog.shape)
out["hog_bit_exact_vs_oracle"] = bool(np.array_equal(hog, whog))         # only the B200 path is bit-identical to the C restatement (the reference's AVX2 leaves contract to FMA)
out["hog_close"] = bool(hog.shape == whog.shape and np.allclose(hog, whog, rtol=0, atol=1e-4))
binar = ((img > 120) * 255).astype(np.uint8)
lsl = oracle.ccl_lsl("ref", binar, threads=1)
wlsl = oracle.ccl_lsl("orc", binar)
out["plsl_equal"] = bool(lsl["na"] == wlsl["na"] and np.array_equal(lsl["labels"], wlsl["labels"]) and np.array_equal(lsl["boxes"], wlsl["boxes"]))
blobs, _ = oracle.ccl_lsl_extract_ref(binar, blob=True, threads=1)
segs, segboxes = oracle.ccl_lsl_extract_ref(binar, blob=False, threads=1)
ok = len(blobs) == int(wlsl["na"])
for a, b in enumerate(blobs):
    ys, xs = np.nonzero(wlsl["labels"] == a + 1)
    ok = ok and np.array_equal(b, np.stack([xs, ys], 1).astype(np.int16))
out["plsl_extract_equal"] = bool(ok)
out["plsl_segment_boxes_ok"] = bool(all(tuple(segboxes[a]) == (s[0::2, 0].min(), s[0::2, 1].min(), s[1::2, 0].max(), s[1::2, 1].max()) for a, s in enumerate(segs)))
ms = oracle.ccl_lmser("ref", img, threads=1)
wms = oracle.ccl_lmser("orc", img)


def region_set(r):
    s = set()
    for n, b, p in zip(r["sizes"], r["boxes"], r["points"]):
        s.add((int(n), tuple(int(v) for v in b), hash(np.sort(p[:, 1].astype(np.int64) * 65536 + p[:, 0]).tobytes())))
    return s


out["mser_equal"] = bool(region_set(ms) == region_set(wms) and len(ms["sizes"]) == len(wms["sizes"]) > 0)
out["gpu_launches"] = int(b200.cvb200_launch_count()) - launches0
print(json.dumps(out))
