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
out["gpu_launches"] = int(b200.cvb200_launch_count()) - launches0
print(json.dumps(out))
