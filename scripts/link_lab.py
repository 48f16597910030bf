#!/usr/bin/env python
"""Times the KHT kernels on device-resident Canny edge maps (1080p frame G) and checks the lines against the oracle.
usage: CVB200_KHT_LINK=<variant> python scripts/link_lab.py [frames] [check]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import compv_b200 as cvb
from frames import frame_g

W, H = 1920, 1080
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
check = len(sys.argv) > 2
cvb.init(0)
frames = np.stack([frame_g(W, H, 12345 + k) for k in range(min(B, 16))])
frames = np.concatenate([frames] * ((B + len(frames) - 1) // len(frames)))[:B]
d_in = torch.from_numpy(frames).cuda()
d_out = torch.empty_like(d_in)
dete = cvb.CompVEdgeDete.newObj(cvb.CANNY_ID, 59.0, 119.0, 3)
dete.set_preblur(5, 1.0)
kht = cvb.CompVHough.newObj(cvb.HOUGHKHT_ID, 1.0, 1.0, 100)
stream = torch.cuda.current_stream().cuda_stream
dete.process_dev(d_in, W, H, W, d_out, batch=B, stream=stream)
for _ in range(2):
    got = kht.process_dev(d_out, W, H, W, batch=B, stream=stream)
cvb.lib().cvb200_profile_begin()
R = 5
for _ in range(R):
    got = kht.process_dev(d_out, W, H, W, batch=B, stream=stream)
buf = C.create_string_buffer(1 << 16)
cvb.check(cvb.lib().cvb200_profile_end(buf, C.c_size_t(len(buf))), "profile_end")
out = {}
for ln in buf.value.decode().splitlines():
    name, cnt, ms = ln.split()
    out[name] = float(ms) / int(cnt)
print("variant", os.environ.get("CVB200_KHT_LINK", "default"), "frames", B, " ".join("%s=%.3f" % (k, v) for k, v in sorted(out.items())), "lines[0]=%d" % len(got[0]), flush=True)
if check:
    import oracle
    edges = d_out[:2].cpu().numpy()
    for k in range(2):
        want, _ = oracle.hough_kht("orc", edges[k], 1.0, 1.0, 100)
        assert len(want) == len(got[k]) and all(np.array_equal(got[k][f], want[f]) for f in ("rho", "theta", "strength")), "MISMATCH frame %d" % k
    print("lines match the oracle", flush=True)
