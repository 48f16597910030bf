#!/usr/bin/env python
"""One traced call of the Canny+KHT pipeline (CVB200_TRACE=1 prints the per-slot timeline). usage: python scripts/pipe_trace.py frames [host]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, time
import compv_b200 as cvb
from frames import frame_g
W, H = 1920, 1080
B = int(sys.argv[1]); host = len(sys.argv) > 2
cvb.init(0)
frames = np.stack([frame_g(W, H, 12345 + k) for k in range(16)])
frames = np.concatenate([frames] * ((B + 15) // 16))[:B]
h_in = torch.from_numpy(frames).pin_memory()
d_in = h_in.cuda()
dete = cvb.CompVEdgeDete.newObj(cvb.CANNY_ID, 59.0, 119.0, 3); dete.set_preblur(5, 1.0)
kht = cvb.CompVHough.newObj(cvb.HOUGHKHT_ID, 1.0, 1.0, 100)
for it in range(3):
    os.environ["CVB200_QUIET"] = "1"
    t0 = time.perf_counter()
    if host: cvb.canny_kht_process_batch(dete, kht, h_in.numpy(), width=W, capacity=512)
    else: cvb.canny_kht_process_batch_dev(dete, kht, d_in, W, H, W, B, capacity=512)
    print("call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
