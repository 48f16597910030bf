"""One small pass of the serial-stage kernels for an ncu capture: Canny + KHT on a few 1080p frames, PLSL on a few text frames.
  ncu --set full --clock-control none --import-source on -k regex:'kht_link|lsl_equiv' -c 2 -o gpurun_out/x python scripts/prof_once.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compv_b200 as cvb  # noqa: E402
from compv_b200 import _ffi  # noqa: E402
from frames import frame_g, frame_text  # noqa: E402

W, H, B = 1920, 1080, int(os.environ.get("PROF_FRAMES", "4"))
cvb.init(0)
stream = torch.cuda.current_stream().cuda_stream
frames = np.stack([frame_g(W, H, 12345 + k) for k in range(B)])
d_in = torch.from_numpy(frames).cuda()
d_edges = torch.empty_like(d_in)
canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
canny.set_preblur(5, 1.0)
canny.process_dev(d_in, W, H, W, d_edges, batch=B, stream=stream)
kht = cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 100)
print("kht lines", [len(x) for x in kht.process_dev(d_edges, W, H, W, batch=B, stream=stream)])
if os.environ.get("PROF_NO_PLSL"):
    sys.exit(0)
text = np.stack([((frame_text(W, H, 20 + k) < 128) * 255).astype(np.uint8) for k in range(B)])
ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
print("plsl labels", ccl.process_dev(torch.from_numpy(text).cuda(), W, H, W, batch=B, stream=stream)[0])
