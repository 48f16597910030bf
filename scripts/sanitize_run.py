"""One small call of every kernel family, for `compute-sanitizer --tool memcheck python scripts/sanitize_run.py` (no torch, small frames, odd sizes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compv_b200 as cvb  # noqa: E402
from compv_b200 import _ffi  # noqa: E402
from frames import frame_g, frame_text, frame_uniform  # noqa: E402

cvb.init(0)
for (w, h) in [(64, 48), (333, 211), (640, 360)]:
    img = frame_g(w, h, 3)
    k = cvb.gauss_kernel(5, 1.0)
    cvb.convlt1("8u32f8u", img, k, k)
    cvb.convlt1("8u16s16s", img, np.array([1, 2, 1], np.int16), np.array([-1, 0, 1], np.int16))
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    edges = canny.process(img)
    cvb.CompVEdgeDete.newObj(_ffi.SOBEL_ID).process(img)
    print(w, h, "kht", len(cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 20).process(edges)),
          "sht", len(cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, 1.0, 40).process(edges)))
    fast = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
    print("fast", len(fast.process(img)))
    cvb.CompVHOG.newObj().process(img) if w >= 16 and h >= 16 else None
    cvb.threshold_otsu(img)
    cvb.threshold_adaptive(img)
    cvb.gradient_fast(img)
    binar = ((frame_text(w, h, 2) < 128) * 255).astype(np.uint8)
    noise = ((frame_uniform(w, h, 5) > 120) * 255).astype(np.uint8)
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    print("plsl", ccl.process(binar).labelsCount(), ccl.process(noise).labelsCount())
    mser = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, delta=2, min_area=0.0005, max_area=0.3, max_variation=0.5, min_diversity=0.2)
    print("mser", mser.process(img).labelsCount(), mser.process(frame_uniform(w, h, 9)).labelsCount())
    for t in (0, 1, 2):
        se = cvb.morph_strel((5, 5), t)
        for op in range(4):
            cvb.morph(binar, se, op)
frames = np.stack([frame_g(320, 200, 50 + i) for i in range(9)])
print("batch", [len(x) for x in cvb.canny_kht_process_batch(canny, cvb.CompVHough.newObj(_ffi.HOUGHKHT_ID, 1.0, 1.0, 30), frames, width=320)])
print("sanitize_run done")
