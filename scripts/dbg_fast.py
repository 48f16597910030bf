import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import compv_b200 as cvb
from frames import frame_g
cvb.init(0)
for (w, h) in [(1920, 1080), (64, 48)]:
    img = frame_g(w, h, 1)
    d = cvb.CompVEdgeDete.newObj(20, 59.0, 119.0, 3)
    try:
        out = d.process(img)
        print(w, h, "ok", int((out == 255).sum()))
    except Exception as e:
        print(w, h, "FAIL", e)
        break
