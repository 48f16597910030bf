import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
import oracle
from scipy import ndimage
r = oracle.ref(1)
def ref_lsl(img):
    h,w = img.shape
    lab = np.zeros((h,w),np.int32); na=C.c_int32(0); boxes=np.zeros((w*h,4),np.int16)
    rc = r.ref_ccl_lsl(img.ctypes.data_as(C.c_void_p), C.c_size_t(w),C.c_size_t(h),C.c_size_t(w), lab.ctypes.data_as(C.c_void_p), C.byref(na), boxes.ctypes.data_as(C.c_void_p), C.c_size_t(w*h), 0, None)
    assert rc==0, rc
    return lab, na.value, boxes[:na.value]
rng=np.random.default_rng(0)
bad=0
for trial in range(300):
    h,w = rng.integers(3,60), rng.integers(17,90)
    p = rng.choice([0.3,0.45,0.5,0.6,0.7])
    img = ((rng.random((h,w))<p)*255).astype(np.uint8)
    lab,na,_ = ref_lsl(img)
    gt,n = ndimage.label(img, structure=np.ones((3,3)))
    # partition equal?
    ok = n==na
    if ok:
        # map
        m = {}
        for a,b in zip(lab.ravel(), gt.ravel()):
            if m.setdefault(a,b)!=b: ok=False;break
    # numbering: by first pixel in raster order?
    if ok:
        firsts=[np.flatnonzero(lab.ravel()==k)[0] for k in range(1,na+1)]
        mono = all(firsts[i]<firsts[i+1] for i in range(len(firsts)-1))
    else: mono=None
    if not ok or not mono:
        bad+=1
        if bad<5: print("trial",trial,h,w,p,"na",na,"gt",n,"ok",ok,"mono",mono)
print("bad",bad)
