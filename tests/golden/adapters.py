"""Adapters that let tests/golden/cases.py run on the compiled reference / the oracle (`OracleOrRef`) or on the CUDA library (`Cuda`)."""
import numpy as np

import oracle


def _mser_canonical(res):
    """Order-free digest input: regions sorted by (size, box); per region size, box and the sorted pixel indices."""
    items = []
    for size, box, pts in zip(res["sizes"], res["boxes"], res["points"]):
        key = np.sort(pts[:, 1].astype(np.int64) * 65536 + pts[:, 0].astype(np.int64))
        items.append((int(size), tuple(int(v) for v in box), key))
    items.sort(key=lambda it: (it[0], it[1], it[2].tobytes()))
    out = [np.array([len(items)], np.int64)]
    for size, box, key in items:
        out += [np.array([size], np.int64), np.array(box, np.int64), key]
    return out


class OracleOrRef:
    def __init__(self, which):
        self.w = which
        self.kw = dict(threads=1) if which == "ref" else {}

    def gauss_kernel(self, n, sigma):
        return oracle.gauss_kernel(self.w, n, sigma)

    def convlt(self, name, img, vt, hz):
        return oracle.convlt1(self.w, name, img, vt, hz)

    def edge(self, img, kind, tlow, thigh):
        # Sobel: the x86 SSE4.1 max quirk is part of what the reference returns on this machine (defect 1)
        if self.w == "ref":
            return oracle.edge_dete("ref", img, kind, tlow, thigh, 3, threads=1)
        return oracle.edge_dete("orc", img, kind, tlow, thigh, 3, sse41_gmax_lanes=True) if kind == "sobel" else oracle.edge_dete("orc", img, kind, tlow, thigh, 3)

    def sht(self, edges, thr):
        return oracle.hough_sht(self.w, edges, 1.0, 1.0, thr, **self.kw)[0]

    def kht(self, edges, thr):
        return oracle.hough_kht(self.w, edges, 1.0, 1.0, thr, **self.kw)[0]

    def fast(self, img, n, t):
        return oracle.fast_detect(self.w, img, n, t, True, **self.kw)

    def otsu(self, img):
        return oracle.threshold(self.w, "otsu", img, **self.kw)

    def adaptive(self, img):
        return oracle.threshold(self.w, "adaptive", img, **self.kw)[0]

    def lsl(self, img):
        return oracle.ccl_lsl(self.w, img, **self.kw)

    def mser_canonical(self, img):
        return _mser_canonical(oracle.ccl_lmser(self.w, img, **self.kw))

    def strel(self, size, t):
        return oracle.morph_strel(self.w, size, t)

    def morph(self, img, se, op):
        return oracle.morph(self.w, img, se, op, **self.kw)


class Cuda:
    def __init__(self, cvb):
        self.cvb = cvb
        from compv_b200 import _ffi
        self.ffi = _ffi

    def gauss_kernel(self, n, sigma):
        return self.cvb.gauss_kernel(n, sigma)

    def convlt(self, name, img, vt, hz):
        return self.cvb.convlt1(name, img, vt, hz)

    def edge(self, img, kind, tlow, thigh):
        ids = {"sobel": self.ffi.SOBEL_ID, "canny": self.ffi.CANNY_ID}
        d = self.cvb.CompVEdgeDete.newObj(ids[kind], tlow, thigh, 3)
        if kind == "sobel":
            d.setBool(self.ffi.EDGE_SET_BOOL_X86_SSE41_GMAX_LANES, True)
        return d.process(img)

    def sht(self, edges, thr):
        return self.cvb.CompVHough.newObj(self.ffi.HOUGHSHT_ID, 1.0, 1.0, thr).process(edges, capacity=1 << 20)

    def kht(self, edges, thr):
        return self.cvb.CompVHough.newObj(self.ffi.HOUGHKHT_ID, 1.0, 1.0, thr).process(edges)

    def fast(self, img, n, t):
        d = self.cvb.CompVCornerDete.newObj(self.ffi.FAST_ID)
        d.setInt(self.ffi.FAST_SET_INT_THRESHOLD, t)
        d.setInt(self.ffi.FAST_SET_INT_FAST_TYPE, self.ffi.FAST_TYPE_9 if n == 9 else self.ffi.FAST_TYPE_12)
        d.setInt(self.ffi.FAST_SET_INT_MAX_FEATURES, -1)
        d.setBool(self.ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, True)
        return d.process(img)

    def otsu(self, img):
        out, thr = self.cvb.threshold_otsu(img)
        return out, thr

    def adaptive(self, img):
        return self.cvb.threshold_adaptive(img)

    def lsl(self, img):
        r = self.cvb.CompVConnectedComponentLabeling.newObj(self.ffi.PLSL_ID).process(img)
        return dict(labels=r.debugFlatten(), boxes=r.boundingBoxes(), na=r.labelsCount())

    def mser_canonical(self, img):
        # cases.py runs MSER with the oracle wrapper's defaults (unittests/ccl_mser.cxx parameters)
        r = self.cvb.CompVConnectedComponentLabeling.newObj(self.ffi.LMSER_ID, delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2,
                                                            connectivity=8).process(img)
        return _mser_canonical(r.regions())

    def strel(self, size, t):
        return self.cvb.morph_strel(size, t)

    def morph(self, img, se, op):
        return self.cvb.morph(img, se, op)
