"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the two CPU checkers:

  * `orc`  : oracle/libcompv_oracle.so, the plain-C restatement (oracle/compv_oracle*.c), built by `make -C oracle`;
  * `ref`  : oracle/_ref/libcompv_refshim.so, the UNMODIFIED reference compiled from /root/reference by oracle/build_ref.sh
             (present here and, as a prebuilt .so, on the GPU box; never rebuilt there).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.  The product (compv_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libcompv_oracle.so")
REFSHIM_SO = os.path.join(_HERE, "_ref", "libcompv_refshim.so")

_orc = None
_ref = None
_ref_threads = None


def build(verbose=False):
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE], stdout=out)
    if os.path.isdir("/root/reference/base"):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")], stdout=out)


def orc():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _orc = C.CDLL(ORACLE_SO)
    return _orc


def have_ref():
    return os.path.exists(REFSHIM_SO)


def ref(threads=1):
    """The compiled reference.  `threads`: 1 = single threaded (CompVBase::init(1)), -1 = one per core."""
    global _ref, _ref_threads
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libcompv_refshim.so is missing: run oracle/build_ref.sh where /root/reference exists")
        _ref = C.CDLL(REFSHIM_SO)
        _ref.ref_cpu_flags.restype = C.c_char_p
        rc = _ref.ref_init(int(threads))
        if rc:
            raise RuntimeError("ref_init failed: %d" % rc)
        _ref_threads = threads
    elif threads != _ref_threads:
        rc = _ref.ref_set_max_threads(int(threads))
        if rc:
            raise RuntimeError("ref_set_max_threads failed: %d" % rc)
        _ref_threads = threads
    return _ref


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _sz(v):
    return C.c_size_t(int(v))


def _chk(rc, what):
    if rc:
        raise RuntimeError("%s failed: %d" % (what, rc))


# numpy dtypes of the convolution variants: name -> (in, kernel, out)
CONV_TYPES = {
    "8u16s16s": (np.uint8, np.int16, np.int16),
    "16s16s16s": (np.int16, np.int16, np.int16),
    "8u32f8u": (np.uint8, np.float32, np.uint8),
    "8u32f32f": (np.uint8, np.float32, np.float32),
    "32f32f32f": (np.float32, np.float32, np.float32),
    "32f32f8u": (np.float32, np.float32, np.uint8),
    "fxp_8u16u8u": (np.uint8, np.uint16, np.uint8),
}


def _frame_args(img, width=None):
    """img is a 2-D array whose row pitch is the stride; `width` (<= pitch) is the logical width."""
    assert img.ndim == 2 and img.flags.c_contiguous
    h, stride = img.shape
    w = stride if width is None else width
    return w, h, stride


def convlt1(which, name, img, vt, hz, width=None, border=0, out=None):
    """which: 'orc' or 'ref'.  Returns the (h, stride) output plane."""
    tin, tk, tout = CONV_TYPES[name]
    assert img.dtype == tin
    vt = np.ascontiguousarray(vt, dtype=tk)
    hz = np.ascontiguousarray(hz, dtype=tk)
    w, h, stride = _frame_args(img, width)
    if out is None:
        out = np.zeros((h, stride), dtype=tout)
    if which == "orc":
        fn = getattr(orc(), "orc_convlt1_" + name)
        _chk(fn(_p(img), _sz(w), _sz(h), _sz(stride), _p(vt), _p(hz), _sz(len(vt)), _p(out), int(border)), "orc_convlt1_" + name)
    else:
        assert border == 0
        fn = getattr(ref(), "ref_convlt1_" + name)
        _chk(fn(_p(img), _sz(w), _sz(h), _sz(stride), _p(vt), _p(hz), _sz(len(vt)), _p(out)), "ref_convlt1_" + name)
    return out


def gauss_kernel(which, size, sigma, fixed_point=False):
    k = np.zeros(size, dtype=np.uint16 if fixed_point else np.float32)
    lib, pre = (orc(), "orc") if which == "orc" else (ref(), "ref")
    fn = getattr(lib, pre + ("_gauss_kernel_dim1_fxp" if fixed_point else "_gauss_kernel_dim1_32f"))
    _chk(fn(_sz(size), C.c_float(sigma), _p(k)), "gauss_kernel")
    return k


EDGE_WHICH = {"sobel": 0, "scharr": 1, "prewitt": 2, "canny": 3}


def sobel_g(which, img, kind="sobel", ks=3, width=None):
    w, h, stride = _frame_args(img, width)
    gx = np.zeros((h, stride), np.int16)
    gy = np.zeros((h, stride), np.int16)
    g = np.zeros((h, stride), np.uint16)
    if which == "orc":
        _chk(orc().orc_sobel_g(_p(img), _sz(w), _sz(h), _sz(stride), EDGE_WHICH[kind], int(ks), _p(gx), _p(gy), _p(g)), "orc_sobel_g")
    else:
        tabs = {("sobel", 3): ([1, 2, 1], [-1, 0, 1]), ("sobel", 5): ([1, 4, 6, 4, 1], [1, 2, 0, -2, -1]),
                ("canny", 3): ([1, 2, 1], [-1, 0, 1]), ("canny", 5): ([1, 4, 6, 4, 1], [1, 2, 0, -2, -1]),
                ("scharr", 3): ([3, 10, 3], [-1, 0, 1]), ("prewitt", 3): ([1, 1, 1], [-1, 0, 1])}
        vt, hz = tabs[(kind, ks)]
        gx = convlt1("ref", "8u16s16s", img, vt, hz, width=width)
        gy = convlt1("ref", "8u16s16s", img, hz, vt, width=width)
        _chk(ref().ref_sum_abs_16s16u(_p(gx), _p(gy), _p(g), _sz(w), _sz(h), _sz(stride)), "ref_sum_abs")
    return gx, gy, g


def edge_dete(which, img, kind="canny", tlow=59.0, thigh=119.0, ks=3, width=None, threshold_type=0, threads=1, simd=True, sse41_gmax_lanes=False):
    """Sobel/Scharr/Prewitt normalised gradient or Canny edge map, (h, stride) uint8.
    simd=False runs the reference's plain C++ path (CompVCpu::flagsDisable(kCpuFlagAll)); sse41_gmax_lanes: see orc_edge_normalized."""
    w, h, stride = _frame_args(img, width)
    out = np.zeros((h, stride), np.uint8)
    if which == "orc":
        if kind == "canny":
            _chk(orc().orc_canny(_p(img), _sz(w), _sz(h), _sz(stride), C.c_float(tlow), C.c_float(thigh), int(ks), int(threshold_type), _p(out)), "orc_canny")
        else:
            _chk(orc().orc_edge_normalized(_p(img), _sz(w), _sz(h), _sz(stride), EDGE_WHICH[kind], int(bool(sse41_gmax_lanes)), _p(out)), "orc_edge_normalized")
    else:
        r = ref(threads)
        if not simd:
            _chk(r.ref_cpu_simd(0), "ref_cpu_simd")
        try:
            _chk(r.ref_edge_dete(EDGE_WHICH[kind], _p(img), _sz(w), _sz(h), _sz(stride), C.c_float(tlow), C.c_float(thigh), int(ks),
                                 int(threshold_type), _p(out)), "ref_edge_dete")
        finally:
            if not simd:
                _chk(r.ref_cpu_simd(1), "ref_cpu_simd")
    return out


def time_edge_dete(img, kind="canny", tlow=59.0, thigh=119.0, ks=3, blur_size=0, blur_sigma=1.0, iters=10, threads=-1, width=None):
    """Times the reference (ms per iteration, numpy array) -- CompVEdgeDete::process with an optional Gaussian pre-blur."""
    w, h, stride = _frame_args(img, width)
    ms = np.zeros(iters, np.float64)
    out = np.zeros((h, stride), np.uint8)
    _chk(ref(threads).ref_time_edge_dete(EDGE_WHICH[kind], _p(img), _sz(w), _sz(h), _sz(stride), C.c_float(tlow), C.c_float(thigh), int(ks),
                                         int(blur_size), C.c_float(blur_sigma), int(iters), _p(ms), _p(out)), "ref_time_edge_dete")
    return ms, out


class RefEdgeSession:
    """Reference detector + pre-wrapped frames created once; run() times blur (optional) + CompVEdgeDete::process over frames."""

    def __init__(self, frames, kind="canny", tlow=59.0, thigh=119.0, ks=3, blur_size=0, blur_sigma=1.0, threads=-1, width=None, kht_threshold=0):
        assert frames.ndim == 3 and frames.flags.c_contiguous and frames.dtype == np.uint8
        n, h, stride = frames.shape
        w = stride if width is None else width
        self._r = ref(threads)
        self._r.ref_edge_session_new.restype = C.c_void_p
        self._r.ref_edge_session_run.restype = C.c_double
        self.shape = (h, stride)
        self.count = n
        self._s = C.c_void_p(self._r.ref_edge_session_new(EDGE_WHICH[kind], _p(frames), _sz(n), _sz(w), _sz(h), _sz(stride), C.c_float(tlow), C.c_float(thigh),
                                                          int(ks), int(blur_size), C.c_float(blur_sigma)))
        if not self._s:
            raise RuntimeError("ref_edge_session_new failed")
        if kht_threshold:
            _chk(self._r.ref_edge_session_add_kht(self._s, _sz(kht_threshold)), "ref_edge_session_add_kht")

    def run(self, first=0, count=None, want_edges=False):
        """Returns (elapsed ms, last edge map or None)."""
        count = self.count if count is None else count
        out = np.zeros(self.shape, np.uint8) if want_edges else None
        ms = self._r.ref_edge_session_run(self._s, _sz(first), _sz(count), _p(out), _sz(self.shape[1]))
        if ms < 0:
            raise RuntimeError("ref_edge_session_run failed: %r" % ms)
        return ms, out

    def close(self):
        if self._s:
            self._r.ref_edge_session_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


POINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("strength", np.float32), ("orient", np.float32), ("level", np.int32), ("size", np.float32)])


def fast_scores(img, N=9, threshold=20, width=None):
    """Oracle strength map (K11), (h, stride) uint8."""
    w, h, stride = _frame_args(img, width)
    out = np.zeros((h, stride), np.uint8)
    _chk(orc().orc_fast_scores(_p(img), _sz(w), _sz(h), _sz(stride), int(N), int(threshold), _p(out)), "orc_fast_scores")
    return out


def fast_detect(which, img, N=9, threshold=20, nms=True, max_features=-1, width=None, threads=1, iters=0):
    """Interest points (structured array, POINT_DTYPE) in the order the implementation produced them.  With which='ref' and iters>0 also returns ms/iter."""
    w, h, stride = _frame_args(img, width)
    cap = w * h
    pts = np.zeros(cap, POINT_DTYPE)
    cnt = C.c_size_t(0)
    if which == "orc":
        assert max_features <= 1
        _chk(orc().orc_fast_detect(_p(img), _sz(w), _sz(h), _sz(stride), int(N), int(threshold), int(bool(nms)), _p(pts), _sz(cap), C.byref(cnt)), "orc_fast_detect")
        return pts[:cnt.value].copy()
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_fast_detect(_p(img), _sz(w), _sz(h), _sz(stride), int(N), int(threshold), int(bool(nms)), int(max_features), _p(pts), _sz(cap), C.byref(cnt),
                                      int(iters), _p(ms)), "ref_fast_detect")
    out = pts[:cnt.value].copy()
    return (out, ms[:iters]) if iters else out


def orb_detect(which, img, threshold=20, nms=True, max_features=2000, n=9, width=None, threads=1, iters=0):
    """The ORB DETECTOR (pyramid of 8 levels at 0.83, FAST per level, per-level quota, orientation): points in the reference's order (level by level)."""
    w, h, stride = _frame_args(img, width)
    cap = 1 << 16
    pts = np.zeros(cap, POINT_DTYPE)
    cnt = C.c_size_t(0)
    if which == "orc":
        _chk(orc().orc_orb_detect(_p(img), _sz(w), _sz(h), _sz(stride), int(threshold), int(bool(nms)), int(n), int(max_features), _p(pts), _sz(cap), C.byref(cnt)), "orc_orb_detect")
        return pts[:cnt.value].copy()
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_orb_detect(_p(img), _sz(w), _sz(h), _sz(stride), int(threshold), int(bool(nms)), int(max_features), _p(pts), _sz(cap), C.byref(cnt), int(iters), _p(ms)), "ref_orb_detect")
    out = pts[:cnt.value].copy()
    return (out, ms[:iters]) if iters else out


def fast_detect_after_ref(img_a, img_b):
    """Reference defect probe (see ref_shim.cxx ref_fast_detect_after): FAST on img_b with a detector that has just processed img_a.  Returns (points, shared_stride)."""
    r = ref(1)
    ha, wa = img_a.shape
    hb, wb = img_b.shape
    pts = np.zeros(wb * hb, POINT_DTYPE)
    cnt = C.c_size_t(0)
    rc = r.ref_fast_detect_after(_p(np.ascontiguousarray(img_a)), _sz(wa), _sz(ha), _p(np.ascontiguousarray(img_b)), _sz(wb), _sz(hb), _p(pts), _sz(len(pts)), C.byref(cnt))
    if rc not in (0, -7):
        _chk(rc, "ref_fast_detect_after")
    return pts[:cnt.value].copy(), rc == 0


def scale_bilinear(which, img, out_w, out_h, width=None):
    """CompVImage::scale(..., COMPV_INTERPOLATION_TYPE_BILINEAR): 8-bit fixed-point bilinear kernel."""
    w, h, stride = _frame_args(img, width)
    out = np.zeros((out_h, out_w), np.uint8)
    if which == "orc":
        _chk(orc().orc_scale_bilinear(_p(img), _sz(w), _sz(h), _sz(stride), _p(out), _sz(out_w), _sz(out_h), _sz(out_w)), "orc_scale_bilinear")
    else:
        _chk(ref(1).ref_scale_bilinear(_p(img), _sz(w), _sz(h), _sz(stride), _p(out), _sz(out_w), _sz(out_h)), "ref_scale_bilinear")
    return out


LINE_DTYPE = np.dtype([("rho", np.float32), ("theta", np.float32), ("strength", np.uint64)])


def hough_kht(which, edges, rho=1.0, theta=1.0, threshold=1, max_lines=0, cluster_min_deviation=2.0, cluster_min_size=10, kernel_min_height=0.002,
              width=None, threads=1, x86_simd=True, iters=0):
    """KHT lines (structured array LINE_DTYPE, in detection order) and Gs.  which: 'orc' | 'ref'."""
    w, h, stride = _frame_args(edges, width)
    cap = 1 << 16
    lines = np.zeros(cap, LINE_DTYPE)
    cnt = C.c_size_t(0)
    gs = C.c_double(0)
    if which == "orc":
        _chk(orc().orc_hough_kht(_p(edges), _sz(w), _sz(h), _sz(stride), C.c_float(rho), C.c_float(theta), _sz(threshold), C.c_float(cluster_min_deviation),
                                 int(cluster_min_size), C.c_float(kernel_min_height), int(max_lines), int(bool(x86_simd)), _p(lines), _sz(cap), C.byref(cnt), C.byref(gs)),
             "orc_hough_kht")
        return lines[:min(cnt.value, cap)].copy(), gs.value
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_hough(1, _p(edges), _sz(w), _sz(h), _sz(stride), C.c_float(rho), C.c_float(theta), _sz(threshold), int(max_lines),
                                C.c_float(cluster_min_deviation), int(cluster_min_size), C.c_float(kernel_min_height), _p(lines), _sz(cap), C.byref(cnt), C.byref(gs),
                                int(iters), _p(ms)), "ref_hough")
    out = lines[:min(cnt.value, cap)].copy()
    return (out, gs.value, ms[:iters]) if iters else (out, gs.value)


def hough_sht(which, edges, rho=1.0, theta=1.0, threshold=1, max_lines=0, width=None, threads=1, x86_simd=True, iters=0, cap=1 << 16):
    """SHT lines (LINE_DTYPE, sorted by strength as the reference returns them) and the untruncated count.  which: 'orc' | 'ref'.
    x86_simd=False selects the generic C++ path (for 'ref': SIMD disabled around the call)."""
    w, h, stride = _frame_args(edges, width)
    lines = np.zeros(cap, LINE_DTYPE)
    cnt = C.c_size_t(0)
    if which == "orc":
        _chk(orc().orc_hough_sht(_p(edges), _sz(w), _sz(h), _sz(stride), C.c_float(rho), C.c_float(theta), _sz(threshold), int(max_lines), int(bool(x86_simd)),
                                 _p(lines), _sz(cap), C.byref(cnt)), "orc_hough_sht")
        return lines[:min(cnt.value, cap)].copy(), cnt.value
    gs = C.c_double(0)
    ms = np.zeros(max(iters, 1), np.float64)
    r = ref(threads)
    if not x86_simd:
        _chk(r.ref_cpu_simd(0), "ref_cpu_simd")
    try:
        _chk(r.ref_hough(0, _p(edges), _sz(w), _sz(h), _sz(stride), C.c_float(rho), C.c_float(theta), _sz(threshold), int(max_lines),
                         C.c_float(0), 0, C.c_float(0), _p(lines), _sz(cap), C.byref(cnt), C.byref(gs), int(iters), _p(ms)), "ref_hough")
    finally:
        if not x86_simd:
            _chk(r.ref_cpu_simd(1), "ref_cpu_simd")
    out = lines[:min(cnt.value, cap)].copy()
    return (out, cnt.value, ms[:iters]) if iters else (out, cnt.value)


RANGE_DTYPE = np.dtype([("a", np.int32), ("start", np.int16), ("end", np.int16)])


def ccl_lsl(which, img, width=None, threads=1, iters=0):
    """PLSL labelling.  Returns dict(labels=int32 (h, w), na=int, boxes=int16 (na, 4) {left, top, right, bottom}[, row_offsets, ranges for 'orc'][, ms])."""
    w, h, stride = _frame_args(img, width)
    labels = np.zeros((h, w), np.int32)
    na = C.c_int32(0)
    boxes = np.zeros((max(1, (w * h + 1) // 2), 4), np.int16)
    if which == "orc":
        row_off = np.zeros(h + 1, np.uint32)
        ranges = np.zeros(max(1, (w + 1) // 2 * h), RANGE_DTYPE)
        nr = C.c_size_t(0)
        _chk(orc().orc_ccl_lsl(_p(img), _sz(w), _sz(h), _sz(stride), _p(labels), C.byref(na), _p(boxes), _sz(len(boxes)), _p(row_off), _p(ranges), _sz(len(ranges)),
                               C.byref(nr)), "orc_ccl_lsl")
        return dict(labels=labels, na=na.value, boxes=boxes[:na.value].copy(), row_offsets=row_off, ranges=ranges[:nr.value].copy())
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_ccl_lsl(_p(img), _sz(w), _sz(h), _sz(stride), _p(labels), C.byref(na), _p(boxes), _sz(len(boxes)), int(iters), _p(ms)), "ref_ccl_lsl")
    out = dict(labels=labels, na=na.value, boxes=boxes[:na.value].copy())
    if iters:
        out["ms"] = ms[:iters]
    return out


def to_grayscale_ref(subtype, data, width, height, stride, threads=1):
    """The REFERENCE's CompVImage::wrap + convertGrayscale on a whole frame buffer (bytes) in `subtype` layout; returns the (height, width) gray plane."""
    r = ref(threads)
    data = np.ascontiguousarray(data, np.uint8)
    out = np.zeros((height, width), np.uint8)
    _chk(r.ref_to_grayscale(int(subtype), _p(data), _sz(width), _sz(height), _sz(stride), _p(out), _sz(width)), "ref_to_grayscale")
    return out


def ccl_lsl_extract_ref(img, blob=True, width=None, threads=1):
    """The REFERENCE's CompVConnectedComponentLabelingResultLSL::extract (ccl_lsl_result.cxx:100-134): list of (n, 2) int16 arrays (x, y), one per label;
    for segments also the boxes the reference derives from them (ccl_lsl_result.cxx:187-230)."""
    r = ref(threads)
    w, h, stride = _frame_args(img, width)
    nl, npnt = C.c_size_t(0), C.c_size_t(0)
    _chk(r.ref_ccl_lsl_extract(_p(img), _sz(w), _sz(h), _sz(stride), int(bool(blob)), None, _sz(0), None, _sz(0), C.byref(nl), C.byref(npnt), None), "ref_ccl_lsl_extract")
    counts = np.zeros(max(nl.value, 1), np.int32)
    pts = np.zeros((max(npnt.value, 1), 2), np.int16)
    boxes = np.zeros((max(nl.value, 1), 4), np.int16)
    _chk(r.ref_ccl_lsl_extract(_p(img), _sz(w), _sz(h), _sz(stride), int(bool(blob)), _p(counts), _sz(len(counts)), _p(pts), _sz(len(pts)), C.byref(nl), C.byref(npnt),
                               None if blob else _p(boxes)), "ref_ccl_lsl_extract")
    out, o = [], 0
    for a in range(nl.value):
        out.append(pts[o:o + counts[a]].copy())
        o += counts[a]
    return out, (None if blob else boxes[:nl.value])


def hough_to_cartesian_ref(kht, width, height, rho, theta):
    """The REFERENCE's CompVHough::toCartesian (houghkht.cxx:1249-1280 / houghsht.cxx:566-592): (n, 4) float32 {a.x, a.y, b.x, b.y}."""
    r = ref(1)
    rho = np.ascontiguousarray(rho, np.float32)
    theta = np.ascontiguousarray(theta, np.float32)
    out = np.zeros((len(rho), 4), np.float32)
    _chk(r.ref_hough_to_cartesian(int(bool(kht)), _sz(width), _sz(height), _sz(len(rho)), _p(rho), _p(theta), _p(out)), "ref_hough_to_cartesian")
    return out


def ccl_lmser(which, img, delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8, width=None, threads=1, iters=0,
              point_cap=None):
    """Linear-time MSER.  Defaults are the reference's unit-test parameters (unittests/ccl_mser.cxx:26-46).
    Returns dict(sizes=int32 (n,), boxes=int16 (n, 4) {left, top, right, bottom}, points=list of (size, 2) int16 (x, y) arrays in the reference's order[, ms])."""
    w, h, stride = _frame_args(img, width)
    region_cap = w * h + 2
    if point_cap is None:
        point_cap = 64 * w * h
    sizes = np.zeros(region_cap, np.int32)
    boxes = np.zeros((region_cap, 4), np.int16)
    pts = np.zeros((point_cap, 2), np.int16)
    nr, npts = C.c_size_t(0), C.c_size_t(0)
    args = [_p(img), _sz(w), _sz(h), _sz(stride), int(delta), C.c_double(min_area), C.c_double(max_area), C.c_double(max_variation), C.c_double(min_diversity), int(connectivity),
            _p(sizes), _p(boxes), _sz(region_cap), _p(pts), _sz(point_cap), C.byref(nr), C.byref(npts)]
    ms = np.zeros(max(iters, 1), np.float64)
    if which == "orc":
        _chk(orc().orc_ccl_lmser(*args), "orc_ccl_lmser")
    else:
        _chk(ref(threads).ref_ccl_lmser(*(args + [int(iters), _p(ms)])), "ref_ccl_lmser")
    n = nr.value
    assert npts.value <= point_cap, "point_cap too small: %d points" % npts.value
    offs = np.concatenate([[0], np.cumsum(sizes[:n])])
    out = dict(sizes=sizes[:n].copy(), boxes=boxes[:n].copy(), points=[pts[offs[i]:offs[i + 1]].copy() for i in range(n)])
    if iters and which != "orc":
        out["ms"] = ms[:iters]
    return out


STREL_RECT, STREL_DIAMOND, STREL_CROSS = 0, 1, 2
MORPH_ERODE, MORPH_DILATE, MORPH_OPEN, MORPH_CLOSE = 0, 1, 2, 3


def morph_strel(which, size, strel_type):
    """CompVMathMorph::buildStructuringElement: (height, width) uint8 array."""
    sw, sh = size
    out = np.zeros((sh, sw), np.uint8)
    if which == "orc":
        _chk(orc().orc_morph_strel(_p(out), _sz(sw), _sz(sh), int(strel_type)), "orc_morph_strel")
    else:
        dummy = np.zeros((max(sh, 1), max(sw, 1)), np.uint8)
        _chk(ref(1).ref_morph(_p(dummy), _sz(sw), _sz(sh), _sz(sw), int(strel_type), None, _sz(sw), _sz(sh), _p(out), 0, 0, None, 0, None), "ref_morph")
    return out


def morph(which, img, strel, op, border=2, width=None, fill=0, threads=1, iters=0):
    """CompVMathMorph::process with an explicit structuring element.  `fill`: value the output is pre-filled with (visible only with border IGNORE)."""
    w, h, stride = _frame_args(img, width)
    strel = np.ascontiguousarray(strel, np.uint8)
    sh, sw = strel.shape
    out = np.full((h, stride), fill, np.uint8)
    if which == "orc":
        _chk(orc().orc_morph_process(_p(img), _sz(w), _sz(h), _sz(stride), _p(strel), _sz(sw), _sz(sh), _sz(sw), _p(out), int(op), int(border)), "orc_morph_process")
        return out
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_morph(_p(img), _sz(w), _sz(h), _sz(stride), -1, _p(strel), _sz(sw), _sz(sh), None, int(op), int(border), _p(out), int(iters), _p(ms)), "ref_morph")
    return (out, ms[:iters]) if iters else out


def histogram(img, width=None):
    w, h, stride = _frame_args(img, width)
    hist = np.zeros(256, np.uint32)
    _chk(orc().orc_histogram_8u(_p(img), _sz(w), _sz(h), _sz(stride), _p(hist)), "orc_histogram_8u")
    return hist


def threshold(which, mode, img, threshold=128.0, block_size=5, delta=8.0, max_val=255.0, invert=False, width=None, threads=1):
    """mode: 'global' | 'otsu' | 'adaptive'.  Returns (out, otsu_threshold_or_None)."""
    w, h, stride = _frame_args(img, width)
    out = np.zeros((h, stride), np.uint8)
    thr = C.c_double(0)
    if which == "orc":
        if mode == "global":
            _chk(orc().orc_threshold_global(_p(img), _sz(w), _sz(h), _sz(stride), C.c_double(threshold), _p(out)), "orc_threshold_global")
            return out, None
        if mode == "otsu":
            _chk(orc().orc_threshold_otsu(_p(img), _sz(w), _sz(h), _sz(stride), C.byref(thr), _p(out)), "orc_threshold_otsu")
            return out, thr.value
        _chk(orc().orc_threshold_adaptive(_p(img), _sz(w), _sz(h), _sz(stride), _sz(block_size), C.c_double(delta), C.c_double(max_val), int(bool(invert)), _p(out)), "orc_threshold_adaptive")
        return out, None
    m = {"global": 0, "otsu": 1, "adaptive": 2}[mode]
    _chk(ref(threads).ref_threshold(m, _p(img), _sz(w), _sz(h), _sz(stride), C.c_double(threshold), _sz(block_size), C.c_double(delta), C.c_double(max_val), int(bool(invert)),
                                    C.byref(thr), _p(out), 0, None), "ref_threshold")
    return out, (thr.value if mode == "otsu" else None)


def gradient_fast(which, img, width=None):
    """Returns dict gx16, gy16, gx32, gy32, mag, dir (degrees)."""
    w, h, stride = _frame_args(img, width)
    o = {"gx16": np.zeros((h, stride), np.int16), "gy16": np.zeros((h, stride), np.int16), "gx32": np.zeros((h, stride), np.float32),
         "gy32": np.zeros((h, stride), np.float32), "mag": np.zeros((h, stride), np.float32), "dir": np.zeros((h, stride), np.float32)}
    fn = orc().orc_gradient_fast_8u if which == "orc" else ref().ref_gradient_fast
    _chk(fn(_p(img), _sz(w), _sz(h), _sz(stride), _p(o["gx16"]), _p(o["gy16"]), _p(o["gx32"]), _p(o["gy32"]), _p(o["mag"]), _p(o["dir"])), "gradient_fast")
    return o


def hog(which, img, block=(16, 16), stride_=(8, 8), cell=(8, 8), nbins=9, block_norm=52, gradient_signed=True, interp=55, width=None, threads=1, iters=0):
    """S-HOG descriptor (float32 vector).  block_norm 48..52 (none, L1, L1sqrt, L2, L2Hys), interp 53..55 (nearest, bilinear LUT, bilinear)."""
    w, h, stride = _frame_args(img, width)
    size = C.c_size_t(0)
    cap = 1 << 24
    out = np.zeros(cap, np.float32)
    if which == "orc":
        _chk(orc().orc_hog(_p(img), _sz(w), _sz(h), _sz(stride), _sz(block[0]), _sz(block[1]), _sz(stride_[0]), _sz(stride_[1]), _sz(cell[0]), _sz(cell[1]), _sz(nbins),
                           int(block_norm), int(bool(gradient_signed)), int(interp), _p(out), _sz(cap), C.byref(size)), "orc_hog")
        return out[:size.value].copy()
    ms = np.zeros(max(iters, 1), np.float64)
    _chk(ref(threads).ref_hog(_p(img), _sz(w), _sz(h), _sz(stride), _sz(block[0]), _sz(block[1]), _sz(stride_[0]), _sz(stride_[1]), _sz(cell[0]), _sz(cell[1]), _sz(nbins),
                              int(block_norm), int(bool(gradient_signed)), int(interp), _p(out), _sz(cap), C.byref(size), int(iters), _p(ms)), "ref_hog")
    o = out[:size.value].copy()
    return (o, ms[:iters]) if iters else o
