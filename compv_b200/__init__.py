"""compv_b200 -- B200-native (sm_100a) implementation of CompV's per-pixel image hot path.

The product is `lib/libcompv_b200.so` (hand-written CUDA kernels behind the C ABI of include/cvb200.h).  This Python
package is the harness-side binding used by tests/ and bench.py: numpy in, numpy out, every call goes through the
C ABI.  Class and method names mirror the reference's C++ API (CompVEdgeDete::newObj / process, CompVMathConvlt::convlt1 ...)
so that the parity tests read like the reference's own tests.  There is no CPU fallback: a missing library or a missing
GPU raises.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import (CvbError, lib, check, vp, sz,  # noqa: F401
                   SOBEL_ID, SCHARR_ID, PREWITT_ID, CANNY_ID, FAST_ID, HOUGHSHT_ID, HOUGHKHT_ID, HOGS_ID, PLSL_ID, LMSER_ID,
                   BORDER_TYPE_ZERO, BORDER_TYPE_IGNORE, BORDER_TYPE_REPLICATE)

_initialised_device = None

CONV_TYPES = {
    "8u16s16s": (np.uint8, np.int16, np.int16),
    "16s16s16s": (np.int16, np.int16, np.int16),
    "8u32f8u": (np.uint8, np.float32, np.uint8),
    "8u32f32f": (np.uint8, np.float32, np.float32),
    "32f32f32f": (np.float32, np.float32, np.float32),
    "32f32f8u": (np.float32, np.float32, np.uint8),
    "fxp_8u16u8u": (np.uint8, np.uint16, np.uint8),
}


def init(device=0):
    """cvb200_init: bind to a CUDA device (replaces CompVGpu::init, gpu/compv_gpu.cxx:36-62)."""
    global _initialised_device
    if _initialised_device != device:
        check(lib().cvb200_init(int(device)), "cvb200_init")
        _initialised_device = device


def launch_count():
    return int(lib().cvb200_launch_count())


def _frame(img, width):
    assert img.ndim == 2 and img.flags.c_contiguous, "frames are (height, stride) C-contiguous arrays"
    h, stride = img.shape
    return (stride if width is None else int(width)), h, stride


# ---- a2: CompVMathConvlt ---------------------------------------------------------------------------
def convlt1(name, img, vt_kern, hz_kern, width=None, border=BORDER_TYPE_ZERO, out=None):
    """CompVMathConvlt::convlt1<In,Kern,Out> / convlt1FixedPoint on host buffers.  Returns the (height, stride) output plane."""
    tin, tk, tout = CONV_TYPES[name]
    assert img.dtype == tin
    vt = np.ascontiguousarray(vt_kern, dtype=tk)
    hz = np.ascontiguousarray(hz_kern, dtype=tk)
    assert len(vt) == len(hz)
    w, h, stride = _frame(img, width)
    if out is None:
        out = np.zeros((h, stride), dtype=tout)
    fn = getattr(lib(), "cvb200_convlt1_" + name)
    check(fn(vp(img), sz(w), sz(h), sz(stride), vp(vt), vp(hz), sz(len(vt)), vp(out), int(border)), "cvb200_convlt1_" + name)
    return out


def gauss_kernel(size, sigma, fixed_point=False):
    """CompVMathGauss::kernelDim1<float> / kernelDim1FixedPoint."""
    k = np.zeros(size, dtype=np.uint16 if fixed_point else np.float32)
    fn = lib().cvb200_gauss_kernel_dim1_fxp if fixed_point else lib().cvb200_gauss_kernel_dim1_32f
    check(fn(sz(size), C.c_float(sigma), vp(k)), "cvb200_gauss_kernel_dim1")
    return k


def sobel_g(img, kind_id=SOBEL_ID, ks=3, width=None):
    """gx, gy (int16) and g = |gx|+|gy| (uint16): convlt1<u8,int16,int16> x2 + CompVMathUtils::sumAbs."""
    w, h, stride = _frame(img, width)
    gx = np.zeros((h, stride), np.int16)
    gy = np.zeros((h, stride), np.int16)
    g = np.zeros((h, stride), np.uint16)
    check(lib().cvb200_sobel_g(vp(img), sz(w), sz(h), sz(stride), int(kind_id), sz(ks), vp(gx), vp(gy), vp(g)), "cvb200_sobel_g")
    return gx, gy, g


# ---- a3 / a5: CompVEdgeDete ------------------------------------------------------------------------
class CompVEdgeDete:
    """Mirror of CompVEdgeDete (base/include/compv/base/compv_features.h:205-215) over cvb200_edge_dete_*."""

    EDGE_SET_BOOL_X86_SSE41_GMAX_LANES = 1000
    EDGE_SET_BOOL_GENERIC_KERNEL = 1001

    def __init__(self, handle, dete_id):
        self._h = handle
        self.id = dete_id

    @staticmethod
    def newObj(dete_id, tLow=0.8, tHigh=1.6, kernSize=3):
        h = C.c_void_p()
        check(lib().cvb200_edge_dete_new(C.byref(h), int(dete_id), C.c_float(tLow), C.c_float(tHigh), sz(kernSize)), "cvb200_edge_dete_new")
        return CompVEdgeDete(h, dete_id)

    def set(self, cap_id, value, ctype):
        v = ctype(value)
        return lib().cvb200_edge_dete_set(self._h, int(cap_id), C.byref(v), sz(C.sizeof(v)))

    def setInt(self, cap_id, value):
        check(self.set(cap_id, value, C.c_int32), "cvb200_edge_dete_set")

    def setFloat32(self, cap_id, value):
        check(self.set(cap_id, value, C.c_float), "cvb200_edge_dete_set")

    def setBool(self, cap_id, value):
        check(self.set(cap_id, bool(value), C.c_bool), "cvb200_edge_dete_set")

    def set_preblur(self, size, sigma):
        check(lib().cvb200_edge_dete_set_preblur(self._h, sz(size), C.c_float(sigma)), "cvb200_edge_dete_set_preblur")

    def process(self, image, width=None, edges=None):
        """Host buffers in, host buffers out (H2D + kernels + D2H inside the call)."""
        w, h, stride = _frame(image, width)
        if edges is None:
            edges = np.zeros((h, stride), np.uint8)
        check(lib().cvb200_edge_dete_process(self._h, vp(image), sz(w), sz(h), sz(stride), vp(edges)), "cvb200_edge_dete_process")
        return edges

    def process_batch(self, images, width=None, edges=None):
        """(batch, height, stride) host array (ideally pinned) -> same-shape edge maps; pipelined H2D/compute/D2H inside the call."""
        assert images.ndim == 3 and images.flags.c_contiguous
        b, h, stride = images.shape
        w = stride if width is None else int(width)
        if edges is None:
            edges = np.zeros_like(images)
        check(lib().cvb200_edge_dete_process_batch(self._h, vp(images), sz(w), sz(h), sz(stride), vp(edges), sz(b), sz(h * stride)), "cvb200_edge_dete_process_batch")
        return edges

    def process_dev(self, d_image, width, height, stride, d_edges, batch=1, frame_pitch=0, stream=0):
        """Device pointers (ints or torch tensors), batched, on `stream` (a cudaStream_t value)."""
        check(lib().cvb200_edge_dete_process_dev(self._h, vp(d_image), sz(width), sz(height), sz(stride), vp(d_edges), sz(batch), sz(frame_pitch),
                                                 C.c_void_p(stream)), "cvb200_edge_dete_process_dev")

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_edge_dete_free(C.byref(self._h))
        except Exception:
            pass


# ---- a8: CompVCornerDete (FAST) --------------------------------------------------------------------
POINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("strength", np.float32), ("orient", np.float32), ("level", np.int32), ("size", np.float32)])


def fast_scores(img, N=9, threshold=20, width=None):
    """K11 strength map through cvb200_fast_scores (shape of the reference's GPU hook processData)."""
    w, h, stride = _frame(img, width)
    out = np.zeros((h, stride), np.uint8)
    check(lib().cvb200_fast_scores(vp(img), sz(w), sz(h), sz(stride), int(N), int(threshold), vp(out)), "cvb200_fast_scores")
    return out


class CompVCornerDete:
    """Mirror of CompVCornerDete (base/include/compv/base/compv_features.h:166-174) over cvb200_corner_dete_*."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def newObj(dete_id=FAST_ID):
        h = C.c_void_p()
        check(lib().cvb200_corner_dete_new(C.byref(h), int(dete_id)), "cvb200_corner_dete_new")
        return CompVCornerDete(h)

    def set(self, cap_id, value, ctype):
        v = ctype(value)
        return lib().cvb200_corner_dete_set(self._h, int(cap_id), C.byref(v), sz(C.sizeof(v)))

    def setInt(self, cap_id, value):
        check(self.set(cap_id, value, C.c_int32), "cvb200_corner_dete_set")

    def setBool(self, cap_id, value):
        check(self.set(cap_id, bool(value), C.c_bool), "cvb200_corner_dete_set")

    def process(self, image, width=None, capacity=None):
        w, h, stride = _frame(image, width)
        cap = int(capacity if capacity is not None else w * h)
        pts = np.zeros(max(cap, 1), POINT_DTYPE)
        cnt = C.c_size_t(0)
        check(lib().cvb200_corner_dete_process(self._h, vp(image), sz(w), sz(h), sz(stride), vp(pts), sz(cap), C.byref(cnt)), "cvb200_corner_dete_process")
        return pts[:min(cnt.value, cap)].copy()

    def process_dev(self, d_image, width, height, stride, d_points, capacity, d_counts, batch=1, frame_pitch=0, stream=0):
        check(lib().cvb200_corner_dete_process_dev(self._h, vp(d_image), sz(width), sz(height), sz(stride), vp(d_points), sz(capacity), vp(d_counts), sz(batch),
                                                   sz(frame_pitch), C.c_void_p(stream)), "cvb200_corner_dete_process_dev")

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_corner_dete_free(C.byref(self._h))
        except Exception:
            pass


# ---- a6 / a7: CompVHough ---------------------------------------------------------------------------
LINE_DTYPE = np.dtype([("rho", np.float32), ("theta", np.float32), ("strength", np.uint64)])


class CompVHough:
    """Mirror of CompVHough (base/include/compv/base/compv_features.h:217-227) over cvb200_hough_*."""

    HOUGH_SET_BOOL_X86_SIMD_SCAN = 1002

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def newObj(hough_id=HOUGHKHT_ID, rho=1.0, theta=1.0, threshold=1):
        h = C.c_void_p()
        check(lib().cvb200_hough_new(C.byref(h), int(hough_id), C.c_float(rho), C.c_float(theta), sz(threshold)), "cvb200_hough_new")
        return CompVHough(h)

    def set(self, cap_id, value, ctype):
        v = ctype(value)
        return lib().cvb200_hough_set(self._h, int(cap_id), C.byref(v), sz(C.sizeof(v)))

    def setInt(self, cap_id, value):
        check(self.set(cap_id, value, C.c_int32), "cvb200_hough_set")

    def setFloat32(self, cap_id, value):
        check(self.set(cap_id, value, C.c_float), "cvb200_hough_set")

    def setBool(self, cap_id, value):
        check(self.set(cap_id, bool(value), C.c_bool), "cvb200_hough_set")

    def getFloat64(self, cap_id):
        v = C.c_double(0)
        check(lib().cvb200_hough_get(self._h, int(cap_id), C.byref(v), sz(8)), "cvb200_hough_get")
        return v.value

    def process(self, edges, width=None, capacity=1 << 16):
        w, h, stride = _frame(edges, width)
        lines = np.zeros(capacity, LINE_DTYPE)
        cnt = C.c_size_t(0)
        check(lib().cvb200_hough_process(self._h, vp(edges), sz(w), sz(h), sz(stride), vp(lines), sz(capacity), C.byref(cnt)), "cvb200_hough_process")
        return lines[:min(cnt.value, capacity)].copy()

    def process_dev(self, d_edges, width, height, stride, batch=1, frame_pitch=0, capacity=4096, stream=0):
        """Device edge maps in; returns a list of per-frame line arrays (host)."""
        lines = np.zeros((batch, capacity), LINE_DTYPE)
        counts = np.zeros(batch, np.uint64)
        check(lib().cvb200_hough_process_dev(self._h, vp(d_edges), sz(width), sz(height), sz(stride), sz(batch), sz(frame_pitch), vp(lines), sz(capacity), vp(counts),
                                             C.c_void_p(stream)), "cvb200_hough_process_dev")
        return [lines[f, :min(int(counts[f]), capacity)].copy() for f in range(batch)]

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_hough_free(C.byref(self._h))
        except Exception:
            pass


RANGE_DTYPE = np.dtype([("a", np.int32), ("start", np.int16), ("end", np.int16)])


class CompVConnectedComponentLabelingResult:
    """Mirror of CompVConnectedComponentLabelingResultLSL (base/include/compv/base/compv_ccl.h:138-156) over cvb200_ccl_result_*."""

    def __init__(self, handle, width, height):
        self._h = handle
        self.width, self.height = width, height

    def labelsCount(self):
        return int(lib().cvb200_ccl_result_labels_count(self._h))

    def labelIds(self):
        return np.arange(1, self.labelsCount() + 1, dtype=np.int32)

    def segments(self):
        """(row_offsets uint32 (height+1), ranges RANGE_DTYPE): the LEA in CSR form."""
        ro, rg, n = C.POINTER(C.c_uint32)(), C.c_void_p(), C.c_size_t(0)
        check(lib().cvb200_ccl_result_segments(self._h, C.byref(ro), C.byref(rg), C.byref(n)), "cvb200_ccl_result_segments")
        row_offsets = np.ctypeslib.as_array(ro, shape=(self.height + 1,)).copy()
        if not n.value:
            return row_offsets, np.zeros(0, RANGE_DTYPE)
        buf = (C.c_char * (n.value * RANGE_DTYPE.itemsize)).from_address(rg.value)
        return row_offsets, np.frombuffer(buf, RANGE_DTYPE).copy()

    def regions(self):
        """LMSER results: dict(sizes int32 (n,), boxes int16 (n, 4), points list of (size, 2) int16 (x, y)) -- CompVConnectedComponentLabelingResultLMSER::points()."""
        ps, pb, pp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nr, npts = C.c_size_t(0), C.c_size_t(0)
        check(lib().cvb200_ccl_result_regions(self._h, C.byref(ps), C.byref(pb), C.byref(pp), C.byref(nr), C.byref(npts)), "cvb200_ccl_result_regions")
        n, m = nr.value, npts.value
        if not n:
            return dict(sizes=np.zeros(0, np.int32), boxes=np.zeros((0, 4), np.int16), points=[])
        sizes = np.frombuffer((C.c_char * (4 * n)).from_address(ps.value), np.int32).copy()
        boxes = np.frombuffer((C.c_char * (8 * n)).from_address(pb.value), np.int16).reshape(n, 4).copy()
        pts = np.frombuffer((C.c_char * (4 * m)).from_address(pp.value), np.int16).reshape(m, 2).copy()
        offs = np.concatenate([[0], np.cumsum(sizes)])
        return dict(sizes=sizes, boxes=boxes, points=[pts[offs[i]:offs[i + 1]] for i in range(n)])

    def debugFlatten(self):
        labels = np.zeros((self.height, self.width), np.int32)
        check(lib().cvb200_ccl_result_flatten(self._h, vp(labels), sz(self.width)), "cvb200_ccl_result_flatten")
        return labels

    def boundingBoxes(self):
        n = C.c_size_t(0)
        boxes = np.zeros((max(1, self.labelsCount()), 4), np.int16)
        check(lib().cvb200_ccl_result_bounding_boxes(self._h, vp(boxes), sz(len(boxes)), C.byref(n)), "cvb200_ccl_result_bounding_boxes")
        return boxes[:n.value].copy()

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_ccl_result_free(C.byref(self._h))
        except Exception:
            pass


class CompVConnectedComponentLabeling:
    """Mirror of CompVConnectedComponentLabeling (base/include/compv/base/compv_ccl.h:173-241) over cvb200_ccl_*."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def newObj(ccl_id=_ffi.PLSL_ID, delta=5, min_area=0.0002, max_area=0.5, max_variation=0.5, min_diversity=0.5, connectivity=8):
        """CompVConnectedComponentLabeling::newObj (compv_ccl.h:229-236), same defaults."""
        h = C.c_void_p()
        check(lib().cvb200_ccl_new_ex(C.byref(h), int(ccl_id), int(delta), C.c_double(min_area), C.c_double(max_area), C.c_double(max_variation), C.c_double(min_diversity),
                                      int(connectivity)), "cvb200_ccl_new_ex")
        return CompVConnectedComponentLabeling(h)

    def set(self, cap_id, value, ctype):
        v = ctype(value)
        return lib().cvb200_ccl_set(self._h, int(cap_id), C.byref(v), sz(C.sizeof(v)))

    def process(self, binar, width=None):
        w, h, stride = _frame(binar, width)
        r = C.c_void_p()
        check(lib().cvb200_ccl_process(self._h, vp(binar), sz(w), sz(h), sz(stride), C.byref(r)), "cvb200_ccl_process")
        return CompVConnectedComponentLabelingResult(r, w, h)

    def process_dev(self, d_binar, width, height, stride, batch=1, frame_pitch=0, d_labels=None, want_results=False, stream=0):
        """Device frames in; returns (na int32 array, list of results or None).  d_labels (device int32, batch*height*width) receives the label images."""
        na = np.zeros(batch, np.int32)
        handles = (C.c_void_p * batch)()
        check(lib().cvb200_ccl_process_dev(self._h, vp(d_binar), sz(width), sz(height), sz(stride), sz(batch), sz(frame_pitch),
                                           vp(d_labels) if d_labels is not None else None, vp(na), handles if want_results else None, C.c_void_p(stream)),
              "cvb200_ccl_process_dev")
        results = [CompVConnectedComponentLabelingResult(C.c_void_p(handles[f]), width, height) for f in range(batch)] if want_results else None
        return na, results

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_ccl_free(C.byref(self._h))
        except Exception:
            pass


# ---- section 8f next row 1: CompVMathMorph -----------------------------------------------------------------
def morph_strel(size, strel_type):
    """CompVMathMorph::buildStructuringElement((width, height), type) -> (height, width) uint8."""
    sw, sh = size
    out = np.zeros((sh, sw), np.uint8)
    check(lib().cvb200_morph_build_strel(vp(out), sz(sw), sz(sh), sz(sw), int(strel_type)), "cvb200_morph_build_strel")
    return out


def morph(img, strel, op, border=BORDER_TYPE_REPLICATE, width=None, fill=0):
    """CompVMathMorph::process on a host frame; `fill` pre-fills the output (visible only with border IGNORE)."""
    w, h, stride = _frame(img, width)
    strel = np.ascontiguousarray(strel, np.uint8)
    sh, sw = strel.shape
    out = np.full((h, stride), fill, np.uint8)
    check(lib().cvb200_morph_process(vp(img), sz(w), sz(h), sz(stride), vp(strel), sz(sw), sz(sh), sz(sw), vp(out), int(op), int(border)), "cvb200_morph_process")
    return out


def morph_dev(d_in, width, height, stride, strel, op, d_out, border=BORDER_TYPE_REPLICATE, batch=1, frame_pitch=0, stream=0):
    strel = np.ascontiguousarray(strel, np.uint8)
    sh, sw = strel.shape
    check(lib().cvb200_morph_process_dev(vp(d_in), sz(width), sz(height), sz(stride), vp(strel), sz(sw), sz(sh), sz(sw), vp(d_out), int(op), int(border), sz(batch), sz(frame_pitch),
                                         C.c_void_p(stream)), "cvb200_morph_process_dev")


def canny_kht_process_batch(canny, hough, images, width=None, capacity=4096):
    """Host frames (batch, height, stride) -> list of per-frame line arrays; cvb200_canny_kht_process_batch."""
    assert images.ndim == 3 and images.flags.c_contiguous
    b, h, stride = images.shape
    w = stride if width is None else int(width)
    lines = np.zeros((b, capacity), LINE_DTYPE)
    counts = np.zeros(b, np.uint64)
    check(lib().cvb200_canny_kht_process_batch(canny._h, hough._h, vp(images), sz(w), sz(h), sz(stride), sz(b), sz(h * stride), vp(lines), sz(capacity), vp(counts)),
          "cvb200_canny_kht_process_batch")
    return [lines[f, :min(int(counts[f]), capacity)].copy() for f in range(b)]


def image_to_grayscale(pixel_format, data, width, height, stride):
    """cvb200_image_to_grayscale: host frame bytes in `pixel_format` (COMPV_SUBTYPE_PIXELS_* value) -> (height, stride) gray plane."""
    data = np.ascontiguousarray(data, np.uint8)
    out = np.zeros((height, stride), np.uint8)
    check(lib().cvb200_image_to_grayscale(int(pixel_format), vp(data), sz(width), sz(height), sz(stride), vp(out)), "cvb200_image_to_grayscale")
    return out


def canny_kht_process_batch_fmt(canny, hough, pixel_format, frames, width, height, stride, frame_pitch_bytes, capacity=4096):
    """Host frames in a camera format (batch, frame_pitch_bytes) uint8 -> list of per-frame line arrays; cvb200_canny_kht_process_batch_fmt."""
    assert frames.ndim == 2 and frames.flags.c_contiguous and frames.shape[1] == frame_pitch_bytes
    b = frames.shape[0]
    lines = np.zeros((b, capacity), LINE_DTYPE)
    counts = np.zeros(b, np.uint64)
    check(lib().cvb200_canny_kht_process_batch_fmt(canny._h, hough._h, int(pixel_format), vp(frames), sz(width), sz(height), sz(stride), sz(b), sz(frame_pitch_bytes), vp(lines), sz(capacity),
                                                   vp(counts)), "cvb200_canny_kht_process_batch_fmt")
    return [lines[f, :min(int(counts[f]), capacity)].copy() for f in range(b)]


def init_devices(count=0):
    """cvb200_init_devices: every device of the process (count <= 0) or the first `count`; returns how many are active."""
    check(lib().cvb200_init_devices(int(count)), "cvb200_init_devices")
    return int(lib().cvb200_active_device_count())


def canny_kht_process_batch_multi(canny, hough, images, width=None, capacity=4096):
    """Host frames spread over every initialised device in-process; cvb200_canny_kht_process_batch_multi."""
    assert images.ndim == 3 and images.flags.c_contiguous
    b, h, stride = images.shape
    w = stride if width is None else int(width)
    lines = np.zeros((b, capacity), LINE_DTYPE)
    counts = np.zeros(b, np.uint64)
    check(lib().cvb200_canny_kht_process_batch_multi(canny._h, hough._h, vp(images), sz(w), sz(h), sz(stride), sz(b), sz(h * stride), vp(lines), sz(capacity), vp(counts)),
          "cvb200_canny_kht_process_batch_multi")
    return [lines[f, :min(int(counts[f]), capacity)].copy() for f in range(b)]


def canny_kht_process_batch_dev(canny, hough, d_images, width, height, stride, batch, capacity=4096, frame_pitch=0, stream=0):
    """Device frames -> list of per-frame line arrays (host); cvb200_canny_kht_process_batch_dev."""
    lines = np.zeros((batch, capacity), LINE_DTYPE)
    counts = np.zeros(batch, np.uint64)
    check(lib().cvb200_canny_kht_process_batch_dev(canny._h, hough._h, vp(d_images), sz(width), sz(height), sz(stride), sz(batch), sz(frame_pitch), vp(lines), sz(capacity), vp(counts),
                                                   C.c_void_p(stream)), "cvb200_canny_kht_process_batch_dev")
    return [lines[f, :min(int(counts[f]), capacity)].copy() for f in range(batch)]


# ---- a10: thresholding ------------------------------------------------------------------------------
def histogram(img, width=None):
    """CompVMathHistogram::build (8-bit): 256 uint32 bins."""
    w, h, stride = _frame(img, width)
    hist = np.zeros(256, np.uint32)
    check(lib().cvb200_histogram_8u(vp(img), sz(w), sz(h), sz(stride), vp(hist)), "cvb200_histogram_8u")
    return hist


def threshold_global(img, threshold, width=None):
    w, h, stride = _frame(img, width)
    out = np.zeros((h, stride), np.uint8)
    check(lib().cvb200_threshold_global(vp(img), sz(w), sz(h), sz(stride), C.c_double(threshold), vp(out)), "cvb200_threshold_global")
    return out


def threshold_otsu(img, width=None, want_output=True):
    """Returns (out or None, threshold)."""
    w, h, stride = _frame(img, width)
    out = np.zeros((h, stride), np.uint8) if want_output else None
    thr = C.c_double(0)
    check(lib().cvb200_threshold_otsu(vp(img), sz(w), sz(h), sz(stride), C.byref(thr), vp(out)), "cvb200_threshold_otsu")
    return out, thr.value


def threshold_adaptive(img, block_size=5, delta=8.0, max_val=255.0, invert=False, width=None):
    w, h, stride = _frame(img, width)
    out = np.zeros((h, stride), np.uint8)
    check(lib().cvb200_threshold_adaptive(vp(img), sz(w), sz(h), sz(stride), sz(block_size), C.c_double(delta), C.c_double(max_val), int(bool(invert)), vp(out)),
          "cvb200_threshold_adaptive")
    return out


# ---- a4: CompVGradientFast, a9: CompVHOG ---------------------------------------------------------------
def gradient_fast(img, width=None):
    w, h, stride = _frame(img, width)
    o = {"gx16": np.zeros((h, stride), np.int16), "gy16": np.zeros((h, stride), np.int16), "gx32": np.zeros((h, stride), np.float32),
         "gy32": np.zeros((h, stride), np.float32), "mag": np.zeros((h, stride), np.float32), "dir": np.zeros((h, stride), np.float32)}
    check(lib().cvb200_gradient_fast_8u(vp(img), sz(w), sz(h), sz(stride), vp(o["gx16"]), vp(o["gy16"]), vp(o["gx32"]), vp(o["gy32"]), vp(o["mag"]), vp(o["dir"])), "cvb200_gradient_fast_8u")
    return o


class CompVHOG:
    """Mirror of CompVHOG (base/include/compv/base/compv_features.h:229-261) over cvb200_hog_*."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def newObj(hog_id=HOGS_ID, blockSize=(16, 16), blockStride=(8, 8), cellSize=(8, 8), nbins=9, blockNorm=_ffi.HOG_BLOCK_NORM_L2HYS, gradientSigned=True,
               interp=_ffi.HOG_INTERPOLATION_BILINEAR):
        h = C.c_void_p()
        check(lib().cvb200_hog_new(C.byref(h), int(hog_id), sz(blockSize[0]), sz(blockSize[1]), sz(blockStride[0]), sz(blockStride[1]), sz(cellSize[0]), sz(cellSize[1]),
                                   sz(nbins), int(blockNorm), int(bool(gradientSigned)), int(interp)), "cvb200_hog_new")
        return CompVHOG(h)

    def set(self, cap_id, value, ctype):
        v = ctype(value)
        return lib().cvb200_hog_set(self._h, int(cap_id), C.byref(v), sz(C.sizeof(v)))

    def descriptorSize(self, width, height):
        n = C.c_size_t(0)
        check(lib().cvb200_hog_descriptor_size(self._h, sz(width), sz(height), C.byref(n)), "cvb200_hog_descriptor_size")
        return n.value

    def process(self, img, width=None):
        w, h, stride = _frame(img, width)
        n = self.descriptorSize(w, h)
        out = np.zeros(n, np.float32)
        size = C.c_size_t(0)
        fn = lib().cvb200_hog_process if img.dtype == np.uint8 else lib().cvb200_hog_process_32f
        check(fn(self._h, vp(img), sz(w), sz(h), sz(stride), vp(out), sz(n), C.byref(size)), "cvb200_hog_process")
        return out[:size.value]

    def process_dev(self, d_in, width, height, stride, d_out, batch=1, frame_pitch=0, stream=0):
        check(lib().cvb200_hog_process_dev(self._h, vp(d_in), sz(width), sz(height), sz(stride), vp(d_out), sz(batch), sz(frame_pitch), C.c_void_p(stream)), "cvb200_hog_process_dev")

    def __del__(self):
        try:
            if self._h:
                lib().cvb200_hog_free(C.byref(self._h))
        except Exception:
            pass
