"""ctypes binding of libcompv_b200.so (the C ABI declared in include/cvb200.h).

This module is plumbing for tests/ and bench.py: it adds no algorithm.  It fails loudly when the CUDA library is
missing -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcompv_b200.so")

S_OK = 0
E_NOT_IMPLEMENTED = 20001
E_NOT_INITIALIZED = 20002
E_INVALID_CALL = 20004
E_INVALID_STATE = 20005
E_INVALID_PARAMETER = 20006
E_OUT_OF_MEMORY = 20013
E_OUT_OF_BOUND = 20014
E_CUDA = 20035

# ids (values of the reference's enums, see include/cvb200.h)
FAST_ID = 1
FAST_SET_INT_THRESHOLD = 2
FAST_SET_INT_MAX_FEATURES = 3
FAST_SET_INT_FAST_TYPE = 4
FAST_SET_BOOL_NON_MAXIMA_SUPP = 5
FAST_TYPE_9 = 6
FAST_TYPE_12 = 7
ORB_ID = 8
ORB_SET_INT_INTERNAL_DETE_ID = 9
ORB_SET_INT_FAST_THRESHOLD = 10
ORB_SET_BOOL_FAST_NON_MAXIMA_SUPP = 11
ORB_SET_INT_PYRAMID_LEVELS = 12
ORB_SET_INT_MAX_FEATURES = 15
EDGE_SET_BOOL_X86_SSE41_GMAX_LANES = 1000
EDGE_SET_BOOL_GENERIC_KERNEL = 1001
HOUGH_SET_BOOL_X86_SIMD_SCAN = 1002
CANNY_ID = 20
CANNY_SET_INT_KERNEL_SIZE = 21
CANNY_SET_INT_THRESHOLD_TYPE = 22
CANNY_SET_FLT32_THRESHOLD_LOW = 23
CANNY_SET_FLT32_THRESHOLD_HIGH = 24
CANNY_THRESHOLD_TYPE_PERCENT_OF_MEAN = 25
CANNY_THRESHOLD_TYPE_COMPARE_TO_GRADIENT = 26
SOBEL_ID = 27
SCHARR_ID = 28
PREWITT_ID = 29
HOUGHSHT_ID = 30
HOUGHKHT_ID = 31
HOUGH_SET_FLT32_RHO = 32
HOUGH_SET_FLT32_THETA = 33
HOUGH_SET_INT_THRESHOLD = 34
HOUGH_SET_INT_MAXLINES = 35
HOUGHKHT_SET_FLT32_CLUSTER_MIN_DEVIATION = 36
HOUGHKHT_SET_INT_CLUSTER_MIN_SIZE = 37
HOUGHKHT_SET_FLT32_KERNEL_MIN_HEIGTH = 38
HOUGHKHT_SET_BOOL_OVERRIDE_INPUT_EDGES = 39
HOUGHKHT_GET_FLT64_GS = 40
HOGS_ID = 41
HOG_BLOCK_NORM_NONE = 48
HOG_BLOCK_NORM_L1 = 49
HOG_BLOCK_NORM_L1SQRT = 50
HOG_BLOCK_NORM_L2 = 51
HOG_BLOCK_NORM_L2HYS = 52
HOG_INTERPOLATION_NEAREST = 53
HOG_INTERPOLATION_BILINEAR_LUT = 54
HOG_INTERPOLATION_BILINEAR = 55
PLSL_ID = 1
CCL_SET_INT_CONNECTIVITY = 0
PLSL_SET_INT_TYPE = 2
PLSL_SET_BOOL_SORT_SEGMENTS = 3
PLSL_TYPE_XRLEZ = 10
PLSL_TYPE_STD = 4
LMSER_ID = 19
BORDER_TYPE_ZERO = 0
BORDER_TYPE_IGNORE = 1
BORDER_TYPE_REPLICATE = 2


class CvbError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = "%s failed: %d" % (what, code)
        try:
            msg += " (%s) %s" % (lib().cvb200_error_string(code).decode(), lib().cvb200_last_cuda_error().decode())
        except Exception:
            pass
        super().__init__(msg)


_lib = None


def lib():
    """Load libcompv_b200.so (once).  Raises if the library has not been built: the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libcompv_b200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.cvb200_error_string.restype = C.c_char_p
        _lib.cvb200_last_cuda_error.restype = C.c_char_p
        _lib.cvb200_launch_count.restype = C.c_uint64
        _lib.cvb200_ccl_result_labels_count.restype = C.c_size_t
    return _lib


def check(code, what=""):
    if code != S_OK:
        raise CvbError(code, what)


def vp(x):
    """void* of a numpy array / int address / None."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "ctypes"):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def sz(x):
    return C.c_size_t(int(x))
