"""Row-strip mode (BASELINE config 4): ONE frame cut into horizontal strips, one strip per GPU / rank.

The reference splits the same stages across its threads by rows -- convolution with overlap rows (base/include/compv/base/math/compv_math_convlt.h:129-159), Canny
NMS + hysteresis per row band (core/features/edges/compv_core_feature_canny_dete.cxx:175-234, 282-306), SHT voting into per-thread accumulators that are summed
(core/features/hough/compv_core_feature_houghsht.cxx:455-477).  Here a rank is a thread of that scheme with its own GPU:

  stage                       what crosses the seams                                       collective
  Canny front (blur/Sobel/NMS) 4 halo rows of the INPUT image (2 blur + 1 Sobel + 1 NMS)     none (every rank uploads its strip + halo)
  Canny hysteresis             the class state of the first / last owned row                neighbour send/recv per round + all-reduce(SUM) of the strong count
  Sobel / Scharr / Prewitt     1 halo row; the frame maximum of |gx|+|gy|                    all-reduce(MAX)
  Otsu                         the 256-bin histogram                                         all-reduce(SUM)
  FAST                         4 halo rows (radius 3 + NMS)                                  all-gather of the points (rank order = raster order)
  SHT                          the accumulator                                               all-reduce(SUM, int32)

KHT linking and the PLSL equivalence pass are order-dependent over the whole frame: they stay frame-sharded.  LMSER: replicas only.

The functions take an `ops` object that runs the per-strip stages (CudaStripOps = the C-ABI stage entry points of libcompv_b200.so on this rank's GPU; the CPU tests
pass a numpy stand-in so that the partitioning / exchange / convergence logic runs over gloo without a GPU) and use torch.distributed only for what the table lists."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

CANNY_HALO = 4   # rows of input a strip needs beyond its own: Gaussian 5x5 (2) + Sobel 3x3 (1) + NMS (1)
FAST_HALO = 4    # Bresenham circle radius 3 + 3x3 non-maximum suppression of the strengths


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _staged(t):
    """NCCL moves device tensors itself; under gloo (CPU tests, or two ranks sharing one GPU) device tensors travel through host memory."""
    return t.is_cuda and dist.get_backend() != "nccl"


def _all_reduce(t, op):
    if _staged(t):
        h = t.cpu()
        dist.all_reduce(h, op=op)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=op)


def strip_range(height, rank, world, align=1):
    """Rows [y0, y1) owned by `rank`: contiguous, multiples of `align` (except the last strip), sizes differ by at most `align`, every row owned once."""
    assert 0 <= rank < world and height > 0 and align >= 1
    units = (height + align - 1) // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align, height), min(u1 * align, height)


def with_halo(y0, y1, height, halo):
    """The rows a rank loads: its own plus `halo` above and below, clipped to the frame (at the frame's border the strip's border IS the image border)."""
    return max(0, y0 - halo), min(height, y1 + halo)


def _exchange_boundary_rows(buf, rank, world):
    """buf = [halo row above | owned rows | halo row below].  My first owned row goes to rank-1's bottom halo, my last owned row to rank+1's top halo."""
    if world == 1:
        return
    staged = _staged(buf)
    ops_, recv = [], []
    if rank > 0:
        up = buf[1].cpu() if staged else buf[1].contiguous()
        got = torch.empty_like(up)
        ops_ += [dist.P2POp(dist.isend, up, rank - 1), dist.P2POp(dist.irecv, got, rank - 1)]
        recv.append((0, got))
    if rank < world - 1:
        down = buf[-2].cpu() if staged else buf[-2].contiguous()
        got = torch.empty_like(down)
        ops_ += [dist.P2POp(dist.isend, down, rank + 1), dist.P2POp(dist.irecv, got, rank + 1)]
        recv.append((buf.shape[0] - 1, got))
    for r in dist.batch_isend_irecv(ops_):
        r.wait()
    for row, got in recv:
        buf[row].copy_(got)


def canny_row_strips(ops, frame, tlow, thigh, blur=(5, 1.0)):
    """frame: the full (H, W) uint8 frame on the host (every rank holds it, or at least its strip + halo: only those rows are touched).
    Returns (y0, y1, edges) with edges = the (y1-y0, W) uint8 edge map {0, 255} of the rows this rank owns -- identical to the single-GPU result for those rows."""
    rank, world = _world()
    H, W = frame.shape
    y0, y1 = strip_range(H, rank, world)
    a0, a1 = with_halo(y0, y1, H, CANNY_HALO)
    cls = ops.canny_front(frame[a0:a1], tlow, thigh, blur)          # class map of the loaded rows: 0 / 0x80 weak / 0xff strong
    own = cls[y0 - a0:y1 - a0]
    buf = torch.zeros((own.shape[0] + 2, W), dtype=torch.uint8, device=own.device)
    buf[1:-1] = own
    prev = -1
    rounds = 0
    while True:
        _exchange_boundary_rows(buf, rank, world)
        ops.canny_closure(buf)                                      # 8-connected closure of the strong pixels inside buf, halo rows included
        total = (buf[1:-1] == 255).sum().to(torch.int64).reshape(1)
        if world > 1:
            _all_reduce(total, dist.ReduceOp.SUM)
        rounds += 1
        if int(total.item()) == prev:                               # a whole round promoted nothing anywhere: the closure is global
            break
        prev = int(total.item())
    ops.canny_finalize(buf)                                         # weak -> 0
    return y0, y1, buf[1:-1], rounds


def sobel_row_strips(ops, frame, kind="sobel"):
    """Sobel / Scharr / Prewitt detector (edge_dete.cxx:55-206): the normalisation needs the maximum of |gx|+|gy| over the WHOLE frame -> all-reduce(MAX)."""
    rank, world = _world()
    H, W = frame.shape
    y0, y1 = strip_range(H, rank, world)
    a0, a1 = with_halo(y0, y1, H, 1)
    gmax = ops.edge_gmax(frame[a0:a1], kind)                        # device uint32 [1]; the halo rows' own border rows contribute 0 (zero border of the convolution)
    g64 = gmax.to(torch.int64)
    if world > 1:
        _all_reduce(g64, dist.ReduceOp.MAX)
    out = ops.edge_normalize(frame[a0:a1], kind, g64.to(torch.int32))
    return y0, y1, out[y0 - a0:y1 - a0]


def otsu_row_strips(ops, frame):
    """Otsu (compv_image_threshold.cxx:52-116): histogram of the strip, all-reduce(SUM), the scan on every rank (256 steps), binarise the strip."""
    rank, world = _world()
    H, W = frame.shape
    y0, y1 = strip_range(H, rank, world)
    hist = ops.histogram(frame[y0:y1]).to(torch.int64)
    if world > 1:
        _all_reduce(hist, dist.ReduceOp.SUM)
    thr = ops.otsu_from_histogram(hist.cpu().numpy().astype(np.uint32), H * W)
    return y0, y1, thr, ops.threshold_global(frame[y0:y1], thr)


def fast_row_strips(ops, frame, threshold=20, n=9):
    """FAST + NMS (fast_dete.cxx:163-422): corners of the owned rows from a strip with 4 halo rows; all ranks' points gathered in rank order = the reference's raster order."""
    rank, world = _world()
    H, W = frame.shape
    y0, y1 = strip_range(H, rank, world)
    a0, a1 = with_halo(y0, y1, H, FAST_HALO)
    pts = ops.fast_points(frame[a0:a1], threshold, n)               # structured array x, y, strength (y relative to a0)
    pts = pts[(pts["y"] + a0 >= y0) & (pts["y"] + a0 < y1)].copy()
    pts["y"] += a0
    if world == 1:
        return pts
    gathered = [None] * world
    dist.all_gather_object(gathered, pts)
    return np.concatenate(gathered)


def sht_row_strips(ops, edges_strip, y0, full_height, threshold):
    """edges_strip: the (rows, W) edge map of the rows this rank owns, on the device (canny_row_strips's output).  Every rank votes its strip into an accumulator of the
    FULL frame's geometry, the accumulators are summed (all-reduce, int32), every rank extracts the same lines (houghsht.cxx:455-477 is the same sum over threads)."""
    rank, world = _world()
    acc = ops.sht_accumulate(edges_strip, y0, full_height, threshold)
    if world > 1:
        _all_reduce(acc, dist.ReduceOp.SUM)
    return ops.sht_lines(acc, edges_strip.shape[1], full_height, threshold)


class CudaStripOps:
    """The per-strip stages on this rank's GPU through the C ABI (cvb200_edge_dete_process_stages_dev, cvb200_hough_sht_accumulate_dev / _lines_dev, ...)."""

    def __init__(self, device):
        import compv_b200 as cvb
        self.cvb = cvb
        self.dev = torch.device("cuda", device)
        self._canny = {}
        self._edge = {}
        self._sht = {}
        self.stream = 0

    def _up(self, rows):
        return torch.from_numpy(np.ascontiguousarray(rows)).to(self.dev)

    def _canny_obj(self, tlow, thigh, blur):
        key = (tlow, thigh, blur)
        if key not in self._canny:
            d = self.cvb.CompVEdgeDete.newObj(self.cvb.CANNY_ID, tlow, thigh, 3)
            if blur:
                d.set_preblur(blur[0], blur[1])
            self._canny[key] = d
        self._last_canny = self._canny[key]
        return self._canny[key]

    def _stages(self, d, image, out, stages, gmax=None):
        h, w = out.shape
        cvb = self.cvb
        cvb.check(cvb.lib().cvb200_edge_dete_process_stages_dev(d._h, cvb.vp(image), cvb.sz(w), cvb.sz(h), cvb.sz(w), cvb.vp(out), cvb.sz(1), cvb.sz(0), int(stages), cvb.vp(gmax),
                                                                C.c_void_p(self.stream)), "cvb200_edge_dete_process_stages_dev")

    def canny_front(self, rows, tlow, thigh, blur):
        img = self._up(rows)
        cls = torch.empty_like(img)
        self._stages(self._canny_obj(tlow, thigh, blur), img, cls, 1)
        return cls

    def canny_closure(self, buf):
        self._stages(self._last_canny, None, buf, 2)

    def canny_finalize(self, buf):
        self._stages(self._last_canny, None, buf, 4)

    def _edge_obj(self, kind):
        ids = {"sobel": self.cvb.SOBEL_ID, "scharr": self.cvb.SCHARR_ID, "prewitt": self.cvb.PREWITT_ID}
        if kind not in self._edge:
            self._edge[kind] = self.cvb.CompVEdgeDete.newObj(ids[kind])
        return self._edge[kind]

    def edge_gmax(self, rows, kind):
        img = self._up(rows)
        gmax = torch.zeros(1, dtype=torch.int32, device=self.dev)
        scratch = torch.empty_like(img)
        self._stages(self._edge_obj(kind), img, scratch, 1, gmax)
        return gmax

    def edge_normalize(self, rows, kind, gmax):
        img = self._up(rows)
        out = torch.empty_like(img)
        self._stages(self._edge_obj(kind), img, out, 2, gmax.contiguous())
        return out

    def histogram(self, rows):
        cvb = self.cvb
        img = self._up(rows)
        h, w = img.shape
        hist = torch.zeros(256, dtype=torch.int32, device=self.dev)
        cvb.check(cvb.lib().cvb200_histogram_8u_dev(cvb.vp(img), cvb.sz(w), cvb.sz(h), cvb.sz(w), cvb.vp(hist), cvb.sz(1), cvb.sz(0), C.c_void_p(self.stream)), "cvb200_histogram_8u_dev")
        torch.cuda.synchronize(self.dev)
        return hist

    def otsu_from_histogram(self, hist_u32, count):
        thr = C.c_double(0)
        self.cvb.check(self.cvb.lib().cvb200_otsu_threshold_from_histogram(self.cvb.vp(np.ascontiguousarray(hist_u32, np.uint32)), self.cvb.sz(count), C.byref(thr)), "cvb200_otsu_threshold_from_histogram")
        return thr.value

    def threshold_global(self, rows, thr):
        cvb = self.cvb
        img = self._up(rows)
        h, w = img.shape
        out = torch.empty_like(img)
        cvb.check(cvb.lib().cvb200_threshold_global_dev(cvb.vp(img), cvb.sz(w), cvb.sz(h), cvb.sz(w), C.c_double(thr), cvb.vp(out), cvb.sz(1), cvb.sz(0), C.c_void_p(self.stream)), "cvb200_threshold_global_dev")
        torch.cuda.synchronize(self.dev)
        return out

    def fast_points(self, rows, threshold, n):
        cvb = self.cvb
        from compv_b200 import _ffi
        d = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
        d.setInt(_ffi.FAST_SET_INT_THRESHOLD, threshold)
        d.setInt(_ffi.FAST_SET_INT_FAST_TYPE, _ffi.FAST_TYPE_9 if n == 9 else _ffi.FAST_TYPE_12)
        d.setInt(_ffi.FAST_SET_INT_MAX_FEATURES, -1)
        d.setBool(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, True)
        return d.process(np.ascontiguousarray(rows))

    def _sht_obj(self, threshold):
        if threshold not in self._sht:
            self._sht[threshold] = self.cvb.CompVHough.newObj(self.cvb.HOUGHSHT_ID, 1.0, 1.0, threshold)
        return self._sht[threshold]

    def sht_accumulate(self, edges_strip, y0, full_height, threshold):
        cvb = self.cvb
        hgh = self._sht_obj(threshold)
        rows, w = edges_strip.shape
        n = C.c_size_t(0)
        cvb.check(cvb.lib().cvb200_hough_sht_acc_size(hgh._h, cvb.sz(w), cvb.sz(full_height), C.byref(n)), "cvb200_hough_sht_acc_size")
        acc = torch.zeros(n.value, dtype=torch.int32, device=self.dev)
        e = edges_strip.contiguous()
        cvb.check(cvb.lib().cvb200_hough_sht_accumulate_dev(hgh._h, cvb.vp(e), cvb.sz(w), cvb.sz(rows), cvb.sz(w), cvb.sz(full_height), cvb.sz(y0), cvb.vp(acc), C.c_void_p(self.stream)),
                  "cvb200_hough_sht_accumulate_dev")
        return acc

    def sht_lines(self, acc, width, full_height, threshold, capacity=1 << 16):
        cvb = self.cvb
        hgh = self._sht_obj(threshold)
        lines = np.zeros(capacity, cvb.LINE_DTYPE)
        cnt = C.c_size_t(0)
        cvb.check(cvb.lib().cvb200_hough_sht_lines_dev(hgh._h, cvb.vp(acc), cvb.sz(width), cvb.sz(full_height), cvb.vp(lines), cvb.sz(capacity), C.byref(cnt), C.c_void_p(self.stream)),
                  "cvb200_hough_sht_lines_dev")
        return lines[:min(cnt.value, capacity)].copy()
