"""Side measurements for the §8 rows that are not the headline workload (bench.py covers Gaussian+Canny+KHT).
Device-resident inputs, CUDA-event timing, per-kernel split through cvb200_profile_*; one JSON line per row on stdout.
  python scripts/bench_rows.py [--rows sht,lsl] [--batch 64] [--steps 5]
The CPU reference leg (oracle/_ref) runs on a few frames when the prebuilt shim is present."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compv_b200 as cvb  # noqa: E402
from compv_b200 import _ffi  # noqa: E402
from frames import frame_g, frame_text  # noqa: E402

W, H = 1920, 1080


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def kernel_split(fn, steps):
    cvb.lib().cvb200_profile_begin()
    for _ in range(steps):
        fn()
    buf = C.create_string_buffer(1 << 16)
    cvb.check(cvb.lib().cvb200_profile_end(buf, C.c_size_t(len(buf))), "cvb200_profile_end")
    out = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, ms = ln.split()
        out[name] = round(float(ms) / steps, 4)
    return out


def emit(row, batch, ms, extra):
    px = batch * W * H
    d = {"row": row, "workload": extra.pop("workload").replace("1080p", "%dx%d" % (W, H)), "frames": batch, "ms_per_batch": round(ms, 4), "Mpixels_per_s": round(px / ms / 1e3, 1)}
    d.update(extra)
    print(json.dumps(d), flush=True)


def edge_batch(batch):
    frames = np.stack([frame_g(W, H, 12345 + k) for k in range(min(batch, 8))])
    d_in = torch.from_numpy(frames).cuda()
    d_in = d_in.repeat((batch + len(frames) - 1) // len(frames), 1, 1)[:batch].contiguous()
    d_edges = torch.empty_like(d_in)
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)
    canny.process_dev(d_in, W, H, W, d_edges, batch=batch, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_edges


def row_sht(args):
    import oracle
    d_edges = edge_batch(args.batch)
    sht = cvb.CompVHough.newObj(_ffi.HOUGHSHT_ID, 1.0, 1.0, 150)
    stream = torch.cuda.current_stream().cuda_stream
    got = []

    def step():
        got[:] = sht.process_dev(d_edges, W, H, W, batch=args.batch, stream=stream, capacity=4096)
    ms = timed(step, args.steps)
    extra = {"workload": "houghsht_1080p rho=1 theta=1deg threshold=150 on Canny edge maps", "lines_frame0": len(got[0]), "kernels_ms": kernel_split(step, args.steps)}
    if oracle.have_ref():
        e = d_edges[0].cpu().numpy()
        _, _, t = oracle.hough_sht("ref", e, 1.0, 1.0, 150, threads=-1, iters=5)
        extra["cpu_reference_ms_per_frame"] = round(float(np.median(t)), 3)
    emit("a6", args.batch, ms, extra)


def row_lsl(args):
    import oracle
    frames = np.stack([((frame_text(W, H, 20 + k) < 128) * 255).astype(np.uint8) for k in range(min(args.batch, 8))])
    d_in = torch.from_numpy(frames).cuda()
    d_in = d_in.repeat((args.batch + len(frames) - 1) // len(frames), 1, 1)[:args.batch].contiguous()
    d_labels = torch.empty((args.batch, H, W), dtype=torch.int32, device="cuda")
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    stream = torch.cuda.current_stream().cuda_stream
    na = []

    def step_lea():
        na[:] = ccl.process_dev(d_in, W, H, W, batch=args.batch, stream=stream)[0]

    def step_flat():
        na[:] = ccl.process_dev(d_in, W, H, W, batch=args.batch, d_labels=d_labels, stream=stream)[0]
    ms = timed(step_lea, args.steps)
    ms_flat = timed(step_flat, args.steps)
    extra = {"workload": "plsl_1080p text frame (dark glyphs = foreground)", "labels_frame0": int(na[0]), "ms_per_batch_with_label_image": round(ms_flat, 4),
             "kernels_ms": kernel_split(step_flat, args.steps)}
    if oracle.have_ref():
        r = oracle.ccl_lsl("ref", frames[0], threads=-1, iters=10)
        extra["cpu_reference_ms_per_frame"] = round(float(np.median(r["ms"])), 3)
    emit("a11", args.batch, ms, extra)


def row_mser(args):
    import oracle
    batch = min(args.batch, 8)
    frames = np.stack([frame_g(W, H, 20 + k) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8)
    stream = torch.cuda.current_stream().cuda_stream
    na = []

    def step():
        na[:] = ccl.process_dev(d_in, W, H, W, batch=batch, want_results=True, stream=stream)[0]
    ms = timed(step, max(2, args.steps // 2), warmup=1)
    extra = {"workload": "lmser_1080p frame G, delta=2 (unittests/ccl_mser.cxx parameters), regions + points returned to the host", "regions_frame0": int(na[0]),
             "kernels_ms": kernel_split(step, 2)}
    if oracle.have_ref():
        r = oracle.ccl_lmser("ref", frames[0], threads=-1, iters=2)
        extra["cpu_reference_ms_per_frame"] = round(float(np.median(r["ms"])), 3)
    emit("a12", batch, ms, extra)


PEAK_GBS = None


def hbm_peak():
    global PEAK_GBS
    if PEAK_GBS is None:
        try:
            PEAK_GBS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            PEAK_GBS = 6574.5
    return PEAK_GBS


def g_frames(batch):
    frames = np.stack([frame_g(W, H, 12345 + k) for k in range(min(batch, 8))])
    d_in = torch.from_numpy(frames).cuda()
    return frames, d_in.repeat((batch + len(frames) - 1) // len(frames), 1, 1)[:batch].contiguous()


def simple_row(args, row, workload, alg_bytes_per_px, step, ref_ms=None, extra=None):
    """alg_bytes_per_px: SURVEY section 8(d)'s algorithmic bytes; achieved GB/s = bytes / time of the whole call (all its kernels)."""
    ms = timed(step, args.steps)
    px = args.batch * W * H
    gbs = alg_bytes_per_px * px / (ms * 1e-3) / 1e9
    d = {"workload": workload, "alg_bytes_per_px": alg_bytes_per_px, "achieved_GBps": round(gbs, 1), "hbm_peak_GBps": hbm_peak(), "frac_of_hbm_peak": round(gbs / hbm_peak(), 4),
         "kernels_ms": kernel_split(step, args.steps)}
    if ref_ms is not None:
        d["cpu_reference_ms_per_frame"] = round(float(ref_ms), 3)
    if extra:
        d.update(extra)
    emit(row, args.batch, ms, d)


def row_convlt(args):
    import oracle
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    k = cvb.gauss_kernel(5, 1.0)
    d_out8 = torch.empty_like(d_in)
    lib = cvb.lib()

    def step8():
        cvb.check(lib.cvb200_convlt1_8u32f8u_dev(_ffi.vp(d_in), _ffi.sz(W), _ffi.sz(H), _ffi.sz(W), _ffi.vp(k), _ffi.vp(k), _ffi.sz(5), _ffi.vp(d_out8), 0, _ffi.sz(args.batch), _ffi.sz(0), C.c_void_p(stream)), "convlt")
    ref = None
    if oracle.have_ref():
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            oracle.convlt1("ref", "8u32f8u", frames[0], k, k)
            t.append((time.perf_counter() - t0) * 1e3)
        ref = min(t)
    simple_row(args, "a2", "convlt1<u8,f32,u8> 5-tap Gaussian (sigma 1), 1080p", 2, step8, ref)
    sob = np.array([-1, 0, 1], np.int16)
    smooth = np.array([1, 2, 1], np.int16)
    d_out16 = torch.empty((args.batch, H, W), dtype=torch.int16, device="cuda")

    def step16():
        cvb.check(lib.cvb200_convlt1_8u16s16s_dev(_ffi.vp(d_in), _ffi.sz(W), _ffi.sz(H), _ffi.sz(W), _ffi.vp(smooth), _ffi.vp(sob), _ffi.sz(3), _ffi.vp(d_out16), 0, _ffi.sz(args.batch), _ffi.sz(0), C.c_void_p(stream)), "convlt")
    simple_row(args, "a2", "convlt1<u8,i16,i16> 3-tap Sobel gx, 1080p", 3, step16)


def row_sobel(args):
    import oracle
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    d_out = torch.empty_like(d_in)
    sob = cvb.CompVEdgeDete.newObj(_ffi.SOBEL_ID)

    def step():
        sob.process_dev(d_in, W, H, W, d_out, batch=args.batch, stream=stream)
    ref = float(np.median(oracle.time_edge_dete(frames[0], kind="sobel", iters=5, threads=-1)[0])) if oracle.have_ref() else None
    simple_row(args, "a3", "Sobel edge detector (gx, gy, L1 magnitude, frame max, normalise), 1080p", 2, step, ref)
    canny = cvb.CompVEdgeDete.newObj(_ffi.CANNY_ID, 59.0, 119.0, 3)
    canny.set_preblur(5, 1.0)

    def stepc():
        canny.process_dev(d_in, W, H, W, d_out, batch=args.batch, stream=stream)
    refc = float(np.median(oracle.time_edge_dete(frames[0], kind="canny", tlow=59.0, thigh=119.0, blur_size=5, blur_sigma=1.0, iters=5, threads=-1)[0])) if oracle.have_ref() else None
    simple_row(args, "a5", "Gaussian 5x5 + Canny 59/119, 1080p (frame G)", 2, stepc, refc)


def row_gradient(args):
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    mag = torch.empty((args.batch, H, W), dtype=torch.float32, device="cuda")
    dire = torch.empty_like(mag)
    lib = cvb.lib()

    def step():
        cvb.check(lib.cvb200_gradient_fast_8u_dev(_ffi.vp(d_in), _ffi.sz(W), _ffi.sz(H), _ffi.sz(W), None, None, None, None, _ffi.vp(mag), _ffi.vp(dire), _ffi.sz(args.batch), _ffi.sz(0), C.c_void_p(stream)), "gradient")
    simple_row(args, "a4", "CompVGradientFast magnitude + direction (f32 out), 1080p", 9, step)


def row_fast(args):
    import oracle
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    cap = 8192
    d_pts = torch.empty((args.batch, cap, 6), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(args.batch, dtype=torch.int32, device="cuda")
    fast = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
    fast.setInt(_ffi.FAST_SET_INT_THRESHOLD, 20)
    fast.setBool(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, True)

    def step():
        fast.process_dev(d_in, W, H, W, d_pts, cap, d_cnt, batch=args.batch, stream=stream)
    ref = None
    if oracle.have_ref():
        _, t = oracle.fast_detect("ref", frames[0], 9, 20, True, threads=-1, iters=5)
        ref = float(np.median(t))
    step()
    torch.cuda.synchronize()
    simple_row(args, "a8", "FAST9 threshold 20 + NMS, 1080p (frame G)", 1, step, ref, {"corners_frame0": int(d_cnt[0].item())})


def row_hog(args):
    import oracle
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    hog = cvb.CompVHOG.newObj()
    n = hog.descriptorSize(W, H)
    d_out = torch.empty((args.batch, n), dtype=torch.float32, device="cuda")

    def step():
        hog.process_dev(d_in, W, H, W, d_out, batch=args.batch, stream=stream)
    ref = None
    if oracle.have_ref():
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            oracle.hog("ref", frames[0], threads=-1)
            t.append((time.perf_counter() - t0) * 1e3)
        ref = min(t)
    simple_row(args, "a9", "S-HOG 8x8 cells, 16x16 blocks, stride 8, 9 bins, L2Hys, bilinear; 1080p as one window", 3.25, step, ref, {"descriptor_floats": int(n)})


def row_threshold(args):
    import oracle
    frames, d_in = g_frames(args.batch)
    stream = torch.cuda.current_stream().cuda_stream
    d_out = torch.empty_like(d_in)
    d_thr = torch.zeros(args.batch, dtype=torch.float64, device="cuda")
    d_hist = torch.zeros(args.batch * 256, dtype=torch.int32, device="cuda")
    lib = cvb.lib()

    def otsu():
        cvb.check(lib.cvb200_threshold_otsu_dev(_ffi.vp(d_in), _ffi.sz(W), _ffi.sz(H), _ffi.sz(W), _ffi.vp(d_thr), _ffi.vp(d_out), _ffi.vp(d_hist), _ffi.sz(args.batch), _ffi.sz(0), C.c_void_p(stream)), "otsu")

    def adaptive():
        cvb.check(lib.cvb200_threshold_adaptive_dev(_ffi.vp(d_in), _ffi.sz(W), _ffi.sz(H), _ffi.sz(W), _ffi.sz(5), C.c_double(8.0), C.c_double(255.0), 0, _ffi.vp(d_out), _ffi.sz(args.batch), _ffi.sz(0), C.c_void_p(stream)), "adaptive")
    r1 = r2 = None
    if oracle.have_ref():
        def tm(mode):
            t = []
            for _ in range(3):
                t0 = time.perf_counter()
                oracle.threshold("ref", mode, frames[0], threads=-1)
                t.append((time.perf_counter() - t0) * 1e3)
            return min(t)
        r1, r2 = tm("otsu"), tm("adaptive")
    simple_row(args, "a10", "Otsu threshold (histogram + threshold + binarise), 1080p", 3, otsu, r1)
    simple_row(args, "a10", "adaptive threshold, 5x5 fixed-point mean, delta 8, 1080p", 2, adaptive, r2)


def row_morph(args):
    import oracle
    frames = np.stack([((frame_text(W, H, 20 + k) < 128) * 255).astype(np.uint8) for k in range(min(args.batch, 8))])
    d_in = torch.from_numpy(frames).cuda()
    d_in = d_in.repeat((args.batch + len(frames) - 1) // len(frames), 1, 1)[:args.batch].contiguous()
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.current_stream().cuda_stream
    se = cvb.morph_strel((3, 3), 0)

    def step():
        cvb.morph_dev(d_in, W, H, W, se, 3, d_out, batch=args.batch, stream=stream)
    ref = None
    if oracle.have_ref():
        _, t = oracle.morph("ref", frames[0], se, 3, threads=-1, iters=5)
        ref = float(np.median(t))
    simple_row(args, "8f-1", "morphological close, 3x3 rectangle (text pipeline step between threshold and PLSL), 1080p", 4, step, ref)


ROWS = {"morph": row_morph, "convlt": row_convlt, "sobel": row_sobel, "gradient": row_gradient, "fast": row_fast, "hog": row_hog, "threshold": row_threshold, "sht": row_sht, "lsl": row_lsl, "mser": row_mser}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="convlt,sobel,gradient,fast,hog,threshold,morph,sht,lsl,mser")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--size", default="1080p", choices=["1080p", "4k"])
    args = ap.parse_args()
    if args.size == "4k":
        W, H = 3840, 2160
    cvb.init(0)
    t0 = time.time()
    for r in args.rows.split(","):
        ROWS[r](args)
    print(json.dumps({"elapsed_s": round(time.time() - t0, 1)}))
