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
    d = {"row": row, "workload": extra.pop("workload"), "frames": batch, "ms_per_batch": round(ms, 4), "Mpixels_per_s": round(px / ms / 1e3, 1)}
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


ROWS = {"sht": row_sht, "lsl": row_lsl, "mser": row_mser}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="sht,lsl,mser")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    cvb.init(0)
    t0 = time.time()
    for r in args.rows.split(","):
        ROWS[r](args)
    print(json.dumps({"elapsed_s": round(time.time() - t0, 1)}))
