#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: Mpixels/s, Canny+HoughKHT @1080p; % HBM roofline; vs ref AVX2 CPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic 1920x1080 frames (BASELINE.json configs[1] + configs[3], the
configuration the metric "Canny+HoughKHT @1080p" is quoted on): Gaussian 5x5 (sigma 1) -> Canny (Sobel 3x3, tLow 59, tHigh 119)
-> HoughKHT (rho 1, theta 1 degree, threshold 100) per frame.  Frames are sharded across ranks with no data-path collective
(weak scaling: every rank owns `--frames` frames).

  value : whole-job Mpixels/s with the frames already resident in HBM (device API, CUDA events, max over ranks)
  e2e   : the same metric through the host call (cvb200_canny_kht_process_batch): pinned host frames in, lines out,
          H2D and D2H inside the timed region
  roofline / cpu_baseline : see DESIGN.md section "Measurement"

--impl reference times the UNMODIFIED reference (oracle/_ref, AVX2 intrinsics, all host threads) on the same config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# The pipeline keeps several sub-batches in flight on their own streams; with the default of 8 hardware work queues the driver aliases streams onto
# shared queues and serialises them (measured: slots 4 and 5 of 6 only started when slots 0 and 1 had finished).  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

W, H = 1920, 1080
TLOW, THIGH, KS = 59.0, 119.0, 3
BLUR, SIGMA = 5, 1.0
KHT_THRESHOLD = 100
# Canny 1 B/px read (frame) + 1 B/px written (edge map), blur/gradients on chip; KHT 1 B/px read of the edge map
ALG_BYTES_PER_PX = {"canny_front": 2.0, "kht_link": 1.0, "kht_bits": 1.0, "canny_hysteresis": 2.0, "canny_finalize": 2.0}
METRIC = "Mpixels/s Canny+HoughKHT @1080p"
WORKLOAD = "gauss5x5+canny+houghkht_1080p"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_frames(n, seed0):
    from frames import frame_g
    return np.stack([frame_g(W, H, seed0 + k) for k in range(n)])


def tile_frames(dst, base):
    """dst[k] = base[k % len(base)] for every frame of dst (the synthetic batch cycles a few dozen distinct frames); returns dst."""
    n = len(base)
    for s0 in range(0, len(dst), n):
        m = min(n, len(dst) - s0)
        dst[s0:s0 + m] = base[:m]
    return dst


def run_reference(args, rank, world):
    """--impl reference: the unmodified reference (oracle/_ref) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    frames_per_step = max(1, min(args.frames, 8))  # bounded sample of the workload per step
    frames = make_frames(frames_per_step, 12345)
    r = oracle.ref(-1)
    threads = r.ref_threads_count()
    sess = oracle.RefEdgeSession(frames, "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=-1, kht_threshold=KHT_THRESHOLD)

    def step():
        return sess.run(0, frames_per_step)[0]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    mpx = frames_per_step * args.steps * W * H / 1e6
    value = mpx / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16 (+f32 blur)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "width": W, "height": H, "frames_per_step": frames_per_step,
                                         "tLow": TLOW, "tHigh": THIGH, "kernSize": KS, "blur": [BLUR, SIGMA], "kht": {"rho": 1, "theta_deg": 1, "threshold": KHT_THRESHOLD}},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": threads, "kind": "reference",
                         "sample": "%d frames/step x %d steps, CompV reference AVX2 intrinsics (asm disabled), %d pool threads on %d host cores" % (frames_per_step, args.steps, threads, cores)},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:  # context: the same reference deployed frame-parallel (one single-threaded session per core); `value` above stays the reference's own threading
        line["cpu_baseline_frame_parallel"] = cpu_frame_parallel()
    except Exception as ex:
        line["cpu_baseline_frame_parallel"] = {"value": None, "sample": "unavailable: %r" % (ex,)}
    print(json.dumps(line), flush=True)


def cpu_worker(n_frames):
    """Hidden mode (--cpu-worker N): ONE single-threaded reference session over N 1080p frames; prints elapsed ms.  bench.py starts one of these per host core to
    measure the frame-parallel CPU deployment (N independent 1-thread sessions), the strategy that suits the reference's poorly scaling KHT best."""
    import oracle
    frames = make_frames(2, 12345)
    oracle.ref(1)
    sess = oracle.RefEdgeSession(frames, "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=1, kht_threshold=KHT_THRESHOLD)
    sess.run(0, 1)
    ms, _ = sess.run(0, n_frames)
    print("CPUWORKER %.3f" % ms, flush=True)


def cpu_frame_parallel(frames_per_worker=6):
    """One single-threaded reference process per host core, all at once: aggregate Mpixels/s (what a frame-parallel CPU deployment of the reference delivers)."""
    cores = os.cpu_count() or 1
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", str(frames_per_worker)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(cores)]
    worst = 0.0
    ok = 0
    for pr in procs:
        out, _ = pr.communicate(timeout=600)
        for ln in out.splitlines():
            if ln.startswith("CPUWORKER"):
                worst = max(worst, float(ln.split()[1]))
                ok += 1
    if not ok:
        return None
    return {"value": ok * frames_per_worker * W * H / 1e6 / (worst * 1e-3), "unit": "Mpixels/s", "cores": ok, "kind": "reference",
            "sample": "%d single-threaded reference processes x %d 1080p frames each, started together; slowest process %.1f ms" % (ok, frames_per_worker, worst)}


def measure_latency(cvb):
    """One 1080p frame through the per-frame host API the reference's samples use (samples/hough_lines/main.cxx:59,106): cvb200_edge_dete_process then
    cvb200_hough_process, host buffers in and out; median of 15 calls.  Beside it the reference, single-threaded, on the same frame."""
    frame = make_frames(1, 12345)[0]
    dete = cvb.CompVEdgeDete.newObj(cvb.CANNY_ID, TLOW, THIGH, KS)
    dete.set_preblur(BLUR, SIGMA)
    kht = cvb.CompVHough.newObj(cvb.HOUGHKHT_ID, 1.0, 1.0, KHT_THRESHOLD)
    edges = np.zeros_like(frame)
    tc, tk = [], []
    for it in range(18):
        t0 = time.perf_counter()
        dete.process(frame, edges=edges)
        t1 = time.perf_counter()
        lines = kht.process(edges, capacity=4096)
        t2 = time.perf_counter()
        if it >= 3:
            tc.append((t1 - t0) * 1e3)
            tk.append((t2 - t1) * 1e3)
    out = {"workload": "gauss5x5+canny then houghkht, ONE 1920x1080 frame, host buffers in/out, per-frame API", "canny_ms": float(np.median(tc)), "kht_ms": float(np.median(tk)),
           "total_ms": float(np.median(np.array(tc) + np.array(tk))), "lines": int(len(lines))}
    try:
        import oracle
        frames = frame[None].copy()
        s1 = oracle.RefEdgeSession(frames, "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=1, kht_threshold=KHT_THRESHOLD)
        s1.run(0, 1)
        out["reference_1thread_ms"] = s1.run(0, 5)[0] / 5
        sN = oracle.RefEdgeSession(frames, "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=-1, kht_threshold=KHT_THRESHOLD)
        sN.run(0, 1)
        out["reference_all_threads_ms"] = sN.run(0, 5)[0] / 5
    except Exception as ex:
        out["reference_1thread_ms"] = None
        out["reference_note"] = "unavailable: %r" % (ex,)
    return out


def measure_rows(cvb, peak):
    """The other configurations BASELINE.json names, device-resident inputs, CUDA events, one figure each with its roofline fraction (algorithmic bytes of
    SURVEY 8(d) / time of the whole call / measured HBM peak) and the reference's time per frame on this host beside it.  Bounded: a few seconds each."""
    import torch
    from compv_b200 import _ffi
    from frames import frame_g, frame_text
    stream = torch.cuda.current_stream().cuda_stream
    rows = {}

    def timed(fn, steps=5, warmup=3):
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

    def dev_frames(w, h, batch, gen):
        distinct = np.stack([gen(k) for k in range(min(batch, 4))])
        d = torch.from_numpy(distinct).cuda()
        return distinct, d.repeat((batch + len(distinct) - 1) // len(distinct), 1, 1)[:batch].contiguous()

    def row(name, w, h, batch, alg_bytes_per_px, ms, cpu_ms_per_frame, cpu_note, extra=None):
        px = batch * w * h
        gbs = alg_bytes_per_px * px / (ms * 1e-3) / 1e9
        r = {"workload": name, "width": w, "height": h, "frames": batch, "ms_per_batch": ms, "value": px / ms / 1e3, "unit": "Mpixels/s",
             "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "alg_bytes_per_px": alg_bytes_per_px, "scope": "whole call (all its kernels)"},
             "cpu_baseline": {"ms_per_frame": cpu_ms_per_frame, "value": (w * h / 1e6 / (cpu_ms_per_frame * 1e-3)) if cpu_ms_per_frame else None, "unit": "Mpixels/s", "kind": "reference", "sample": cpu_note}}
        if extra:
            r.update(extra)
        return r

    try:
        import oracle
        have_ref = oracle.have_ref()
    except Exception:
        oracle, have_ref = None, False

    # ---- config 1: Sobel 3x3, 640x480, single frame, reference on 1 thread ----
    w, h = 640, 480
    img, d_in = dev_frames(w, h, 1, lambda k: frame_g(w, h, 12345 + k))
    d_out = torch.empty_like(d_in)
    sob = cvb.CompVEdgeDete.newObj(cvb.SOBEL_ID)
    ms = timed(lambda: sob.process_dev(d_in, w, h, w, d_out, batch=1, stream=stream), steps=20)
    host = []
    for it in range(13):
        t0 = time.perf_counter()
        sob.process(img[0])
        if it >= 3:
            host.append((time.perf_counter() - t0) * 1e3)
    cpu = float(np.median(oracle.time_edge_dete(img[0], kind="sobel", iters=10, threads=1)[0])) if have_ref else None
    rows["config1_sobel_640x480_single_frame"] = row("Sobel 3x3 edge detector (gx, gy, |gx|+|gy|, frame max, normalise)", w, h, 1, 2.0, ms, cpu, "CompVEdgeDete SOBEL, 1 thread, median of 10",
                                                     {"host_api_latency_ms": float(np.median(host))})

    # ---- config 3: FAST9_16 + NMS, 3840x2160, 64 frames ----
    w, h, batch = 3840, 2160, 64
    img4k, d4k = dev_frames(w, h, batch, lambda k: frame_g(w, h, 12345 + k))
    cap = 16384
    d_pts = torch.empty((batch, cap, 6), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(batch, dtype=torch.int32, device="cuda")
    fast = cvb.CompVCornerDete.newObj(_ffi.FAST_ID)
    fast.setInt(_ffi.FAST_SET_INT_THRESHOLD, 20)
    fast.setBool(_ffi.FAST_SET_BOOL_NON_MAXIMA_SUPP, True)
    ms = timed(lambda: fast.process_dev(d4k, w, h, w, d_pts, cap, d_cnt, batch=batch, stream=stream))
    cpu = float(np.median(oracle.fast_detect("ref", img4k[0], 9, 20, True, threads=-1, iters=5)[1])) if have_ref else None
    rows["config3_fast9_16_nms_4k_x64"] = row("FAST9_16 threshold 20 + 3x3 NMS, points out", w, h, batch, 1.0, ms, cpu, "CompVCornerDeteFAST, all host threads, median of 5",
                                              {"corners_frame0": int(d_cnt[0].item())})

    # ---- Canny + KHT at 3840x2160 (the headline metric's second size) ----
    batch = 256
    d_big = d4k.repeat(batch // 64, 1, 1).contiguous()
    dete = cvb.CompVEdgeDete.newObj(cvb.CANNY_ID, TLOW, THIGH, KS)
    dete.set_preblur(BLUR, SIGMA)
    kht = cvb.CompVHough.newObj(cvb.HOUGHKHT_ID, 1.0, 1.0, KHT_THRESHOLD)
    lines_buf = np.zeros((batch, 512), cvb.LINE_DTYPE)
    counts_buf = np.zeros(batch, np.uint64)
    a = (dete._h, kht._h, cvb.vp(d_big), cvb.sz(w), cvb.sz(h), cvb.sz(w), cvb.sz(batch), cvb.sz(h * w), cvb.vp(lines_buf), cvb.sz(512), cvb.vp(counts_buf), C.c_void_p(stream))
    ms = timed(lambda: cvb.check(cvb.lib().cvb200_canny_kht_process_batch_dev(*a), "canny_kht"), steps=3, warmup=2)
    cpu = None
    if have_ref:
        sess = oracle.RefEdgeSession(img4k[:2].copy(), "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=-1, kht_threshold=KHT_THRESHOLD)
        sess.run(0, 1)
        cpu = sess.run(0, 4)[0] / 4
    rows["canny_kht_4k_x256"] = row("gauss5x5+canny+houghkht (the headline path) at 3840x2160", w, h, batch, 3.0, ms, cpu, "reference, all host threads, 4 frames",
                                    {"lines_frame0": int(counts_buf[0])})
    del d_big

    # ---- config 5: S-HOG 8/16/8, 9 bins + PLSL + MSER at 3840x2160 ----
    hog = cvb.CompVHOG.newObj()
    n = hog.descriptorSize(w, h)
    d_desc = torch.empty((batch // 4, n), dtype=torch.float32, device="cuda")
    ms = timed(lambda: hog.process_dev(d4k, w, h, w, d_desc, batch=64, stream=stream))
    cpu = None
    if have_ref:
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            oracle.hog("ref", img4k[0], threads=-1)
            t.append((time.perf_counter() - t0) * 1e3)
        cpu = min(t)
    rows["config5_hog_s_4k_x64"] = row("S-HOG 8x8 cells, 16x16 blocks, stride 8, 9 bins, L2Hys, bilinear; frame = one window", w, h, 64, 3.25, ms, cpu, "CompVHogStd, all host threads, best of 3",
                                       {"descriptor_floats": int(n)})
    del d_desc
    txt, d_txt = dev_frames(w, h, 64, lambda k: ((frame_text(w, h, 20 + k) < 128) * 255).astype(np.uint8))
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    na = []
    ms = timed(lambda: na.__setitem__(slice(None), ccl.process_dev(d_txt, w, h, w, batch=64, stream=stream)[0]))
    cpu = float(np.median(oracle.ccl_lsl("ref", txt[0], threads=-1, iters=5)["ms"])) if have_ref else None
    rows["config5_plsl_4k_x64"] = row("PLSL connected components (text frame, dark glyphs = foreground), segments + label ids out", w, h, 64, 1.0, ms, cpu, "CompVConnectedComponentLabelingLSL, all host threads, median of 5",
                                      {"labels_frame0": int(na[0])})
    mser = cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8)
    d8 = d4k[:8].contiguous()
    ms = timed(lambda: na.__setitem__(slice(None), mser.process_dev(d8, w, h, w, batch=8, want_results=True, stream=stream)[0]), steps=2, warmup=1)
    cpu = float(np.median(oracle.ccl_lmser("ref", img4k[0], threads=-1, iters=2)["ms"])) if have_ref else None
    rows["config5_mser_4k_x8"] = row("LMSER regions (delta 2, unittests/ccl_mser.cxx parameters), regions + points returned to the host", w, h, 8, 1.0, ms, cpu, "CompVConnectedComponentLabelingLMSER, all host threads, 2 runs",
                                     {"regions_frame0": int(na[0])})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per rank per step (the KHT linking stage runs one warp per frame: throughput grows with frames in flight; 4096 x 1080p = 8.5 GB)")
    ap.add_argument("--cpu-worker", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-rows", action="store_true", help="skip the extra BASELINE configurations (rows) and the single-frame latency")
    ap.add_argument("--cpu-frames", type=int, default=400, help="frames timed for cpu_baseline (rank 0, N=1)")
    args = ap.parse_args()
    if args.cpu_worker:
        cpu_worker(args.cpu_worker)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import compv_b200 as cvb
    from compv_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    cvb.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not land on stdout next to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.frames
    distinct = min(B, 64)
    base = make_frames(distinct, shard.weak_seed(12345, rank))
    try:  # the batch (B, H, W) uint8 -- larger than the 126 MB L2 -- is tiled straight into pinned memory: no pageable copy of the 8.5 GB beside it
        h_in = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
        tile_frames(h_in.numpy(), base)
        host_batch = "tiled into pinned memory"
    except Exception:
        h_in = torch.from_numpy(tile_frames(np.empty((B, H, W), np.uint8), base)).pin_memory()
        host_batch = "pageable array copied into pinned memory"
    d_in = h_in.cuda()
    frames = base[:8].copy()  # the CPU baseline below cycles 8 frames
    dete = cvb.CompVEdgeDete.newObj(cvb.CANNY_ID, TLOW, THIGH, KS)
    dete.set_preblur(BLUR, SIGMA)
    kht = cvb.CompVHough.newObj(cvb.HOUGHKHT_ID, 1.0, 1.0, KHT_THRESHOLD)
    stream = torch.cuda.current_stream().cuda_stream
    nlines = [0]

    # result buffers are allocated once, as an application would: the calls below are the C ABI entry points themselves
    CAP = 512
    lines_buf = np.zeros((B, CAP), cvb.LINE_DTYPE)
    counts_buf = np.zeros(B, np.uint64)
    h_frames = h_in.numpy()
    vp, sz = cvb.vp, cvb.sz
    args_dev = (dete._h, kht._h, vp(d_in), sz(W), sz(H), sz(W), sz(B), sz(H * W), vp(lines_buf), sz(CAP), vp(counts_buf), C.c_void_p(stream))
    args_e2e = (dete._h, kht._h, vp(h_frames), sz(W), sz(H), sz(W), sz(B), sz(H * W), vp(lines_buf), sz(CAP), vp(counts_buf))

    def step_dev():
        cvb.check(cvb.lib().cvb200_canny_kht_process_batch_dev(*args_dev), "cvb200_canny_kht_process_batch_dev")
        nlines[0] = int(counts_buf.sum())

    def step_e2e():
        cvb.check(cvb.lib().cvb200_canny_kht_process_batch(*args_e2e), "cvb200_canny_kht_process_batch")
        nlines[0] = int(counts_buf.sum())

    dev = torch.device("cuda", local_rank)

    def barrier():
        shard.barrier(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        dev_ms_ = e0.elapsed_time(e1)
        barrier()
        dev_max, wall_max = shard.max_over_ranks([dev_ms_, wall], dev)
        return dev_max, wall_max

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = cvb.launch_count()
    dev_ms, _ = timed(step_dev, args.steps)
    launches = cvb.launch_count() - l0
    # ---- end to end through the host call (H2D + D2H inside) ----
    for _ in range(2):
        step_e2e()
    _, e2e_ms = timed(step_e2e, args.steps)   # the host call is synchronous: wall clock around it is the honest number
    clocks = sampler.stop() if rank == 0 else None

    value = shard.whole_job_throughput(B * W * H, args.steps, world, dev_ms) / 1e6
    e2e_value = shard.whole_job_throughput(B * W * H, args.steps, world, e2e_ms) / 1e6
    lines_all = shard.gather_counts(nlines[0], dev)

    # ---- roofline of the dominant kernel: per-kernel CUDA-event timing over extra steps (rank 0) ----
    roof, kernels = None, {}
    if rank == 0:
        # Per-kernel CUDA events.  The timed region above overlaps sub-batches on several streams, where a kernel's event-to-event time includes waiting for the others;
        # for the roofline the same steps are run once more with ONE slot (CVB200_PIPE_SLOTS=1: every kernel launched once per step on the whole batch, nothing overlapped).
        os.environ["CVB200_PIPE_SLOTS"] = "1"
        for _ in range(2):
            step_dev()
        cvb.lib().cvb200_profile_begin()
        psteps = min(args.steps, 5)
        for _ in range(psteps):
            step_dev()
        buf = C.create_string_buffer(1 << 16)
        cvb.check(cvb.lib().cvb200_profile_end(buf, C.c_size_t(len(buf))), "cvb200_profile_end")
        os.environ.pop("CVB200_PIPE_SLOTS", None)
        for ln in buf.value.decode().splitlines():
            name, cnt, ms = ln.split()
            kernels[name] = {"launches": int(cnt), "total_ms": float(ms), "avg_ms": float(ms) / max(int(cnt), 1)}
        step_ms = sum(k["total_ms"] for k in kernels.values()) / psteps
        top = max(kernels, key=lambda n: kernels[n]["total_ms"])
        for k in kernels.values():
            k["share"] = k["total_ms"] / psteps / step_ms
        peak, how = load_peaks()
        # the dominant kernel by device time; in this serial pass every launch of it processes the whole batch
        frames_per_launch = B * psteps / max(kernels[top]["launches"], 1)
        alg_bytes = ALG_BYTES_PER_PX.get(top, 1.0) * frames_per_launch * W * H
        achieved = alg_bytes / (kernels[top]["avg_ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # dram bytes read + written by that kernel in one `ncu --set full` capture (profiles/), per frame, scaled to this launch's frame count
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")))
            if top in tj["bytes_per_frame"]:
                traffic = tj["bytes_per_frame"][top] * frames_per_launch
                traffic_src = tj["source"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": how, "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": kernels[top]["avg_ms"],
                "frames_per_launch": frames_per_launch, "serial_step_ms": step_ms,
                "note": ("kht_link is the order-dependent linking walk (one warp per frame): bound by the issue latency of its dependent instruction chain, not by HBM, see DESIGN.md section 5. " if top == "kht_link" else "")
                + "Measured in a serial pass (one pipeline slot, no overlap) after the timed region; `kernels[*].share` is the share of that serial step."}
        if "canny_front" in kernels and top != "canny_front":
            fpl = B * psteps / max(kernels["canny_front"]["launches"], 1)
            cf = ALG_BYTES_PER_PX["canny_front"] * fpl * W * H / (kernels["canny_front"]["avg_ms"] * 1e-3) / 1e9
            roof["canny_front"] = {"achieved": cf, "frac": cf / peak, "avg_launch_ms": kernels["canny_front"]["avg_ms"], "frames_per_launch": fpl}

    # ---- CPU baseline: the compiled reference on this host's cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1:
        try:
            import oracle
            r = oracle.ref(-1)
            n = max(1, args.cpu_frames)
            sess = oracle.RefEdgeSession(frames[:8], "canny", TLOW, THIGH, KS, BLUR, SIGMA, threads=-1, kht_threshold=KHT_THRESHOLD)
            sess.run(0, 8)  # warm-up
            ms, _ = sess.run(0, n)
            cpu = {"value": n * W * H / 1e6 / (ms * 1e-3), "unit": "Mpixels/s", "cores": int(r.ref_threads_count()), "kind": "reference",
                   "sample": "%d x 1080p frames (Gaussian5x5 + Canny + HoughKHT, 8 distinct frames cycled), CompV reference, AVX2 intrinsics, asm disabled, %.3f ms/frame, %d host cores" % (n, ms / n, os.cpu_count())}
        except Exception as ex:  # the bench line must still come out
            cpu = {"value": None, "unit": "Mpixels/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (ex,)}

    cpu_fp, latency, rows = None, None, None
    if rank == 0 and world == 1:
        try:
            cpu_fp = cpu_frame_parallel()
        except Exception as ex:
            cpu_fp = {"value": None, "sample": "unavailable: %r" % (ex,)}
        if not args.no_rows:
            sampler2 = ClockSampler(local_rank)
            sampler2.start()
            try:
                latency = measure_latency(cvb)
            except Exception as ex:
                latency = {"error": repr(ex)}
            try:
                rows = measure_rows(cvb, load_peaks()[0])
            except Exception as ex:
                rows = {"error": repr(ex)}
            c2 = sampler2.stop()
            if isinstance(rows, dict):
                rows["clocks"] = c2

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16 (+f32 blur)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "frames_per_gpu_per_step": B, "tLow": TLOW, "tHigh": THIGH,
                       "kernSize": KS, "blur": [BLUR, SIGMA], "kht": {"rho": 1, "theta_deg": 1, "threshold": KHT_THRESHOLD}, "lines_per_step_per_rank": lines_all, "l2": "inputs (%.0f MB/GPU) larger than the 126 MB L2" % (B * W * H / 1e6), "host_batch": host_batch,
                       "parallelism": "frames sharded across %d GPU(s), no collective" % world},
            "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": B * W * H * world, "d2h_bytes_per_step": sum(lines_all) * 16,
                    "ms_per_step": e2e_ms / args.steps, "api": "cvb200_canny_kht_process_batch (pinned host frames in, lines out)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "cpu_baseline_frame_parallel": cpu_fp,
            "latency": latency, "rows": rows, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
