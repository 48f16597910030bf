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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=512, help="frames per rank per step (the KHT linking stage runs one warp per frame: throughput grows with frames in flight)")
    ap.add_argument("--cpu-frames", type=int, default=400, help="frames timed for cpu_baseline (rank 0, N=1)")
    args = ap.parse_args()
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
    frames = np.concatenate([make_frames(distinct, shard.weak_seed(12345, rank))] * ((B + distinct - 1) // distinct))[:B]  # (B, H, W) uint8: larger than the 126 MB L2
    h_in = torch.from_numpy(frames).pin_memory()
    d_in = h_in.cuda()
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
        cvb.lib().cvb200_profile_begin()
        psteps = min(args.steps, 10)
        for _ in range(psteps):
            step_dev()
        buf = C.create_string_buffer(1 << 16)
        cvb.check(cvb.lib().cvb200_profile_end(buf, C.c_size_t(len(buf))), "cvb200_profile_end")
        for ln in buf.value.decode().splitlines():
            name, cnt, ms = ln.split()
            kernels[name] = {"launches": int(cnt), "total_ms": float(ms), "avg_ms": float(ms) / max(int(cnt), 1)}
        step_ms = sum(k["total_ms"] for k in kernels.values()) / psteps
        top = max(kernels, key=lambda n: kernels[n]["total_ms"])
        for k in kernels.values():
            k["share"] = k["total_ms"] / psteps / step_ms
        peak, how = load_peaks()
        # the dominant kernel by device time; every launch of it processes the whole batch
        alg_bytes = ALG_BYTES_PER_PX.get(top, 1.0) * B * W * H
        achieved = alg_bytes / (kernels[top]["avg_ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # dram bytes read + written by that kernel in one `ncu --set full` capture of this command (profiles/), scaled to this launch's frame count
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
            if top in tj["bytes_per_launch"]:
                traffic = tj["bytes_per_launch"][top] * B / tj["frames_per_launch"]
                traffic_src = tj["source"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": how, "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": kernels[top]["avg_ms"],
                "note": "kht_link is the order-dependent linking walk (one warp per frame): latency-bound by construction, see DESIGN.md" if top == "kht_link" else ""}
        if "canny_front" in kernels and top != "canny_front":
            cf = ALG_BYTES_PER_PX["canny_front"] * B * W * H / (kernels["canny_front"]["avg_ms"] * 1e-3) / 1e9
            roof["canny_front"] = {"achieved": cf, "frac": cf / peak, "avg_launch_ms": kernels["canny_front"]["avg_ms"]}

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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16 (+f32 blur)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "frames_per_gpu_per_step": B, "tLow": TLOW, "tHigh": THIGH,
                       "kernSize": KS, "blur": [BLUR, SIGMA], "kht": {"rho": 1, "theta_deg": 1, "threshold": KHT_THRESHOLD}, "lines_per_step_per_rank": lines_all, "l2": "inputs (%.0f MB/GPU) larger than the 126 MB L2" % (B * W * H / 1e6),
                       "parallelism": "frames sharded across %d GPU(s), no collective" % world},
            "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": B * W * H * world, "d2h_bytes_per_step": sum(lines_all) * 16,
                    "ms_per_step": e2e_ms / args.steps, "api": "cvb200_canny_kht_process_batch (pinned host frames in, lines out)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
