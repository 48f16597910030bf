"""Row-strip mode (compv_b200/strips.py): one frame cut into one strip per rank.
CPU: world_size 2 and 3 gloo jobs run the partitioning / halo / seam-exchange / convergence / all-reduce logic with a numpy stand-in for the per-strip stages
(a stand-in with the same locality: 4 halo rows in, 8-connected closure, sums and maxima) and must reproduce the single-process result exactly.
GPU: two ranks drive the real stage entry points of libcompv_b200.so (NCCL when the box has two GPUs, else both ranks on GPU 0 with gloo staging the seams through
host memory) and every rank's strip must equal the oracle's full-frame result on the rows it owns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import ndimage

from compv_b200 import strips
from frames import frame_g, frame_smooth


def test_strip_ranges_cover_the_frame_once():
    for h in [1, 7, 480, 1080, 2160]:
        for world in [1, 2, 3, 8]:
            for align in [1, 8]:
                got = [strips.strip_range(h, r, world, align) for r in range(world)]
                assert got[0][0] == 0 and got[-1][1] == h
                for (a0, a1), (b0, b1) in zip(got, got[1:]):
                    assert a1 == b0 and a0 <= a1
                assert all(y0 % align == 0 or y0 == h for y0, _ in got)
    assert strips.with_halo(10, 20, 100, 4) == (6, 24) and strips.with_halo(0, 20, 22, 4) == (0, 22)


class NumpyOps:
    """Stand-in stages with the real ones' data dependencies: class map from a 9-row vertical window (4 halo rows), closure = 8-connected hysteresis."""

    def canny_front(self, rows, tlow, thigh, blur):
        f = rows.astype(np.int32)
        p = np.pad(f, ((4, 4), (1, 1)))           # zero border, like the sub-image border of the real kernels
        acc = sum(p[k:k + f.shape[0], 1:-1] * (5 - abs(k - 4)) for k in range(9)) + p[4:4 + f.shape[0], :-2] - p[4:4 + f.shape[0], 2:]
        g = np.abs(acc - 25 * f) % 251
        cls = np.where(g > thigh, 255, np.where(g > tlow, 128, 0)).astype(np.uint8)
        # the rows within 4 of a SUB-image border that is not the frame's border are garbage by construction (zero padding): exactly what the halo must absorb
        return torch.from_numpy(cls)

    def canny_closure(self, buf):
        b = buf.numpy()
        lab, n = ndimage.label(b > 0, structure=np.ones((3, 3)))
        if n:
            strong = np.unique(lab[b == 255])
            b[np.isin(lab, strong[strong > 0]) & (b > 0)] = 255

    def canny_finalize(self, buf):
        buf[buf != 255] = 0

    def edge_gmax(self, rows, kind):
        return torch.tensor([int(rows[1:-1].max()) if rows.shape[0] > 2 else 0], dtype=torch.int32)

    def edge_normalize(self, rows, kind, gmax):
        return torch.from_numpy((rows.astype(np.float32) * (255.0 / max(int(gmax), 1))).astype(np.uint8))

    def histogram(self, rows):
        return torch.from_numpy(np.bincount(rows.reshape(-1), minlength=256).astype(np.int32))

    def otsu_from_histogram(self, hist, count):
        return float(np.argmax(np.cumsum(hist) * 2 >= count))     # stand-in: the median

    def threshold_global(self, rows, thr):
        return torch.from_numpy(((rows > thr) * 255).astype(np.uint8))

    def sht_accumulate(self, edges_strip, y0, full_height, threshold):
        e = edges_strip.numpy()
        ys, xs = np.nonzero(e)
        acc = np.zeros((32, 2 * (e.shape[1] + full_height) + 1), np.int32)
        for t in range(32):
            rho = np.floor(xs * np.cos(t * np.pi / 32) + (ys + y0) * np.sin(t * np.pi / 32)).astype(np.int64) + e.shape[1] + full_height
            np.add.at(acc[t], rho, 1)
        return torch.from_numpy(acc.reshape(-1))

    def sht_lines(self, acc, width, full_height, threshold):
        a = acc.numpy()
        return np.nonzero(a > threshold)[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _single_process_expectation(frame):
    ops = NumpyOps()
    cls = ops.canny_front(frame, 60, 200, None)
    ops.canny_closure(cls)
    ops.canny_finalize(cls)
    hist = np.bincount(frame.reshape(-1), minlength=256)
    thr = ops.otsu_from_histogram(hist, frame.size)
    acc = ops.sht_accumulate(cls, 0, frame.shape[0], 10)
    return cls.numpy(), thr, acc.numpy()


def _cpu_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = NumpyOps()
        frame = frame_g(160, 150, 5)
        y0, y1, edges, rounds = strips.canny_row_strips(ops, frame, 60, 200, None)
        _, _, thr, binar = strips.otsu_row_strips(ops, frame)
        lines = strips.sht_row_strips(ops, edges, y0, frame.shape[0], 10)
        _, _, sob = strips.sobel_row_strips(ops, frame)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), y0=y0, y1=y1, edges=edges.numpy(), thr=thr, binar=binar.numpy(), lines=lines, rounds=rounds, sob=sob.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_strip_protocol_over_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_cpu_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    frame = frame_g(160, 150, 5)
    want_edges, want_thr, want_acc = _single_process_expectation(frame)
    assert (want_edges == 255).sum() > 500
    got = [np.load(tmp_path / ("r%d.npz" % r)) for r in range(world)]
    assert int(got[0]["y0"]) == 0 and int(got[-1]["y1"]) == frame.shape[0]
    edges = np.concatenate([g["edges"] for g in got])
    np.testing.assert_array_equal(edges, want_edges)                                  # seam closure == global closure
    assert max(int(g["rounds"]) for g in got) >= 2
    for g in got:
        assert float(g["thr"]) == want_thr                                            # summed histogram -> same threshold on every rank
        np.testing.assert_array_equal(g["lines"], np.nonzero(want_acc > 10)[0])       # summed accumulators -> same cells on every rank
    np.testing.assert_array_equal(np.concatenate([g["binar"] for g in got]), ((frame > want_thr) * 255).astype(np.uint8))
    gmax = max(int(frame[max(int(g["y0"]) - 1, 0):int(g["y1"]) + 1][1:-1].max()) for g in got)
    np.testing.assert_array_equal(np.concatenate([g["sob"] for g in got]), (frame.astype(np.float32) * (255.0 / gmax)).astype(np.uint8))


# ---------------------------------------------------------------- GPU: the real stages, two ranks
def _gpu_worker(rank, world, port, out_dir, use_nccl):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    device = rank if use_nccl else 0
    torch.cuda.set_device(device)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import compv_b200 as cvb
        cvb.init(device)
        ops = strips.CudaStripOps(device)
        out = {}
        for name, frame in (("g", frame_g(640, 480, 21)), ("s", frame_smooth(640, 480, 4))):
            y0, y1, edges, rounds = strips.canny_row_strips(ops, frame, 59.0, 119.0, (5, 1.0))
            lines = strips.sht_row_strips(ops, edges, y0, frame.shape[0], 120)
            _, _, sob = strips.sobel_row_strips(ops, frame, "sobel")
            _, _, thr, binar = strips.otsu_row_strips(ops, frame)
            pts = strips.fast_row_strips(ops, frame, 20, 9)
            out.update({name + "_y": np.array([y0, y1]), name + "_edges": edges.cpu().numpy(), name + "_rounds": rounds, name + "_lines": lines, name + "_sob": sob.cpu().numpy(),
                        name + "_thr": thr, name + "_binar": binar.cpu().numpy(), name + "_pts": pts})
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_row_strips_cuda_two_ranks_match_the_oracle(tmp_path):
    import oracle
    world, port = 2, _free_port()
    use_nccl = torch.cuda.device_count() >= 2
    mp.spawn(_gpu_worker, args=(world, port, str(tmp_path), use_nccl), nprocs=world, join=True)
    got = [np.load(tmp_path / ("r%d.npz" % r), allow_pickle=True) for r in range(world)]
    kern = oracle.gauss_kernel("orc", 5, 1.0)
    for name, frame in (("g", frame_g(640, 480, 21)), ("s", frame_smooth(640, 480, 4))):
        want_edges = oracle.edge_dete("orc", oracle.convlt1("orc", "8u32f8u", frame, kern, kern), "canny", 59.0, 119.0, 3)
        np.testing.assert_array_equal(np.concatenate([g[name + "_edges"] for g in got]), want_edges)          # strips + seam exchange == the whole frame
        want_lines, _ = oracle.hough_sht("orc", want_edges, 1.0, 1.0, 120)
        for g in got:                                                                                          # summed accumulators: every rank holds the full answer
            assert len(g[name + "_lines"]) == len(want_lines)
            for key in ("rho", "theta", "strength"):
                np.testing.assert_array_equal(g[name + "_lines"][key], want_lines[key])
        np.testing.assert_array_equal(np.concatenate([g[name + "_sob"] for g in got]), oracle.edge_dete("orc", frame, "sobel", 0.0, 0.0, 3))
        want_binar, want_thr = oracle.threshold("orc", "otsu", frame)
        for g in got:
            assert float(g[name + "_thr"]) == want_thr
        np.testing.assert_array_equal(np.concatenate([g[name + "_binar"] for g in got]), want_binar)
        want_pts = oracle.fast_detect("orc", frame, 9, 20, True)
        for g in got:
            assert len(g[name + "_pts"]) == len(want_pts)
            for key in ("x", "y", "strength"):
                np.testing.assert_array_equal(g[name + "_pts"][key], want_pts[key])
