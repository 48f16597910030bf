"""a11 PLSL connected component labeling: oracle pinned on the compiled reference (CPU, bit-exact label image / boxes / count);
CUDA vs oracle / reference (GPU).

Not compared with the reference: frames containing a row that alternates 1,0,1,...,1 over an ODD width.  Such a row has width+1 relative labels and
the reference writes RLCi[width], one element past its own RLC row (ccl_lsl.cxx:205 with a `width`-wide table, :627); the run end it stores there is
then overwritten by the next row.  The oracle and the CUDA path keep the correct end.  Random frames here use widths >= 9 where this does not occur
in practice, and the structured frames never alternate."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_text, frame_uniform

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


def binar_frames(w, h):
    out = [((frame_text(w, h, 7) < 128) * 255).astype(np.uint8),       # dark glyphs -> foreground
           ((frame_g(w, h, 5) > 128) * 255).astype(np.uint8),           # large blobs with ragged borders
           ((frame_uniform(w, h, 3) > 100) * 1).astype(np.uint8),      # noise, foreground coded 0x01: defeats every run-length shortcut
           ((frame_uniform(w, h, 4) > 200) * 255).astype(np.uint8)]   # sparse noise
    spiral = np.zeros((h, w), np.uint8)                                 # nested U shapes: labels that only merge at the bottom rows
    for k in range(0, min(w, h) // 2 - 2, 4):
        spiral[k:h - k, k] = 255
        spiral[k:h - k, w - 1 - k] = 255
        spiral[h - 1 - k, k:w - k] = 255
    out.append(spiral)
    return out


def same_result(a, b):
    assert a["na"] == b["na"]
    np.testing.assert_array_equal(a["labels"], b["labels"])
    np.testing.assert_array_equal(a["boxes"], b["boxes"])


@needs_ref
@pytest.mark.parametrize("w,h", [(64, 48), (321, 200), (640, 480), (1920, 1080)])
def test_oracle_lsl_vs_reference(w, h):
    for img in binar_frames(w, h):
        same_result(oracle.ccl_lsl("orc", img), oracle.ccl_lsl("ref", img, threads=1))


@needs_ref
def test_oracle_lsl_random_frames_vs_reference():
    rng = np.random.default_rng(1)
    for _ in range(150):
        h, w = int(rng.integers(1, 80)), int(rng.integers(9, 200))
        p = rng.choice([0.1, 0.3, 0.45, 0.5, 0.6, 0.7, 0.95])
        img = ((rng.random((h, w)) < p) * int(rng.choice([1, 255]))).astype(np.uint8)
        same_result(oracle.ccl_lsl("orc", img), oracle.ccl_lsl("ref", img, threads=1))


@needs_ref
def test_oracle_lsl_multithreaded_reference_agrees():
    img = ((frame_text(1122, 1182, 3) < 128) * 255).astype(np.uint8)
    same_result(oracle.ccl_lsl("orc", img), oracle.ccl_lsl("ref", img, threads=-1))


@needs_ref
def test_oracle_lsl_black_white_and_strided():
    z = np.zeros((40, 64), np.uint8)
    same_result(oracle.ccl_lsl("orc", z), oracle.ccl_lsl("ref", z, threads=1))
    f = np.full((24, 32), 255, np.uint8)
    same_result(oracle.ccl_lsl("orc", f), oracle.ccl_lsl("ref", f, threads=1))
    img = np.zeros((50, 96), np.uint8)
    img[:, :77] = ((frame_uniform(77, 50, 9) > 128) * 255).astype(np.uint8)
    img[:, 77:] = 255  # padding beyond the width must be ignored
    same_result(oracle.ccl_lsl("orc", img, width=77), oracle.ccl_lsl("ref", img, width=77, threads=1))


def test_oracle_lsl_segments_are_the_label_image():
    img = ((frame_g(200, 120, 11) > 128) * 255).astype(np.uint8)
    r = oracle.ccl_lsl("orc", img)
    lab = np.zeros_like(r["labels"])
    for j in range(120):
        for s in r["ranges"][r["row_offsets"][j]:r["row_offsets"][j + 1]]:
            lab[j, s["start"]:s["end"]] = s["a"]
    np.testing.assert_array_equal(lab, r["labels"])
    assert set(np.unique(r["labels"])) == set(range(0, r["na"] + 1))


# ---------------------------------------------------------------- CUDA (C ABI) vs oracle / reference
def check_cuda_result(res, want):
    assert res.labelsCount() == want["na"]
    np.testing.assert_array_equal(res.debugFlatten(), want["labels"])
    np.testing.assert_array_equal(res.boundingBoxes(), want["boxes"])
    ro, rg = res.segments()
    np.testing.assert_array_equal(ro, want["row_offsets"])
    for k in ("a", "start", "end"):
        np.testing.assert_array_equal(rg[k], want["ranges"][k])
    np.testing.assert_array_equal(res.labelIds(), np.arange(1, want["na"] + 1))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 48), (321, 200), (640, 480), (1920, 1080)])
def test_cuda_lsl(cvb, w, h):
    from compv_b200 import _ffi
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    for img in binar_frames(w, h):
        res = ccl.process(img)
        check_cuda_result(res, oracle.ccl_lsl("orc", img))
        if oracle.have_ref():
            r = oracle.ccl_lsl("ref", img, threads=1)
            assert res.labelsCount() == r["na"]
            np.testing.assert_array_equal(res.debugFlatten(), r["labels"])
            np.testing.assert_array_equal(res.boundingBoxes(), r["boxes"])


@pytest.mark.gpu
def test_cuda_lsl_random_frames(cvb):
    ccl = cvb.CompVConnectedComponentLabeling.newObj()
    rng = np.random.default_rng(2)
    for _ in range(120):
        h, w = int(rng.integers(1, 80)), int(rng.integers(1, 200))
        p = rng.choice([0.1, 0.3, 0.45, 0.5, 0.6, 0.7, 0.95])
        img = ((rng.random((h, w)) < p) * int(rng.choice([1, 255]))).astype(np.uint8)
        check_cuda_result(ccl.process(img), oracle.ccl_lsl("orc", img))


@pytest.mark.gpu
def test_cuda_lsl_caps_and_edge_cases(cvb):
    import ctypes
    from compv_b200 import _ffi
    h = ctypes.c_void_p()
    assert cvb.lib().cvb200_ccl_new(ctypes.byref(h), 12345) == _ffi.E_INVALID_PARAMETER
    ccl = cvb.CompVConnectedComponentLabeling.newObj(_ffi.PLSL_ID)
    assert ccl.set(_ffi.PLSL_SET_INT_TYPE, _ffi.PLSL_TYPE_XRLEZ, ctypes.c_int32) == _ffi.S_OK
    assert ccl.set(_ffi.PLSL_SET_INT_TYPE, _ffi.PLSL_TYPE_STD, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED      # ccl_lsl.cxx:136
    assert ccl.set(_ffi.PLSL_SET_BOOL_SORT_SEGMENTS, True, ctypes.c_bool) == _ffi.S_OK
    assert ccl.set(_ffi.PLSL_SET_BOOL_SORT_SEGMENTS, 1, ctypes.c_int32) == _ffi.E_INVALID_PARAMETER            # :141
    assert ccl.set(_ffi.CCL_SET_INT_CONNECTIVITY, 4, ctypes.c_int32) == _ffi.S_OK
    assert ccl.set(_ffi.CCL_SET_INT_CONNECTIVITY, 6, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED                # compv_ccl.cxx:33
    assert ccl.set(777, 1, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED
    z = np.zeros((40, 64), np.uint8)
    res = ccl.process(z)
    assert res.labelsCount() == 0 and len(res.boundingBoxes()) == 0 and not res.debugFlatten().any()
    f = np.full((24, 32), 255, np.uint8)
    check_cuda_result(ccl.process(f), oracle.ccl_lsl("orc", f))
    img = np.zeros((50, 99), np.uint8)   # stride 99: not a multiple of 4 -> byte-load variant; padding beyond the width is ignored
    img[:, :77] = ((frame_uniform(77, 50, 9) > 128) * 255).astype(np.uint8)
    img[:, 77:] = 255
    check_cuda_result(ccl.process(img, width=77), oracle.ccl_lsl("orc", img, width=77))
    # a frame with more provisional labels than the shared-memory EQ cache holds (isolated pixels on a 2x2 lattice + noise)
    big = np.zeros((600, 800), np.uint8)
    big[::2, ::2] = 255
    big[1::2] = ((frame_uniform(800, 300, 5) > 250) * 255).astype(np.uint8)
    check_cuda_result(ccl.process(big), oracle.ccl_lsl("orc", big))


@pytest.mark.gpu
def test_cuda_lsl_batched_on_device(cvb):
    import torch
    w, h, batch = 1122, 1182, 5
    frames = np.stack([((frame_text(w, h, 20 + k) < 128) * 255).astype(np.uint8) if k % 2 == 0 else ((frame_g(w, h, k) > 120) * 255).astype(np.uint8) for k in range(batch)])
    frames[3] = 0  # an empty frame in the middle of the batch
    d_in = torch.from_numpy(frames).cuda()
    d_labels = torch.empty((batch, h, w), dtype=torch.int32, device="cuda")
    ccl = cvb.CompVConnectedComponentLabeling.newObj()
    na, results = ccl.process_dev(d_in, w, h, w, batch=batch, d_labels=d_labels, want_results=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    labels = d_labels.cpu().numpy()
    for k in range(batch):
        want = oracle.ccl_lsl("orc", frames[k])
        assert na[k] == want["na"]
        np.testing.assert_array_equal(labels[k], want["labels"])
        check_cuda_result(results[k], want)
