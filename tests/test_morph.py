"""Section 8f next row 1, CompVMathMorph: oracle pinned on the compiled reference (CPU, bit-exact); CUDA vs oracle / reference (GPU)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_uniform, frame_text

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
SIZES = [(3, 3), (5, 5), (3, 5), (7, 3), (4, 4), (2, 3), (9, 9)]


def inputs(w, h):
    return [frame_g(w, h, 3), frame_uniform(w, h, 1), ((frame_text(w, h, 2) < 128) * 255).astype(np.uint8)]


def elements():
    for size in SIZES:
        for t in (oracle.STREL_RECT, oracle.STREL_DIAMOND, oracle.STREL_CROSS):
            if t == oracle.STREL_DIAMOND and size[0] != size[1]:
                continue   # the reference writes outside a non-square diamond
            yield size, t


@needs_ref
def test_oracle_strel_vs_reference():
    for size, t in elements():
        np.testing.assert_array_equal(oracle.morph_strel("orc", size, t), oracle.morph_strel("ref", size, t))


@needs_ref
@pytest.mark.parametrize("w,h", [(64, 48), (101, 37), (320, 200)])
@pytest.mark.parametrize("border", [0, 2])
def test_oracle_morph_vs_reference(w, h, border):
    for img in inputs(w, h):
        for size, t in elements():
            se = oracle.morph_strel("orc", size, t)
            for op in range(4):
                np.testing.assert_array_equal(oracle.morph("orc", img, se, op, border), oracle.morph("ref", img, se, op, border))


@needs_ref
def test_oracle_morph_1080p_multithreaded_reference_and_custom_element():
    img = ((frame_text(1920, 1080, 4) < 128) * 255).astype(np.uint8)
    se = oracle.morph_strel("orc", (3, 3), oracle.STREL_RECT)
    np.testing.assert_array_equal(oracle.morph("orc", img, se, oracle.MORPH_CLOSE), oracle.morph("ref", img, se, oracle.MORPH_CLOSE, threads=-1))  # samples/text_recognition/main.cxx:93-104
    custom = np.array([[1, 0, 0, 7], [0, 0, 2, 0], [0, 255, 0, 0]], np.uint8)   # any non-zero cell counts
    small = frame_g(90, 70, 5)
    for op in range(4):
        np.testing.assert_array_equal(oracle.morph("orc", small, custom, op), oracle.morph("ref", small, custom, op))


def test_oracle_morph_rejects_bad_arguments():
    img = frame_g(32, 32, 1)
    with pytest.raises(Exception):
        oracle.morph("orc", img, np.zeros((3, 3), np.uint8), 0)           # all-zero element
    with pytest.raises(Exception):
        oracle.morph("orc", img, np.ones((40, 3), np.uint8), 0)          # element taller than the image
    with pytest.raises(Exception):
        oracle.morph("orc", img, np.ones((3, 3), np.uint8), 4)           # GRADIENT: not implemented by the reference either


# ---------------------------------------------------------------- CUDA (C ABI) vs oracle / reference
@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 48), (101, 37), (320, 200), (1920, 1080)])
@pytest.mark.parametrize("border", [0, 2, 1])
def test_cuda_morph(cvb, w, h, border):
    for img in inputs(w, h)[:2 if w > 1000 else 3]:
        for size, t in elements():
            se = cvb.morph_strel(size, t)
            np.testing.assert_array_equal(se, oracle.morph_strel("orc", size, t))
            for op in range(4):
                got = cvb.morph(img, se, op, border, fill=77)
                want = oracle.morph("orc", img, se, op, border, fill=77)
                np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_cuda_morph_custom_element_strided_batch_and_errors(cvb):
    import torch
    from compv_b200 import _ffi
    custom = np.array([[1, 0, 0, 7], [0, 0, 2, 0], [0, 255, 0, 0]], np.uint8)
    img = np.zeros((70, 99), np.uint8)
    img[:, :90] = frame_g(90, 70, 5)
    img[:, 90:] = 200
    for op in range(4):
        got = cvb.morph(img, custom, op, width=90)
        np.testing.assert_array_equal(got[:, :90], oracle.morph("orc", img, custom, op, width=90)[:, :90])
    w, h, batch = 640, 360, 5
    frames = np.stack([((frame_text(w, h, k) < 128) * 255).astype(np.uint8) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros_like(d_in)
    se = cvb.morph_strel((5, 5), 2)
    cvb.morph_dev(d_in, w, h, w, se, 3, d_out, batch=batch, stream=torch.cuda.current_stream().cuda_stream)
    out = d_out.cpu().numpy()
    for k in range(batch):
        np.testing.assert_array_equal(out[k], oracle.morph("orc", frames[k], se, 3))
    lib = cvb.lib()
    small = frame_g(32, 32, 1)
    o = np.zeros_like(small)
    z = np.zeros((3, 3), np.uint8)
    one = np.ones((3, 3), np.uint8)
    call = lambda strel, op, b=2: lib.cvb200_morph_process(_ffi.vp(small), _ffi.sz(32), _ffi.sz(32), _ffi.sz(32), _ffi.vp(strel), _ffi.sz(strel.shape[1]), _ffi.sz(strel.shape[0]), _ffi.sz(strel.shape[1]), _ffi.vp(o), op, b)
    assert call(z, 0) == _ffi.E_INVALID_PARAMETER                      # compv_math_morph.cxx:483
    assert call(np.ones((40, 3), np.uint8), 0) == _ffi.E_INVALID_PARAMETER   # :135
    assert call(one, 4) == _ffi.E_NOT_IMPLEMENTED                      # :119-122
    assert call(one, 0, 7) == _ffi.E_NOT_IMPLEMENTED                   # :607
