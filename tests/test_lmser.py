"""a12 linear-time MSER: oracle pinned on the compiled reference (CPU: same regions in the same order with the same point order and boxes);
CUDA vs oracle (GPU: same regions as SETS -- the component tree is canonical, the order in which a flood visits it is not)."""
import numpy as np
import pytest

import oracle
from frames import frame_g, frame_text, frame_smooth, frame_uniform

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")

PARAMS = [dict(),                                                                                   # unittests/ccl_mser.cxx:26-46
          dict(delta=5, min_area=0.0002, max_area=0.5, max_variation=0.5, min_diversity=0.5),      # compv_ccl.h:23-28 defaults
          dict(connectivity=4),
          dict(delta=1, min_area=0.0, max_area=1.0, max_variation=1.0, min_diversity=0.0)]           # everything the criteria let through


def gray_frames(w, h):
    rng = np.random.default_rng(w * 1000 + h)
    return [frame_g(w, h, 3), frame_text(w, h, 5), frame_smooth(w, h, 2), frame_uniform(w, h, 1), (rng.integers(0, 16, (h, w)) * 16).astype(np.uint8)]


def same_regions_in_order(a, b):
    np.testing.assert_array_equal(a["sizes"], b["sizes"])
    np.testing.assert_array_equal(a["boxes"], b["boxes"])
    for x, y in zip(a["points"], b["points"]):
        np.testing.assert_array_equal(x, y)


def canonical(res):
    """Order-free form: sorted list of (box, sorted point list) per region."""
    out = []
    for box, pts in zip(res["boxes"], res["points"]):
        key = np.sort(pts[:, 1].astype(np.int64) * 65536 + pts[:, 0].astype(np.int64))
        out.append((tuple(int(v) for v in box), key.tobytes()))
    return sorted(out)


@needs_ref
@pytest.mark.parametrize("w,h", [(64, 48), (100, 37), (320, 200), (333, 211)])   # 64 and 320: stride == width, the reference's row wrap-around (see oracle header)
@pytest.mark.parametrize("kw", PARAMS)
def test_oracle_lmser_vs_reference(w, h, kw):
    for img in gray_frames(w, h):
        same_regions_in_order(oracle.ccl_lmser("orc", img, **kw), oracle.ccl_lmser("ref", img, threads=1, **kw))


@needs_ref
def test_oracle_lmser_strided_and_multithreaded_reference():
    img = np.zeros((90, 160), np.uint8)
    img[:, :150] = frame_g(150, 90, 8)
    img[:, 150:] = 7   # padding beyond the width is never looked at
    same_regions_in_order(oracle.ccl_lmser("orc", img, width=150), oracle.ccl_lmser("ref", img, width=150, threads=1))
    big = frame_text(640, 360, 2)
    same_regions_in_order(oracle.ccl_lmser("orc", big), oracle.ccl_lmser("ref", big, threads=-1))


@needs_ref
def test_oracle_lmser_constant_frame():
    c = np.full((20, 30), 99, np.uint8)
    for kw in PARAMS:
        same_regions_in_order(oracle.ccl_lmser("orc", c, **kw), oracle.ccl_lmser("ref", c, threads=1, **kw))


def test_oracle_lmser_rejects_bad_parameters():
    img = np.zeros((8, 8), np.uint8)
    for kw in [dict(delta=0), dict(delta=256), dict(min_area=0.6, max_area=0.5), dict(max_variation=1.5), dict(min_diversity=-0.1), dict(connectivity=6)]:
        with pytest.raises(Exception):
            oracle.ccl_lmser("orc", img, **kw)


# ---------------------------------------------------------------- CUDA (C ABI) vs oracle
def new_mser(cvb, kw):
    from compv_b200 import _ffi
    d = dict(delta=2, min_area=0.0055 * 0.0055, max_area=0.8 * 0.15, max_variation=0.3, min_diversity=0.2, connectivity=8)
    d.update(kw)
    return cvb.CompVConnectedComponentLabeling.newObj(_ffi.LMSER_ID, **d)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 48), (100, 37), (320, 200), (333, 211)])
@pytest.mark.parametrize("kw", PARAMS)
def test_cuda_lmser(cvb, w, h, kw):
    ccl = new_mser(cvb, kw)
    for img in gray_frames(w, h):
        got = ccl.process(img)
        want = oracle.ccl_lmser("orc", img, **kw)
        assert got.labelsCount() == len(want["sizes"])
        assert canonical(got.regions()) == canonical(want)
        if len(want["sizes"]):
            np.testing.assert_array_equal(np.sort(got.boundingBoxes(), axis=0), np.sort(want["boxes"], axis=0))


@pytest.mark.gpu
def test_cuda_lmser_is_deterministic_in_region_order(cvb):
    img = frame_g(320, 200, 3)
    ccl = new_mser(cvb, {})
    a, b = ccl.process(img).regions(), ccl.process(img).regions()
    np.testing.assert_array_equal(a["sizes"], b["sizes"])
    np.testing.assert_array_equal(a["boxes"], b["boxes"])


@pytest.mark.gpu
def test_cuda_lmser_strided_constant_and_caps(cvb):
    import ctypes
    from compv_b200 import _ffi
    img = np.zeros((90, 160), np.uint8)
    img[:, :150] = frame_g(150, 90, 8)
    img[:, 150:] = 7
    ccl = new_mser(cvb, {})
    assert canonical(ccl.process(img, width=150).regions()) == canonical(oracle.ccl_lmser("orc", img, width=150))
    c = np.full((20, 30), 99, np.uint8)
    for kw in PARAMS:
        assert canonical(new_mser(cvb, kw).process(c).regions()) == canonical(oracle.ccl_lmser("orc", c, **kw))
    h = ctypes.c_void_p()
    bad = [dict(delta=0), dict(delta=256), dict(min_area=0.6, max_area=0.5), dict(max_variation=1.5), dict(min_diversity=-0.1), dict(connectivity=6)]
    for kw in bad:
        d = dict(delta=2, min_area=0.1, max_area=0.5, max_variation=0.3, min_diversity=0.2, connectivity=8)
        d.update(kw)
        rc = cvb.lib().cvb200_ccl_new_ex(ctypes.byref(h), _ffi.LMSER_ID, d["delta"], ctypes.c_double(d["min_area"]), ctypes.c_double(d["max_area"]), ctypes.c_double(d["max_variation"]),
                                         ctypes.c_double(d["min_diversity"]), d["connectivity"])
        assert rc == _ffi.E_INVALID_PARAMETER                                                # compv_ccl.cxx:76-83
    assert ccl.set(_ffi.CCL_SET_INT_CONNECTIVITY, 4, ctypes.c_int32) == _ffi.S_OK             # ccl_lmser.cxx:138-146 -> compv_ccl.cxx:29-35
    assert canonical(ccl.process(img, width=150).regions()) == canonical(oracle.ccl_lmser("orc", img, width=150, connectivity=4))
    assert ccl.set(_ffi.PLSL_SET_INT_TYPE, _ffi.PLSL_TYPE_XRLEZ, ctypes.c_int32) == _ffi.E_NOT_IMPLEMENTED
    with pytest.raises(Exception):
        ccl.process(img, width=150).debugFlatten()                                           # lmser_result.cxx:35-39


@pytest.mark.gpu
def test_cuda_lmser_batched_on_device(cvb):
    import torch
    w, h, batch = 640, 360, 4
    frames = np.stack([frame_text(w, h, 2 + k) if k % 2 == 0 else frame_g(w, h, 30 + k) for k in range(batch)])
    d_in = torch.from_numpy(frames).cuda()
    ccl = new_mser(cvb, {})
    na, results = ccl.process_dev(d_in, w, h, w, batch=batch, want_results=True, stream=torch.cuda.current_stream().cuda_stream)
    for k in range(batch):
        want = oracle.ccl_lmser("orc", frames[k])
        assert na[k] == len(want["sizes"])
        assert canonical(results[k].regions()) == canonical(want)
