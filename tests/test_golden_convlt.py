"""The reference's own golden vectors for K1-K3 (unittests/math_convlt.cxx:17-26): oracle, compiled reference and CUDA library."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle

W, H, STRIDE, KS = 1285, 720, 1344, 7
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "convlt_md5.json")))


def build_case(name):
    """Input plane + kernel exactly as unittests/math_convlt.cxx:99-148 builds them."""
    tin, tk, tout = oracle.CONV_TYPES[name]
    i = np.arange(W, dtype=np.uint64)[None, :]
    j = np.arange(H, dtype=np.uint64)[:, None]
    ij = i * j
    data = np.zeros((H, STRIDE), tin)
    if tin == np.uint8:
        data[:, :W] = ((ij + 53) & 0xff).astype(np.uint8)
    elif tin == np.float32:
        sign = np.where((ij & 1) == 1, np.float32(1), np.float32(-1))
        data[:, :W] = ((ij.astype(np.float32) + np.float32(53.558)) * sign) / np.float32(1.2)
    else:
        v = (ij.astype(np.int64) + 53) * np.where((i & 1) == 1, -1, 1)
        data[:, :W] = (v & 0xffff).astype(np.uint16).view(np.int16)
    k = np.arange(KS)
    if name.startswith("fxp"):
        # the test hands the FLOAT gaussian kernel (sigma 3.5) to convlt1FixedPoint reinterpreted as uint16 taps (math_convlt.cxx:65-70,130-131)
        kern = oracle.gauss_kernel("orc", KS, 3.5).view(np.uint16)[:KS].copy()
    elif tk == np.float32:
        kern = ((k.astype(np.float32) + np.float32(53.558)) * np.where((k & 1) == 1, np.float32(1), np.float32(-1))) / np.float32(1.2)
    else:
        kern = ((k + 53) * np.where((k & 1) == 1, -1, 1)).astype(np.int16)
    return data, kern.astype(tk)


def md5_rows(out):
    """tests/tests_common.cxx:98-117: MD5 over rowInBytes of every row (stride padding skipped)."""
    return hashlib.md5(np.ascontiguousarray(out[:, :W]).tobytes()).hexdigest()


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_oracle_matches_reference_goldens(case):
    data, kern = build_case(case["name"])
    lib = oracle.orc()
    try:
        lib.orc_set_fma(1)
        assert md5_rows(oracle.convlt1("orc", case["name"], data, kern, kern, width=W)) == case["md5_fma"]
        lib.orc_set_fma(0)
        assert md5_rows(oracle.convlt1("orc", case["name"], data, kern, kern, width=W)) == case["md5"]
    finally:
        lib.orc_set_fma(1)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_compiled_reference_matches_its_goldens(case):
    data, kern = build_case(case["name"])
    assert md5_rows(oracle.convlt1("ref", case["name"], data, kern, kern, width=W)) == case["md5_fma"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_cuda_matches_reference_goldens(case, cvb):
    data, kern = build_case(case["name"])
    out = cvb.convlt1(case["name"], data, kern, kern, width=W)
    assert md5_rows(out) == case["md5_fma"]
