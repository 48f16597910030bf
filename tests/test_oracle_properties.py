"""Property-style checks of the oracle against the compiled reference on generated inputs (hypothesis), plus size-independent properties of the oracle itself."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
SETTINGS = dict(max_examples=25, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])  # fixed example set: the suite must not depend on a seed


def binary_image(draw, min_w=9, max_w=96, max_h=40):
    w = draw(st.integers(min_w, max_w))
    h = draw(st.integers(1, max_h))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    density = draw(st.sampled_from([0.05, 0.3, 0.5, 0.7, 0.95]))
    fg = draw(st.sampled_from([1, 255]))
    rng = np.random.default_rng(seed)
    return ((rng.random((h, w)) < density) * fg).astype(np.uint8)


@needs_ref
@settings(**SETTINGS)
@given(st.data())
def test_plsl_oracle_equals_reference_on_random_binary_images(data):
    img = binary_image(data.draw)
    a, r = oracle.ccl_lsl("orc", img), oracle.ccl_lsl("ref", img, threads=1)
    assert a["na"] == r["na"]
    np.testing.assert_array_equal(a["labels"], r["labels"])
    np.testing.assert_array_equal(a["boxes"], r["boxes"])


@settings(**SETTINGS)
@given(st.data())
def test_plsl_oracle_labels_are_a_refinement_of_connectivity(data):
    """Whatever the reference's equivalence quirk does, a label never spans two 8-connected components and labels are 1..na with none missing."""
    from scipy import ndimage
    img = binary_image(data.draw)
    r = oracle.ccl_lsl("orc", img)
    truth, n = ndimage.label(img, structure=np.ones((3, 3)))
    assert r["na"] >= n
    assert set(np.unique(r["labels"]).tolist()) - {0} == set(range(1, r["na"] + 1))
    pairs = {(int(a), int(b)) for a, b in zip(r["labels"].ravel(), truth.ravel()) if a}
    assert len(pairs) == r["na"]                      # each label sits inside exactly one true component
    assert ((r["labels"] != 0) == (img != 0)).all()


@needs_ref
@settings(**SETTINGS)
@given(st.data())
def test_morph_oracle_equals_reference_on_random_elements(data):
    w, h = data.draw(st.integers(12, 80)), data.draw(st.integers(12, 50))
    sw, sh = data.draw(st.integers(1, 7)), data.draw(st.integers(1, 7))
    seed = data.draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    se = (rng.random((sh, sw)) < 0.6).astype(np.uint8) * rng.integers(1, 256, (sh, sw), dtype=np.uint8)
    if not se.any():
        se[sh // 2, sw // 2] = 1
    op, border = data.draw(st.integers(0, 3)), data.draw(st.sampled_from([0, 2]))
    np.testing.assert_array_equal(oracle.morph("orc", img, se, op, border), oracle.morph("ref", img, se, op, border))


@settings(**SETTINGS)
@given(st.data())
def test_morph_oracle_duality_and_idempotence(data):
    """Erosion and dilation are dual under complement for a symmetric element; opening and closing are idempotent away from the border band."""
    w, h = data.draw(st.integers(16, 60)), data.draw(st.integers(16, 40))
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 31 - 1)))
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    se = oracle.morph_strel("orc", (3, 3), data.draw(st.sampled_from([0, 2])))
    er, di = oracle.morph("orc", img, se, 0, 0), oracle.morph("orc", 255 - img, se, 1, 0)
    np.testing.assert_array_equal(er[2:-2, 1:-1], 255 - di[2:-2, 1:-1])
    for op in (2, 3):
        once = oracle.morph("orc", img, se, op, 2)
        twice = oracle.morph("orc", once, se, op, 2)
        np.testing.assert_array_equal(once[4:-4, 3:-3], twice[4:-4, 3:-3])


@needs_ref
@settings(max_examples=12, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(20, 120), st.integers(20, 90), st.integers(0, 2 ** 31 - 1), st.sampled_from([0.5, 1.0, 2.0, 3.0]), st.integers(5, 60))
def test_sht_oracle_equals_reference_on_random_edge_maps(w, h, seed, theta, threshold):
    rng = np.random.default_rng(seed)
    e = ((rng.random((h, w)) < 0.06) * 255).astype(np.uint8)
    e[h // 2, 2:w - 2] = 255
    a, na = oracle.hough_sht("orc", e, 1.0, theta, threshold, cap=1 << 18)
    r, nr = oracle.hough_sht("ref", e, 1.0, theta, threshold, threads=1, cap=1 << 18)
    assert na == nr
    for k in ("rho", "theta", "strength"):
        np.testing.assert_array_equal(a[k], r[k])


@needs_ref
@settings(max_examples=10, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(24, 90), st.integers(24, 70), st.integers(0, 2 ** 31 - 1), st.integers(1, 4), st.sampled_from([4, 8]))
def test_lmser_oracle_equals_reference_on_random_frames(w, h, seed, delta, conn):
    rng = np.random.default_rng(seed)
    img = (rng.integers(0, 12, (h, w)) * 20).astype(np.uint8)
    kw = dict(delta=delta, min_area=0.001, max_area=0.6, max_variation=0.7, min_diversity=0.3, connectivity=conn)
    a, r = oracle.ccl_lmser("orc", img, **kw), oracle.ccl_lmser("ref", img, threads=1, **kw)
    np.testing.assert_array_equal(a["sizes"], r["sizes"])
    np.testing.assert_array_equal(a["boxes"], r["boxes"])
    for x, y in zip(a["points"], r["points"]):
        np.testing.assert_array_equal(x, y)
