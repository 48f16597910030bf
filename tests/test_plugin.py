"""integration/compv_b200_plugin.cxx: the CompVFeature::addFactory adapter compiled against the UNMODIFIED reference headers.  Unmodified reference API calls
(CompVEdgeDete::newObj(..., COMPV_CANNY_ID) ... process) must run on the B200 path once it is registered, and stay on the reference's CPU path when it cannot be."""
import json
import os
import subprocess
import sys

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "oracle", "_ref", "libcompv_b200_plugin.so")
needs_plugin = pytest.mark.skipif(not (oracle.have_ref() and os.path.exists(PLUGIN)), reason="oracle/_ref/libcompv_b200_plugin.so not built (needs /root/reference at build time)")


def drive():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plugin_driver.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


@needs_plugin
def test_plugin_without_a_gpu_registers_nothing():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = drive()
    assert out["register_rc"] == 20035                     # E_CUDA from cvb200_init: nothing registered
    assert out["gpu_launches"] == 0
    # the reference's own CPU implementations still answer (x86 build: Sobel shows the SSE4.1 max quirk)
    assert out["canny_equal"] and out["kht_equal"] and out["sht_equal"] and out["fast_equal"] and out["sobel_equal_x86_quirk"]
    assert out["hog_close"] and out["plsl_equal"] and out["plsl_extract_equal"] and out["plsl_segment_boxes_ok"] and out["mser_equal"]


@pytest.mark.gpu
@needs_plugin
def test_plugin_routes_the_reference_api_to_the_gpu():
    out = drive()
    assert out["register_rc"] == 0
    assert out["gpu_launches"] > 0                          # the reference's factory handed out the B200 objects
    assert out["canny_equal"] and out["kht_equal"] and out["sht_equal"] and out["fast_equal"]
    assert out["sobel_equal"]                               # true frame maximum: the B200 default, not the x86 SSE4.1 lane quirk
    # COMPV_HOGS_ID through CompVHOG::newObj, COMPV_PLSL_ID / COMPV_LMSER_ID through CompVConnectedComponentLabeling::newObj (+ the result classes' accessors)
    assert out["hog_size_equal"] and out["hog_bit_exact_vs_oracle"]
    assert out["plsl_equal"] and out["plsl_extract_equal"] and out["plsl_segment_boxes_ok"] and out["mser_equal"]
