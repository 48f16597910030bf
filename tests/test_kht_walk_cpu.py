"""CPU replays of the two order-dependent pieces of the KHT device path (no GPU needed):
  * compv_b200/csrc/kht_walk.cuh  -- the linking walker (Algorithm 5/6, houghkht.cxx:544-760) against a byte-map restatement of the reference's procedure;
  * compv_b200/csrc/std_sort_emu.cuh -- the device-side replacement of the reference's std::sort (houghkht.cxx:1195-1204), pinned on libstdc++'s std::sort itself.
Both headers are the very code the CUDA kernels compile (host+device functions)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(src, exe, arg):
    out = os.path.join(ROOT, "tests", "cpp", exe)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", out, os.path.join(ROOT, "tests", "cpp", src)])
    r = subprocess.run([out, str(arg)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_link_walker_matches_bytemap_procedure():
    assert "0 mismatches" in _build_and_run("link_check.cpp", "link_check", 220)


def test_sort_emulation_matches_std_sort():
    assert "0 mismatches" in _build_and_run("sort_check.cpp", "sort_check", 432)
