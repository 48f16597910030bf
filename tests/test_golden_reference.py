"""Golden vectors generated from the UNMODIFIED compiled reference (tests/golden/make_reference_md5.py -> tests/golden/reference_md5.json, committed):
one MD5 per row on seeded 640x360 frames.  The oracle must reproduce every one of them on the CPU, the CUDA library on the GPU -- also where
/root/reference and oracle/_ref do not exist."""
import json
import os

import pytest

from golden.adapters import Cuda, OracleOrRef
from golden.cases import cases

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_md5.json")))


def test_oracle_reproduces_the_reference_goldens():
    got = cases(OracleOrRef("orc"))
    assert sorted(got) == sorted(GOLDEN)
    bad = {k: (got[k], GOLDEN[k]) for k in GOLDEN if got[k] != GOLDEN[k]}
    assert not bad, bad


@pytest.mark.gpu
def test_cuda_reproduces_the_reference_goldens(cvb):
    got = cases(Cuda(cvb))
    bad = {k: (got[k], GOLDEN[k]) for k in GOLDEN if got[k] != GOLDEN[k]}
    assert not bad, bad
