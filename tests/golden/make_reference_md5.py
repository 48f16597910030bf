"""Generates tests/golden/reference_md5.json from the UNMODIFIED reference compiled under oracle/_ref (run where /root/reference exists):
    python tests/golden/make_reference_md5.py
The reference runs single threaded (its multi-threaded Canny hysteresis is racy) with its x86 SIMD paths on."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from golden.cases import cases  # noqa: E402
from golden.adapters import OracleOrRef  # noqa: E402

if __name__ == "__main__":
    assert oracle.have_ref(), "oracle/_ref not built"
    out = cases(OracleOrRef("ref"))
    json.dump(out, open(os.path.join(HERE, "reference_md5.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
