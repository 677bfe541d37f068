"""Writes tests/golden/oracle_checksums.json: SHA-256 of the oracle's fields after a few steps of the
cases in tests/test_oracle.py::CASES.  These are REGRESSION fixtures of our own oracle (the reference
ships no golden vectors and cannot run here); they pin the oracle against accidental change and let the
GPU box verify the prebuilt liboracle_fluid.so behaves as it did in the build container.
    python tests/golden/make_oracle_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tests.test_oracle import CASES, _checksums, run_case  # noqa: E402

out = {}
for name, case in sorted(CASES.items()):
    o, s = run_case(oracle, case)
    out[name] = {"s_exec": s, "sha256": _checksums(o, oracle)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_checksums.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
