"""Regenerates tests/golden/legacy_summary.json from the UNMODIFIED legacy SummaryGenerator of the variant module
(oracle/_ref/pv_ref_legacy, compiled from /root/reference/pepper_variant/modules/cpp/summary_generator.cpp). Run in the build
container:  python tests/golden/make_legacy_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import legacy_cases as LC  # noqa: E402

m = LC.ref_mod()
out = {name: LC.as_golden(LC.run_reference(m, b, chunk)) for name, (b, chunk) in LC.cases().items()}
with open(os.path.join(HERE, "legacy_summary.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print({k: (len(v["image"]), len(v["chunk_ids"])) for k, v in out.items()})
