"""Regenerates tests/golden/kat_*.json from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference/pepper_variant/modules/cpp/region_summary.cpp). Run in the build container:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import helpers as H  # noqa: E402
import pyoracle as O  # noqa: E402

for name, make in H.KATS.items():
    b = make()
    a = O.ref_summary(b, 0, H.R9)
    g = dict(position=[int(x) for x in a["position"]], depth=[int(x) for x in a["depth"]],
             frequency=[int(x) for x in a["frequency"]], alleles=[x.decode("latin-1") for x in a["alleles"]],
             images=a["images"].astype(int).tolist())
    with open(os.path.join(HERE, "kat_%s.json" % name), "w") as f:
        json.dump(g, f, separators=(",", ":"))
    print(name, len(g["position"]), g["alleles"][:4])
