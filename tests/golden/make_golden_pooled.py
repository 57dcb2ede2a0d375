#!/usr/bin/env python3
"""Regenerate tests/golden/pooled_fst.json: the two inline python snippets of the UNMODIFIED reference wrapper
scripts/run_fst_impg.sh (:199-218) executed on a grid of 8-decimal pi texts -- build container only."""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refload  # noqa: E402

src = open(os.path.join(refload.REFERENCE_SCRIPTS, "run_fst_impg.sh")).read()
snips = re.findall(r"<<'PY'\n(.*?)\nPY\n", src, flags=re.S)
assert len(snips) == 2, len(snips)
grid = ["0.00000000", "0.00000001", "0.00000743", "0.00001234", "0.00012000", "0.00099999", "0.01000000", "0.12345678"]
cases = []
for a in grid:
    for b in grid[::2]:
        for c in grid[::3] + ["0.00000500"]:
            avg = subprocess.run([sys.executable, "-", a, b], input=snips[0], capture_output=True, text=True).stdout.strip()
            fst = subprocess.run([sys.executable, "-", a, b, c], input=snips[1], capture_output=True, text=True).stdout.strip()
            cases.append([a, b, c, avg, fst])
json.dump(cases, open(os.path.join(HERE, "pooled_fst.json"), "w"))
print(len(cases), "cases", cases[:2], cases[-1])
