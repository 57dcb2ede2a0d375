"""Row f-1 (ingest): pure-Python restatement of the GFA v1 -> presence-matrix reader.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: the reference has no GFA reader of
its own -- it hands the extracted window graph to `odgi similarity -i tmp.gfa` (scripts/run_pica2_odgi.sh:96),
which is not under /root/reference.  The restatement follows the GFA v1 specification for S / P / W lines and
`odgi similarity`'s default of one group per path; it is the contract `impop_gfa_fill` is tested against.
"""
from __future__ import annotations

import re

import numpy as np


def parse(text: str):
    """-> (names, x [n, m] uint8 presence, counts [n, m] int64 visit counts, node_len [m] int64)."""
    seg_index, node_len, paths = {}, [], []
    for raw in text.split("\n"):
        line = raw[:-1] if raw.endswith("\r") else raw
        f = line.split("\t")
        if len(f) < 2 or len(f[0]) != 1:
            continue
        if f[0] == "S":
            seq = f[2]
            if seq == "*":
                ln = 0
                for tag in f[3:]:
                    if tag.startswith("LN:i:"):
                        ln = int(tag[5:])
                        break
            else:
                ln = len(seq)
            if f[1] in seg_index:
                raise ValueError("duplicate segment " + f[1])
            seg_index[f[1]] = len(node_len)
            node_len.append(ln)
        elif f[0] == "P":
            steps = [] if f[2] == "*" else [s[:-1] for s in f[2].split(",")]
            paths.append((f[1], steps))
        elif f[0] == "W":
            name = f"{f[1]}#{f[2]}#{f[3]}"
            if f[4] != "*" and f[5] != "*":
                name += f":{f[4]}-{f[5]}"
            steps = [] if f[6] == "*" else re.findall(r"[<>]([^<>]+)", f[6])
            paths.append((name, steps))
    n, m = len(paths), len(node_len)
    counts = np.zeros((n, m), dtype=np.int64)
    for i, (_, steps) in enumerate(paths):
        for s in steps:
            counts[i, seg_index[s]] += 1
    return [p[0] for p in paths], (counts > 0).astype(np.uint8), counts, np.asarray(node_len, dtype=np.int64)
