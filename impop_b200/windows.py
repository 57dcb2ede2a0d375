"""Windowed statistics straight from presence matrices (matrix mode), written with the column sets of the
reference's per-window wrappers so the R plotters keep working (SURVEY.md 8 b):

  pi     : REGION [SUBSET] LENGTH THRESHOLD R_VALUE PICA_OUTPUT       run_pica2_impg.sh:119-122, :184-188
  fst    : REGION LENGTH FST PI_A PI_B PI_XY DXY DA                   run_h-fst.sh:148, :91
  tajd   : REGION LENGTH SAMPLES SEGREGATING_SITES PI TAJIMAS_D       run_tajd.sh:101, :192-196 (NaN -> NA)

The numbers come from one `WindowBatch.stats()` call (fused similarity + reductions on the GPU); this module
only formats rows.  Multi-GPU: each rank formats its own shard of windows, or rank 0 formats the gathered rows.
"""
from __future__ import annotations

import math

from ._native import ST

HEADERS = {
    "pi": ["REGION", "LENGTH", "THRESHOLD", "R_VALUE", "PICA_OUTPUT"],
    "pi_subset": ["REGION", "SUBSET", "LENGTH", "THRESHOLD", "R_VALUE", "PICA_OUTPUT"],
    "fst": ["REGION", "LENGTH", "FST", "PI_A", "PI_B", "PI_XY", "DXY", "DA"],
    "tajd": ["REGION", "LENGTH", "SAMPLES", "SEGREGATING_SITES", "PI", "TAJIMAS_D"],
}


def region_name(chrom: str, start: int, end: int, prefix: str = "CHM13#0#") -> str:
    """`CHM13#0#chr2:109332703-109382703` -- the form plot_*_trend.R parses (plot_pi_trend.R:191)."""
    return f"{prefix}{chrom}:{start}-{end}"


def pi_rows(regions, lengths, stats, threshold="1.0", r_value="NA", subset=None):
    """PICA_OUTPUT is pica2's stdout line: `{pi_per_site:.8f} (sequence length: L)` (pica2.py:225-228)."""
    for reg, L, row in zip(regions, lengths, stats):
        out = (f"{row[ST['pi_per_site']]:.8f} (sequence length: {int(L)})" if L
               else f"{row[ST['pi']]:.6f} (sequence length: None)")
        yield [reg] + ([subset] if subset is not None else []) + [str(int(L)), str(threshold), str(r_value), out]


def fst_rows(regions, lengths, stats):
    """Six `%.8f` fields as h-fst.py:338-339 prints them."""
    keys = ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da")
    for reg, L, row in zip(regions, lengths, stats):
        yield [reg, str(int(L))] + [f"{row[ST[k]]:.8f}" for k in keys]


def tajd_rows(regions, lengths, stats, counts):
    """PI = per-site pi at 8 decimals (run_tajd.sh:166-174), D as tj_d.py prints a float, NaN -> NA (run_tajd.sh:192-194)."""
    for reg, L, row, cnt in zip(regions, lengths, stats, counts):
        d = float(row[ST["tajima_d"]])
        pi = row[ST["pi_per_site"]] if L else row[ST["pi"]]
        yield [reg, str(int(L)), str(int(cnt[0])), str(int(cnt[7])), f"{pi:.8f}", "NA" if math.isnan(d) else repr(d)]


def write_tsv(handle, kind: str, rows):
    handle.write("\t".join(HEADERS[kind]) + "\n")
    for row in rows:
        handle.write("\t".join(row) + "\n")
