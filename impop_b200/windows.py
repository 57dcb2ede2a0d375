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
    "pooled_fst": ["REGION", "LENGTH", "THRESHOLD", "R_VALUE", "PI_A", "PI_B", "PI_C", "PI_AB_AVG", "FST"],
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


def pooled_fst_text(pi_a: str, pi_b: str, pi_c: str):
    """(PI_AB_AVG, FST) from the three 8-decimal per-site pi texts pica2 printed, exactly as the two inline python
    snippets of run_fst_impg.sh:199-218 form them: `NA` when pi_C == 0."""
    a, b, c = float(pi_a), float(pi_b), float(pi_c)
    avg = 0.5 * (a + b)
    return f"{avg:.8f}", ("NA" if c == 0 else f"{(c - avg) / c:.8f}")


def pooled_fst_rows(regions, lengths, stats_a, stats_b, stats_c, threshold="1.0", r_value="NA"):
    """run_fst_impg.sh:158, :220: pica2's per-site pi over subset A, subset B and their union C (three passes whose SUBSET
    class is A, B, A + B), then the pooled estimator on the printed values."""
    for reg, L, ra, rb, rc in zip(regions, lengths, stats_a, stats_b, stats_c):
        key = ST["pi_per_site"] if L else ST["pi"]
        pa, pb, pc = (f"{row[key]:.8f}" if L else f"{row[key]:.6f}" for row in (ra, rb, rc))
        avg, fst = pooled_fst_text(pa, pb, pc)
        yield [reg, str(int(L)), str(threshold), str(r_value), pa, pb, pc, avg, fst]


def write_tsv(handle, kind: str, rows):
    handle.write("\t".join(HEADERS[kind]) + "\n")
    for row in rows:
        handle.write("\t".join(row) + "\n")


# ------------------------------------------------------------------------------------------------------------
# Matrix-mode driver: window graphs in, the wrappers' TSVs out (replaces the per-window loops of
# run_pica2_impg.sh:126-192, run_h-fst.sh:155-194 and run_tajd.sh:103-198 -- one batch on the GPU instead of
# 3-6 process spawns per BED row)
# ------------------------------------------------------------------------------------------------------------
def batch_from_graphs(ctx, graphs, pop_a_ids=None, pop_b_ids=None, subset_ids=None):
    """GraphWindow list (impop_b200.ingest) -> WindowBatch with per-window labels.  Population / subset identifiers are
    assembly names or PanSN prefixes, expanded per window against its row names exactly as h-fst.py:64-82 does."""
    from . import ingest
    from .engine import WindowBatch
    from .hfst import expand_population
    wins = []
    for g in graphs:
        pa = expand_population(pop_a_ids, g.names)[0] if pop_a_ids else None
        pb = expand_population(pop_b_ids, g.names)[0] if pop_b_ids else None
        sub = expand_population(subset_ids, g.names)[0] if subset_ids else None
        lab = ingest.labels_from_names(g.names, pa, pb, sub, sub)
        wins.append((g.x_bits, g.node_len, lab, g.length))
    return WindowBatch.from_windows(ctx, wins)


def main(argv=None):
    """impop-windows: windowed pi / Hudson Fst / Tajima's D straight from window graphs.

        impop-windows.py --gfa-list windows.tsv [-a popA.txt -b popB.txt] [-s subset.txt]
                         [--pi-out pi.tsv] [--fst-out fst.tsv] [--tajd-out tajd.tsv] [--save-batch windows.npz]
        impop-windows.py --batch windows.npz ...

    windows.tsv: one line per window, `REGION<TAB>path/to/window.gfa` with REGION like CHM13#0#chr2:100000-150000
    (its end - start is the window length L the wrappers pass as -l)."""
    import argparse
    import re
    import sys

    from . import ingest
    from .engine import Context
    from .hfst import read_subset_file

    ap = argparse.ArgumentParser(prog="impop-windows", description=main.__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--gfa-list", help="TSV: REGION <tab> GFA file of that window")
    src.add_argument("--batch", help="binary container written by --save-batch")
    ap.add_argument("-a", "--pop-a", help="file listing population A (h-fst.py -a)")
    ap.add_argument("-b", "--pop-b", help="file listing population B (h-fst.py -b)")
    ap.add_argument("-s", "--subset", help="file listing the samples pi / S / Tajima's D are computed over (run_tajd.sh -l)")
    ap.add_argument("--pi-out"), ap.add_argument("--fst-out"), ap.add_argument("--tajd-out")
    ap.add_argument("--pooled-fst-out", help="run_fst_impg.sh's table: pica2 pi of subset A, B and their union, pooled Fst (needs -a, -b)")
    ap.add_argument("--save-batch", help="also write the parsed windows as one binary container")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    if args.batch:
        graphs = ingest.load_batch(args.batch)
    else:
        todo = []
        with open(args.gfa_list) as fh:
            for line in fh:
                if not line.strip() or line.startswith("#"):
                    continue
                region, path = line.rstrip("\n").split("\t")[:2]
                mt = re.search(r":(\d+)-(\d+)$", region)
                todo.append((region, path, int(mt.group(2)) - int(mt.group(1)) if mt else 0))
        graphs = ingest.read_gfa_many(todo)               # parsed on all host cores (the reader runs outside the GIL)
    if args.save_batch:
        ingest.save_batch(args.save_batch, graphs)
    if (args.fst_out or args.pooled_fst_out) and (args.pop_a is None or args.pop_b is None):
        ap.error("--fst-out / --pooled-fst-out need both -a and -b")
    pop_a = read_subset_file(args.pop_a) if args.pop_a else None
    pop_b = read_subset_file(args.pop_b) if args.pop_b else None
    subset = read_subset_file(args.subset) if args.subset else None
    ctx = Context(args.device)
    batch = batch_from_graphs(ctx, graphs, pop_a, pop_b, subset)
    stats, counts = batch.stats()
    ctx.check()
    stats, counts = stats.cpu().numpy(), counts.cpu().numpy()
    batch.close()
    regions = [g.region or f"window{i}" for i, g in enumerate(graphs)]
    lengths = [g.length for g in graphs]
    if args.pooled_fst_out:
        per_subset = []
        for ids in (pop_a, pop_b, set(pop_a) | set(pop_b)):           # run_fst_impg.sh:143-147: C = union list
            b = batch_from_graphs(ctx, graphs, subset_ids=ids)
            per_subset.append(b.stats()[0].cpu().numpy())
            ctx.check()
            b.close()
        with (sys.stdout if args.pooled_fst_out == "-" else open(args.pooled_fst_out, "w")) as out:
            write_tsv(out, "pooled_fst", pooled_fst_rows(regions, lengths, *per_subset))
    for path, kind, rows in ((args.pi_out, "pi", pi_rows(regions, lengths, stats)),
                             (args.fst_out, "fst", fst_rows(regions, lengths, stats)),
                             (args.tajd_out, "tajd", tajd_rows(regions, lengths, stats, counts))):
        if path:
            with (sys.stdout if path == "-" else open(path, "w")) as out:
                write_tsv(out, kind, rows)
    return 0
