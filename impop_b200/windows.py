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


def tajd_rows(regions, lengths, stats, counts, d_from_text=None, samples=None):
    """PI = per-site pi at 8 decimals (run_tajd.sh:166-174), D as tj_d.py prints a float, NaN -> NA (run_tajd.sh:192-194).
    `d_from_text`: D recomputed from the PRINTED pi (what run_tajd.sh:180 hands tj_d.py is pica2's 8-decimal text), so that
    the D of a row follows from the PI of the same row; default: the device's D from the unrounded pi.  `samples`: SAMPLES
    column / n of tj_d.py when it is the size of the sample list (run_tajd.sh:83) rather than the rows found in the window."""
    for k, (reg, L, row, cnt) in enumerate(zip(regions, lengths, stats, counts)):
        d = float(row[ST["tajima_d"]]) if d_from_text is None else float(d_from_text[k])
        pi = row[ST["pi_per_site"]] if L else row[ST["pi"]]
        yield [reg, str(int(L)), str(int(cnt[0]) if samples is None else int(samples)), str(int(cnt[7])), f"{pi:.8f}",
               "NA" if math.isnan(d) else repr(d)]


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
        wins.append((g.x_bits, g.node_len, lab, g.length, g.row_adj, g.win_const, g.col_mult))     # affine form when compacted
    runs = [getattr(g, "site_runs", -1) for g in graphs]      # counted before compaction (the node order is gone afterwards)
    return WindowBatch.from_windows(ctx, wins, site_runs=runs if any(r >= 0 for r in runs) else None)


def batch_from_flat(ctx, flat, pop_a_ids=None, pop_b_ids=None, subset_ids=None):
    """ingest.FlatBatch (memory-mapped container) -> WindowBatch: one upload of the presence words, node lengths and labels;
    the classes are decided once per unique haplotype (h-fst.py:18-61 prefixes) and gathered per row -- no per-window loop."""
    import numpy as np

    from .engine import WindowBatch
    from .hfst import canonicalize_identifier

    def prefixes(ids):
        return None if ids is None else [p for p in (canonicalize_identifier(i) for i in ids) if p]
    lab = flat.labels(prefixes(pop_a_ids), prefixes(pop_b_ids), prefixes(subset_ids))
    both = (lab & 6) == 6                                  # h-fst.py:181-185: listed in both populations -> in neither
    lab = np.where(both, lab & ~np.uint8(6), lab).astype(np.uint8)
    x = ctx.upload(np.ascontiguousarray(flat.x).view(np.int32))
    nl = ctx.upload(np.ascontiguousarray(flat.node_len).view(np.int32))
    extra = {}
    if flat.row_adj is not None:                           # affine form + the variant-site counts taken before compaction
        runs = np.ascontiguousarray(flat.site_runs)
        extra = dict(row_adj=ctx.upload(np.ascontiguousarray(flat.row_adj)), win_const=np.ascontiguousarray(flat.win_const),
                     col_mult=ctx.upload(np.ascontiguousarray(flat.col_mult)), site_runs=runs if (runs >= 0).any() else None)
    q = (np.asarray(flat.node_len).astype(np.int64) // 255 + 254) // 255        # heavy-table entries per node -> per window
    csum = np.concatenate([[0], np.cumsum(q)])
    lo = np.asarray(flat.len_off, dtype=np.int64)
    heavy = (csum[lo + np.asarray(flat.m, dtype=np.int64)] - csum[lo]).astype(np.int32)
    return WindowBatch(ctx, flat.n, flat.m, flat.pitch, flat.x_off, flat.len_off, flat.row_off, flat.length, x, nl,
                       ctx.upload(lab), heavy_entries=heavy, **extra)


def stats_disjoint_absent(ctx, batch, labels_host, lab_off):
    """Per-window statistics with the convention of a similarity tool that prints NO row for a pair of paths that share no
    node [UPSTREAM-UNVERIFIED: odgi similarity]: such a pair is absent -- not counted in any denominator -- which is how
    pica2.py:132-134 and h-fst.py:147-153 treat a missing row.  (The fused path counts it with pi_ij = 1.)  Materialises
    every window's table on the device and reduces it with NaN in place of those pairs."""
    import numpy as np
    stats = np.zeros((batch.windows, 20), dtype=np.float64)
    counts = np.zeros((batch.windows, 8), dtype=np.int64)
    fused_s, fused_c = batch.stats()
    ctx.check()
    fused_s, fused_c = fused_s.cpu().numpy(), fused_c.cpu().numpy()
    for w in range(batch.windows):
        n = int(batch.n[w])
        I, _, pi = batch.pairwise(w)
        ctx.check()
        ident = 1.0 - pi.cpu().numpy()
        disjoint = (I.cpu().numpy() == 0) & ~np.eye(n, dtype=bool)
        ident[disjoint] = np.nan
        lab = ctx.upload(np.ascontiguousarray(labels_host[int(lab_off[w]):int(lab_off[w]) + n]))
        st, ct, _ = ctx.reduce_identity(ctx.upload(ident), lab, None, length=int(batch.length[w]), seg_sites=float(fused_c[w][7]))
        ctx.check()
        stats[w], counts[w] = st.cpu().numpy(), ct.cpu().numpy()
        counts[w][7] = fused_c[w][7]
        stats[w][19] = fused_s[w][19]
    return stats, counts


def main(argv=None):
    """impop-windows: windowed pi / Hudson Fst / Tajima's D straight from window graphs.

        impop-windows.py --gfa-list windows.tsv [-a popA.txt -b popB.txt] [-s subset.txt]
                         [--pi-out pi.tsv] [--fst-out fst.tsv] [--tajd-out tajd.tsv] [--save-batch windows.impw]
        impop-windows.py --batch windows.impw ...

    windows.tsv: one line per window, `REGION<TAB>path/to/window.gfa` with REGION like CHM13#0#chr2:100000-150000
    (its end - start is the window length L the wrappers pass as -l)."""
    import argparse
    import re
    import sys
    import time

    import numpy as np

    from . import ingest
    from .engine import Context
    from .hfst import read_subset_file

    ap = argparse.ArgumentParser(prog="impop-windows", description=main.__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--gfa-list", help="TSV: REGION <tab> GFA file of that window")
    src.add_argument("--batch", help="binary container written by --save-batch (flat .impw, or .npz)")
    ap.add_argument("-a", "--pop-a", help="file listing population A (h-fst.py -a)")
    ap.add_argument("-b", "--pop-b", help="file listing population B (h-fst.py -b)")
    ap.add_argument("-s", "--subset", help="file listing the samples pi / S / Tajima's D are computed over (run_tajd.sh -l)")
    ap.add_argument("--pi-out"), ap.add_argument("--fst-out"), ap.add_argument("--tajd-out")
    ap.add_argument("--pooled-fst-out", help="run_fst_impg.sh's table: pica2 pi of subset A, B and their union, pooled Fst (needs -a, -b)")
    ap.add_argument("--save-batch", help="also write the parsed (ingested) windows as one binary container (.npz: numpy archive, else flat)")
    ap.add_argument("--presence-only", action="store_true",
                    help="a path that visits a node several times counts it once (default: multiset coverage, min(count_a, count_b) per node, as the similarity tools accumulate steps)")
    ap.add_argument("--no-compact", action="store_true", help="keep every node as a matrix column (default: constant columns merged, empty ones dropped at ingest)")
    ap.add_argument("--disjoint-absent", action="store_true",
                    help="a pair of paths sharing no node is treated as absent from the table (not counted), as the scripts treat a row the similarity tool did not print")
    ap.add_argument("--tajd-sites", choices=["nodes", "bubbles"], default="nodes",
                    help="S of Tajima's D: segregating nodes (default) or variant sites counted like a bubble caller's VCF records "
                         "(runs of segregating nodes between nodes every haplotype carries; run_tajd.sh:126-148 counts povu's records)")
    ap.add_argument("--tajd-samples-from-list", action="store_true", help="SAMPLES / n of Tajima's D = size of the -s list (run_tajd.sh:83) instead of the rows found per window")
    ap.add_argument("--timings", action="store_true", help="print stage times (parse / ingest / upload / kernels / format) on stderr")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    if (args.fst_out or args.pooled_fst_out) and (args.pop_a is None or args.pop_b is None):
        ap.error("--fst-out / --pooled-fst-out need both -a and -b")
    t = {"start": time.perf_counter()}

    def mark(name):
        t[name] = time.perf_counter()

    flat = None
    if args.batch:
        with open(args.batch, "rb") as fh:
            magic = fh.read(8)
        if len(magic) == 8 and int.from_bytes(magic, "little") == ingest.FLAT_MAGIC:
            flat = ingest.load_flat(args.batch)
            graphs = None
        else:
            graphs = ingest.load_batch(args.batch)
        mark("load")
    else:
        todo = []
        with open(args.gfa_list) as fh:
            for line in fh:
                if not line.strip() or line.startswith("#"):
                    continue
                region, path = line.rstrip("\n").split("\t")[:2]
                mt = re.search(r":(\d+)-(\d+)$", region)
                todo.append((region, path, int(mt.group(2)) - int(mt.group(1)) if mt else 0))
        # parsed on all host cores (the reader runs outside the GIL), visit counts kept: a path that revisits a node
        # (duplication, inversion, loop) contributes min(count_a, count_b) * len to an intersection
        graphs = ingest.read_gfa_many(todo, want_counts=False if args.presence_only else "auto")   # counts only where a path revisits a node
        mark("parse")
        if not args.presence_only:
            revisits = sum(1 for g in graphs if g.counts is not None and g.counts.size and int(g.counts.max()) > 1)
            if revisits:
                print(f"impop-windows: {revisits} window(s) have paths that visit a node more than once: multiset coverage "
                      f"(copy-node expansion); --presence-only counts a node once", file=sys.stderr)
                graphs = [ingest.multiset_expand(g) if (g.counts is not None and g.counts.size and int(g.counts.max()) > 1) else g
                          for g in graphs]
        if not args.no_compact:
            graphs = ingest.compact_windows(graphs)           # one native call, all host threads
        mark("ingest")
    if args.save_batch:
        src_graphs = graphs if graphs is not None else [flat.window(w) for w in range(flat.windows)]
        (ingest.save_batch if args.save_batch.endswith(".npz") else ingest.save_flat)(args.save_batch, src_graphs)
    pop_a = read_subset_file(args.pop_a) if args.pop_a else None
    pop_b = read_subset_file(args.pop_b) if args.pop_b else None
    subset = read_subset_file(args.subset) if args.subset else None
    ctx = Context(args.device, lite=True)           # plain device buffers over the C ABI: no torch import on this command line
    mark("context")
    make = (lambda **kw: batch_from_flat(ctx, flat, **kw)) if flat is not None else (lambda **kw: batch_from_graphs(ctx, graphs, **kw))
    batch = make(pop_a_ids=pop_a, pop_b_ids=pop_b, subset_ids=subset)
    mark("upload")
    if args.disjoint_absent:
        stats, counts = stats_disjoint_absent(ctx, batch, batch.labels.cpu().numpy(), batch.lab_off)
    else:
        stats, counts = batch.stats()
        ctx.check()
        stats, counts = stats.cpu().numpy(), counts.cpu().numpy()
    batch.close()
    mark("kernels")
    if flat is not None:
        regions = [r or f"window{i}" for i, r in enumerate(flat.regions)]
        lengths = [int(v) for v in flat.length]
    else:
        regions = [g.region or f"window{i}" for i, g in enumerate(graphs)]
        lengths = [g.length for g in graphs]
    if args.pooled_fst_out:
        per_subset = []
        for ids in (pop_a, pop_b, set(pop_a) | set(pop_b)):           # run_fst_impg.sh:143-147: C = union list
            b = make(subset_ids=ids)
            per_subset.append(b.stats()[0].cpu().numpy())
            ctx.check()
            b.close()
        with (sys.stdout if args.pooled_fst_out == "-" else open(args.pooled_fst_out, "w")) as out:
            write_tsv(out, "pooled_fst", pooled_fst_rows(regions, lengths, *per_subset))
    d_text = None
    if args.tajd_out:
        # run_tajd.sh:166-180: tj_d.py receives the pi pica2 PRINTED (8 decimals) and the sample count
        from .tj_d import tajimas_d_batch
        if args.tajd_sites == "bubbles":
            counts = counts.copy()
            counts[:, 7] = np.asarray(stats)[:, ST["S_bubbles"]].astype(np.int64)
        nS = [int(len(subset)) if (args.tajd_samples_from_list and subset is not None) else int(c[0]) for c in counts]
        pis = [float(f"{(row[ST['pi_per_site']] if L else row[ST['pi']]):.8f}") for row, L in zip(stats, lengths)]
        ok = [k for k in range(len(nS)) if nS[k] >= 2]
        d_text = [float("nan")] * len(nS)
        if ok:
            vals = tajimas_d_batch([nS[k] for k in ok], [float(counts[k][7]) for k in ok], [pis[k] for k in ok], ctx=ctx)
            for k, v in zip(ok, vals):
                d_text[k] = v
    samples = len(subset) if (args.tajd_samples_from_list and subset is not None) else None
    for path, kind, rows in ((args.pi_out, "pi", pi_rows(regions, lengths, stats)),
                             (args.fst_out, "fst", fst_rows(regions, lengths, stats)),
                             (args.tajd_out, "tajd", tajd_rows(regions, lengths, stats, counts, d_text, samples))):
        if path:
            with (sys.stdout if path == "-" else open(path, "w")) as out:
                write_tsv(out, kind, rows)
    mark("format")
    if args.timings:
        keys = list(t)
        print("impop-windows timings (s): " + ", ".join(f"{b} {t[b] - t[a]:.3f}" for a, b in zip(keys, keys[1:]))
              + f", total {t[keys[-1]] - t['start']:.3f}; windows {len(regions)}", file=sys.stderr)
    return 0
