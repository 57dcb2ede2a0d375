"""Haplotype-cluster frequencies -- device-backed drop-in for the reference's scripts/af.py
(af.py:7-89), plus the per-site allele-count mode BASELINE.json config 4 asks for.

Cluster mode (the reference's semantics): names are cut at the first ':' (af.py:13-14), samples
are linked when estimated.identity >= threshold (af.py:38) and the connected components are
ordered by (-size, sorted members) (af.py:43).  Linking + component labelling run in
libimpop_b200 (impop_cluster); ordering and text output are host work.

Site mode (`--sites`; no CLI for it exists in the reference, whose only per-site allele
counting is the unfinished scripts/wip/op-afs.py:26-45): a site-major bit matrix and
population masks give counts[site, pop] and count / |pop| (impop_site_counts).
No CPU fallback in either mode.
"""
from __future__ import annotations

import argparse
import csv
import sys

import numpy as np

from .runtime import default_context
from .tables import SimilarityTable


def load_pairs(path):
    """(rows [(a, b, identity)], sorted sample names) -- af.py:7-19 (KeyError / ValueError propagate as there)."""
    rows, samples = [], set()
    with open(path) as handle:
        for rec in csv.DictReader(handle, delimiter="\t"):
            a = rec["group.a"].split(":", 1)[0]
            b = rec["group.b"].split(":", 1)[0]
            rows.append((a, b, float(rec["estimated.identity"])))
            samples.update((a, b))
    return rows, sorted(samples)


def load_table_fast(path):
    """The table through libimpop_b200's native reader (tables.read_table_fast) when that is equivalent to load_pairs:
    machine-clean text, every row a distinct pair of two different samples (then "any row reaching the threshold links",
    af.py:37-39, and "the last row of a pair stays" coincide) and names that stay distinct once cut at the first ':'
    (af.py:13-14).  None otherwise -- the caller reads the rows as the reference does."""
    from .tables import read_table_fast
    got = read_table_fast(path)
    if got is None:
        return None
    table, data_rows = got
    n = len(table.names)
    if n == 0 or np.count_nonzero(~np.isnan(table.matrix[np.triu_indices(n, 1)])) != data_rows:
        return None                                             # repeated pairs or self pairs: rows are not one per pair
    if not np.isnan(np.diagonal(table.matrix)).all():
        return None
    short = [s.split(":", 1)[0] for s in table.names]
    if len(set(short)) != n:
        return None
    order = sorted(range(n), key=lambda i: short[i])           # cutting the coordinates can change the sort order
    if order != list(range(n)):
        table = SimilarityTable([short[i] for i in order], np.ascontiguousarray(table.matrix[np.ix_(order, order)]))
    else:
        table = SimilarityTable(short, table.matrix)
    return table


def cluster(rows, samples, threshold, ctx=None, table=None):
    """Connected components of {identity >= threshold}, as lists of names ordered like af.py:35-44."""
    ctx = ctx or default_context()
    if table is None:
        samples = list(samples)
        if not samples:
            return []
        table = SimilarityTable.from_rows(rows, combine="max")     # any row reaching the threshold links (af.py:37-39)
        extra = sorted(set(samples) - set(table.names))
        if extra:                                                   # samples that occur in no row stay singletons
            table = _with_names(table, extra)
    comp = ctx.cluster(table.device(ctx), threshold)
    ctx.check()
    comp = comp.cpu().numpy()
    groups = {}
    for i, c in enumerate(comp.tolist()):
        groups.setdefault(c, []).append(table.names[i])
    return sorted(groups.values(), key=lambda members: (-len(members), sorted(members)))


def _with_names(table, extra):
    names = sorted(table.names + list(extra))
    mat = np.full((len(names), len(names)), np.nan)
    where = [names.index(s) for s in table.names]
    mat[np.ix_(where, where)] = table.matrix
    return SimilarityTable(names, mat)


def build_summary(clusters):
    """[(cluster_id, count, frequency, sorted members)] -- af.py:46-54."""
    total = sum(len(c) for c in clusters)
    return [(f"c{k}", len(members), (len(members) / total) if total else 0.0, sorted(members))
            for k, members in enumerate(clusters, 1)]


def write_summary(summary, out_file):
    writer = csv.writer(out_file, delimiter="\t")
    writer.writerow(["cluster_id", "count", "frequency"])
    for cid, count, freq, _ in summary:
        writer.writerow([cid, count, f"{freq:.6f}"])


def write_details(summary, threshold, path):
    with open(path, "w", newline="") as handle:
        writer = csv.writer(handle, delimiter="\t")
        writer.writerow(["sample_id", "cluster_id", "threshold"])
        for cid, _, _, members in summary:
            for sample in members:
                writer.writerow([sample, cid, threshold])


# ----------------------------------------------------------------------------------------------
# per-site allele counts (BASELINE.json config 4)
# ----------------------------------------------------------------------------------------------
def site_allele_counts(site_bits: np.ndarray, pop_masks: np.ndarray, ctx=None, want_freq: bool = True):
    """counts[s, p] = popcount(site_bits[s] & pop_masks[p]) (int32), freq = count / |pop| (fp64).

    site_bits: (sites, words) uint64, bit h of a row = haplotype h carries the allele;
    pop_masks: (pops, words) uint64."""
    ctx = ctx or default_context()
    sites = ctx.upload(np.ascontiguousarray(site_bits, dtype=np.uint64).view(np.int64))
    masks = ctx.upload(np.ascontiguousarray(pop_masks, dtype=np.uint64).view(np.int64))
    counts, freq = ctx.site_counts(sites, masks, want_freq=want_freq)
    ctx.check()
    return counts.cpu().numpy(), (freq.cpu().numpy() if freq is not None else None)


def _run_sites(args):
    """--sites FILE.npz with arrays `sites` (M x words uint64) and `masks` (P x words uint64), optional `pops` names."""
    data = np.load(args.sites, allow_pickle=False)
    counts, freq = site_allele_counts(data["sites"], data["masks"])
    names = [str(s) for s in data["pops"]] if "pops" in data.files else [f"pop{p}" for p in range(counts.shape[1])]
    out = open(args.output, "w", newline="") if args.output else sys.stdout
    try:
        writer = csv.writer(out, delimiter="\t")
        writer.writerow(["site"] + [f"{p}.count" for p in names] + [f"{p}.freq" for p in names])
        for s in range(counts.shape[0]):
            writer.writerow([s] + counts[s].tolist() + [f"{v:.6f}" for v in freq[s].tolist()])
    finally:
        if args.output:
            out.close()
    return 0


def main(argv=None):
    parser = argparse.ArgumentParser(description="Cluster samples in loc.sim-style tables by identity threshold.")
    parser.add_argument("--input", default="loc.sim", help="Path to the similarity table (default: loc.sim)")
    parser.add_argument("--threshold", type=float, default=1.0,
                        help="Minimum estimated.identity to link samples (default: 1.0)")
    parser.add_argument("--output", help="Optional output TSV path for cluster summary; stdout if omitted")
    parser.add_argument("--details", help="Optional path to write detailed sample assignments")
    parser.add_argument("--sites", help="(extension) .npz site-major bit matrix: per-site allele counts / frequencies instead")
    args = parser.parse_args(argv)
    if args.sites:
        return _run_sites(args)
    table = load_table_fast(args.input)
    if table is not None:
        clusters = cluster(None, None, args.threshold, table=table)
    else:
        rows, samples = load_pairs(args.input)
        clusters = cluster(rows, samples, args.threshold)
    summary = build_summary(clusters)
    if args.output:
        with open(args.output, "w", newline="") as handle:
            write_summary(summary, handle)
    else:
        write_summary(summary, sys.stdout)
    if args.details:
        write_details(summary, args.threshold, args.details)
    return 0


if __name__ == "__main__":
    sys.exit(main())
