"""Drop-in for the reference's scripts/hudson/hud.py: Hudson Fst by the `direct` or the `grouped` method.

`direct` is h-fst.py's estimator (hud.py:130-170 == h-fst.py:130-171) and runs through `hfst.calculate_fst`.
`grouped` (hud.py:64-128, 235-263) groups near-identical sequences inside each population and weighs the groups'
representatives by group frequency.  The reference seeds its groups with `set.pop()`, whose order depends on
PYTHONHASHSEED, so its output is only defined when `similarity > threshold` is an equivalence relation inside each
population; here the seed is always the smallest remaining name (as in the pica2 drop-in), which agrees with the
reference wherever the reference agrees with itself.  On a table with absent pairs the reference takes the first
member pair it finds between two groups (hud.py:86-97); here it is the representatives' pair or nothing
(odgi / impg tables are complete).  Grouping, the weighted sums and the pair counts run on the GPU
(`impop_greedy_groups`, `impop_reduce_identity`); sub-matrices of the (device-rounded) table are selected on the host.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import hfst
from .hfst import read_similarity_file, read_subset_file  # noqa: F401  (same readers: hud.py:15-62)
from .runtime import default_context
from .tables import SimilarityTable

calculate_diversity_direct = hfst.calculate_diversity      # hud.py:130-170


def _population(table: SimilarityTable, ctx, ident, members):
    """(row indices sorted by name, group id per row, weight per row, grouped diversity, groups, group pairs with data)."""
    idx = sorted(table.index[s] for s in members if s in table.index)
    n = len(idx)
    sel = np.asarray(idx, dtype=np.int64)
    sub = ctx.upload(np.ascontiguousarray(ident[np.ix_(sel, sel)])) if n else None
    return idx, sel, sub, n


def group_sequences(similarities, sequences, threshold=0.999, round_digits=None, ctx=None):
    """Sorted list of sorted groups (hud.py:64-84), seed = smallest remaining name."""
    ctx = ctx or default_context()
    table = SimilarityTable.from_mapping(similarities, sequences)
    ident = table.host(ctx, round_digits)
    idx, sel, sub, n = _population(table, ctx, ident, sequences)
    if n == 0:
        return []
    group, _ = ctx.greedy_groups(sub, threshold)
    ctx.check()
    members = {}
    for row, seed in enumerate(group.cpu().tolist()):
        members.setdefault(seed, []).append(table.names[idx[row]])
    return sorted(sorted(g) for g in members.values())


def calculate_diversity_grouped(similarities, sequences, threshold=0.999, round_digits=None, ctx=None):
    """(diversity, number of groups, group pairs without data) -- hud.py:99-128."""
    ctx = ctx or default_context()
    table = SimilarityTable.from_mapping(similarities, sequences)
    ident = table.host(ctx, round_digits)
    div, groups, missing, _ = _grouped(table, ctx, ident, sequences, threshold)
    return div, groups, missing


def _grouped(table, ctx, ident, members, threshold):
    idx, sel, sub, n = _population(table, ctx, ident, members)
    if n == 0:
        return 0.0, 0, 0, (idx, None)
    group, weight = ctx.greedy_groups(sub, threshold)
    _, _, wsum = ctx.reduce_identity(sub, None, weight)
    ctx.check()
    weight = weight.cpu().numpy()
    g = int((weight > 0).sum())
    if n <= 1:
        return 0.0, g, 0, (idx, weight)                                       # hud.py:104-105
    ws = wsum.cpu().tolist()
    return ws[2], g, g * (g - 1) // 2 - int(ws[1]), (idx, weight)             # n/(n-1) * sum 2 f_i f_j (1 - s): hud.py:119, :125


def calculate_fst(similarities, pop_a, pop_b, sequence_length=None, round_digits=None, log_file=None, method="direct",
                  threshold=0.999, ctx=None):
    """dict(fst, pi_a, pi_b, pi_xy, dxy, da) -- hud.py:172-300."""
    if method != "grouped":
        return hfst.calculate_fst(similarities, pop_a, pop_b, sequence_length=sequence_length, round_digits=round_digits,
                                  log_file=log_file, ctx=ctx)
    ctx = ctx or default_context()
    table = SimilarityTable.from_mapping(similarities)

    def log_print(msg):
        if log_file:
            print(msg, file=log_file)

    pop_a, pop_b = set(pop_a), set(pop_b)
    overlap = pop_a & pop_b
    if overlap:
        print(f"Warning: {len(overlap)} sequences appear in both populations", file=sys.stderr)
        pop_a, pop_b = pop_a - overlap, pop_b - overlap
    log_print("FST Calculation")
    log_print("=" * 50)
    log_print(f"Population A: {len(pop_a)} sequences")
    log_print(f"Population B: {len(pop_b)} sequences")
    log_print(f"Method: {method}")
    log_print(f"Grouping threshold: {threshold}")
    if round_digits is not None:
        log_print(f"Rounding similarities to {round_digits} decimal places")
    log_print("")
    ident = table.host(ctx, round_digits)
    pi_a, groups_a, miss_a, (ia, wa) = _grouped(table, ctx, ident, pop_a, threshold)
    pi_b, groups_b, miss_b, (ib, wb) = _grouped(table, ctx, ident, pop_b, threshold)
    log_print("Within-population diversity (π) using grouped method:")
    log_print(f"  πA = {pi_a:.6f} ({groups_a} groups from {len(pop_a)} sequences, {miss_a} missing pairs)")
    log_print(f"  πB = {pi_b:.6f} ({groups_b} groups from {len(pop_b)} sequences, {miss_b} missing pairs)")
    pi_xy = 0.5 * (pi_a + pi_b)                                                # hud.py:230
    log_print(f"  πXY = {pi_xy:.6f} (average of πA and πB)")
    log_print("")
    log_print("Between-population diversity (Dxy):")
    dxy, missing = 0.0, groups_a * groups_b
    if wa is not None and wb is not None:
        # representatives of both populations, weights |G_a| / n_A and |G_b| / n_B, one cross-population reduction
        rows = np.asarray(ia + ib, dtype=np.int64)
        sub = ctx.upload(np.ascontiguousarray(ident[np.ix_(rows, rows)]))
        lab = ctx.upload(np.concatenate([np.full(len(ia), 2, dtype=np.uint8), np.full(len(ib), 4, dtype=np.uint8)]))
        _, _, wsum = ctx.reduce_identity(sub, lab, ctx.upload(np.concatenate([wa, wb])))
        ctx.check()
        ws = wsum.cpu().tolist()
        dxy, missing = ws[0], groups_a * groups_b - int(ws[1])                 # hud.py:246-258
    log_print(f"  Dxy = {dxy:.6f} (from {groups_a} x {groups_b} group pairs, {missing} missing)")
    log_print("")
    if dxy > 0:
        fst = (dxy - pi_xy) / dxy                                              # hud.py:268
        log_print("FST calculation:")
        log_print("  FST = (Dxy - πXY) / Dxy")
        log_print(f"      = ({dxy:.6f} - {pi_xy:.6f}) / {dxy:.6f}")
        log_print(f"      = {fst:.6f}")
    else:
        fst = 0.0
        log_print("FST = 0 (Dxy = 0)")
    if sequence_length and sequence_length > 0:
        L = sequence_length
        log_print("")
        log_print(f"Per-site values (sequence length = {L:,}):")
        log_print(f"  πA per site = {pi_a / L:.8f}")
        log_print(f"  πB per site = {pi_b / L:.8f}")
        log_print(f"  πXY per site = {pi_xy / L:.8f}")
        log_print(f"  Dxy per site = {dxy / L:.8f}")
        return {"fst": fst, "pi_a": pi_a / L, "pi_b": pi_b / L, "pi_xy": pi_xy / L, "dxy": dxy / L, "da": (dxy - pi_xy) / L}
    return {"fst": fst, "pi_a": pi_a, "pi_b": pi_b, "pi_xy": pi_xy, "dxy": dxy, "da": dxy - pi_xy}


def main(argv=None):
    parser = argparse.ArgumentParser(description="Calculate FST from pairwise sequence similarities",
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("similarity_file", help="TSV file with columns: group.a, group.b, estimated.identity")
    parser.add_argument("-a", "--pop-a", required=True, help="File listing sequence IDs for population A")
    parser.add_argument("-b", "--pop-b", required=True, help="File listing sequence IDs for population B")
    parser.add_argument("-l", "--length", type=int, default=None, help="Sequence length for per-site calculations")
    parser.add_argument("-r", "--round", type=int, default=None, help="Round similarities to N decimal places")
    parser.add_argument("-m", "--method", choices=["direct", "grouped"], default="direct",
                        help="Calculation method: direct or grouped (default: direct)")
    parser.add_argument("-t", "--threshold", type=float, default=0.999,
                        help="Similarity threshold for grouping (default: 0.999, used only with -m grouped)")
    parser.add_argument("-d", "--log-dir", default=".", help="Directory for log file (default: current directory)")
    parser.add_argument("-v", "--verbose", action="store_true", help="Print detailed progress to stderr")
    args = parser.parse_args(argv)
    if args.verbose:
        print(f"Reading similarity file: {args.similarity_file}", file=sys.stderr)
    similarities, all_sequences = read_similarity_file(args.similarity_file)
    if args.verbose:
        print("Reading population files...", file=sys.stderr)
    pop_a, pop_b = read_subset_file(args.pop_a), read_subset_file(args.pop_b)      # raw IDs: hud.py does not expand (hud.py:365-366)
    if args.verbose:
        print(f"Population A: {len(pop_a)} sequences", file=sys.stderr)
        print(f"Population B: {len(pop_b)} sequences", file=sys.stderr)
        print(f"Method: {args.method}", file=sys.stderr)
        if args.method == "grouped":
            print(f"Grouping threshold: {args.threshold}", file=sys.stderr)
    missing_a, missing_b = pop_a - all_sequences, pop_b - all_sequences
    if missing_a:
        print(f"Warning: {len(missing_a)} sequences from population A not found in similarity file", file=sys.stderr)
    if missing_b:
        print(f"Warning: {len(missing_b)} sequences from population B not found in similarity file", file=sys.stderr)
    pop_a, pop_b = pop_a & all_sequences, pop_b & all_sequences
    if not pop_a or not pop_b:
        print("Error: No valid sequences found in one or both populations", file=sys.stderr)
        return 1
    base_name = os.path.splitext(os.path.basename(args.similarity_file))[0]
    log_path = os.path.join(args.log_dir, f"{base_name}_fst.log")
    os.makedirs(args.log_dir, exist_ok=True)
    with open(log_path, "w") as log_file:
        results = calculate_fst(similarities, pop_a, pop_b, sequence_length=args.length, round_digits=args.round,
                                log_file=log_file, method=args.method, threshold=args.threshold)
    print(f"{results['fst']:.8f}\t{results['pi_a']:.8f}\t{results['pi_b']:.8f}\t"
          f"{results['pi_xy']:.8f}\t{results['dxy']:.8f}\t{results['da']:.8f}")
    if args.verbose:
        print(f"Detailed log saved to: {log_path}", file=sys.stderr)
    return 0
