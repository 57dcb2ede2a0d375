"""Nucleotide diversity (pi) from an all-pairs similarity table -- device-backed drop-in for
the reference's scripts/pica2.py (same function names, arguments, return values, messages,
exit codes and stdout format; pica2.py:6-228).

What runs where: the table is parsed on the host (text), uploaded once, and everything
numeric -- greedy grouping (pica2.py:94-112), the weighted pair sum (:118-145) and
pi = n/(n-1) * sum(2 * term) (:154) with the optional / L (:163-164) -- runs in
libimpop_b200 (impop_greedy_groups + impop_reduce_identity).  No CPU fallback.

Determinism: the reference seeds each group with `set.pop()`, which depends on
PYTHONHASHSEED; here the seed is always the smallest remaining name.
The two agree whenever `similarity > threshold` is transitive on the table, and always when
threshold >= every identity.
"""
from __future__ import annotations

import argparse
import os
import sys

from .runtime import default_context
from .tables import SimilarityTable, TableFormatError, read_rows, read_table_fast


def read_similarity_file(filename):
    """(similarity table, set of element names, number of rows) -- pica2.py:6-58.

    The first item is a `SimilarityTable`: a Mapping keyed by (name_a, name_b) like the
    reference's dict, backed by the dense matrix the device reduces."""
    fast = read_table_fast(filename)        # machine-clean text: native reader; anything else: the csv path below
    if fast is not None:
        table, pair_count = fast
        if pair_count == 0:
            print(f"Warning: No similarity entries found in {filename}")
        return table, set(table.names), pair_count
    try:
        with open(filename, newline="") as handle:
            try:
                rows, pair_count, bad = read_rows(handle, on_bad_value="error")
            except TableFormatError as exc:
                if exc.args[0] == "empty":
                    print(f"Error: File {filename} is empty or missing a header")
                else:
                    print(f"Error: File must contain columns: {sorted(['group.a', 'group.b', 'estimated.identity'])}")
                    print(f"Found columns: {exc.args[0][1]}")
                sys.exit(1)
            if bad:
                line_no, text = bad[0]
                print(f"Error: Invalid similarity value on line {line_no}: {text}")
                sys.exit(1)
            if pair_count == 0:
                print(f"Warning: No similarity entries found in {filename}")
            table = SimilarityTable.from_rows(rows)
            return table, set(table.names), pair_count
    except FileNotFoundError:
        print(f"Error: File not found {filename}")
        sys.exit(1)
    except SystemExit:
        raise
    except Exception as exc:  # pica2.py:56-58
        print(f"Error reading file {filename}: {exc}")
        sys.exit(1)


def analyze_similarity_matrix(similarity_dict, elements, pair_count, threshold=1.0, sequence_length=None,
                              log_file=None, round_digits=None, ctx=None):
    """(pi, pi_per_site) as pica2.analyze_similarity_matrix returns them (pica2.py:60-169)."""
    ctx = ctx or default_context()
    table = SimilarityTable.from_mapping(similarity_dict, elements)

    def log_print(message):
        if log_file:
            print(message, file=log_file)

    n = len(table.names)
    log_print(f"Loaded {pair_count} pairwise similarities")
    log_print(f"Found {n} unique elements")
    if round_digits is not None:
        log_print(f"Rounded similarities to {round_digits} decimal places")
    if n == 0:
        log_print("Warning: No elements available to compute group pairs")
        return 0.0, 0.0                                                    # pica2.py:122-124

    ident = table.device(ctx, round_digits)
    group, weight = ctx.greedy_groups(ident, threshold)
    _, _, wsum = ctx.reduce_identity(ident, None, weight, length=sequence_length or 0)
    ctx.check()
    wsum_h = wsum.cpu().tolist()
    group_h = group.cpu().tolist()

    seeds = sorted(set(group_h))
    log_print(f"\nStep 1: Grouping elements (threshold > {threshold})")
    log_print(f"Found {len(seeds)} groups:")
    if log_file:
        members = {s: [] for s in seeds}
        for i, s in enumerate(group_h):
            members[s].append(table.names[i])
        for k, s in enumerate(seeds, 1):
            log_print(f"  G{k}: {members[s]} (size: {len(members[s])})")
    log_print("\nStep 2: Calculating group pairs")
    log_print(f"  {int(wsum_h[1])} group pairs with similarity data (per-pair terms are reduced on the device)")
    log_print("\nStep 3: Calculating pi")
    if wsum_h[1] == 0:
        log_print("Warning: No group pairs found with similarity data!")
        return 0.0, 0.0                                                    # pica2.py:150-152
    pi = wsum_h[2]
    log_print(f"  n (total elements) = {n}")
    log_print(f"  Number of group pairs with data = {int(wsum_h[1])}")
    log_print(f"  Sum of 2 * group_pairs = {2 * wsum_h[0]:.6f}")
    log_print(f"  pi = {n}/{n - 1} * {2 * wsum_h[0]:.6f} = {pi:.6f}")
    pi_per_site = None
    if sequence_length:
        pi_per_site = wsum_h[3]
        log_print("\nNormalization:")
        log_print(f"  Sequence length = {sequence_length}")
        log_print(f"  pi per site = {pi:.6f} / {sequence_length} = {pi_per_site:.8f}")
    return pi, pi_per_site


def main(argv=None):
    parser = argparse.ArgumentParser(
        description="Analyze similarity matrix with customizable threshold and sequence length normalization")
    parser.add_argument("input_file",
                        help="Input file with similarity data (TSV format with group.a, group.b, estimated.identity columns)")
    parser.add_argument("--threshold", "-t", type=float, default=0.99,
                        help="Similarity threshold for grouping elements (default: 0.99)")
    parser.add_argument("--sequence-length", "-l", type=int, help="Sequence length for normalizing pi per site")
    parser.add_argument("--log-dir", "-d", type=str, default=".", help="Directory to save log file (default: current directory)")
    parser.add_argument("--round-digits", "-r", type=int, default=None,
                        help="Round similarity values to specified decimal places (default: no rounding)")
    args = parser.parse_args(argv)

    base_name = os.path.splitext(os.path.basename(args.input_file))[0]
    log_filename = os.path.join(args.log_dir, f"{base_name}.log")
    os.makedirs(args.log_dir, exist_ok=True)
    similarity, elements, pair_count = read_similarity_file(args.input_file)
    with open(log_filename, "w") as log_file:
        log_file.write("Nucleotide Diversity Analysis Log\n=================================\n")
        log_file.write(f"Input file: {args.input_file}\nThreshold: {args.threshold}\n")
        if args.sequence_length:
            log_file.write(f"Sequence length: {args.sequence_length}\n")
        if args.round_digits is not None:
            log_file.write(f"Similarity rounding: {args.round_digits} decimal places\n")
        log_file.write(f"Log file: {log_filename}\n\n")
        pi, pi_per_site = analyze_similarity_matrix(similarity, elements, pair_count, threshold=args.threshold,
                                                    sequence_length=args.sequence_length, log_file=log_file,
                                                    round_digits=args.round_digits)
        log_file.write("\n" + "=" * 50 + "\nFINAL RESULTS:\n" + f"pi = {pi:.6f}\n")
        if pi_per_site is not None:
            log_file.write(f"pi per site = {pi_per_site:.8f}\n")
    if args.sequence_length:
        print(f"{pi_per_site:.8f} (sequence length: {args.sequence_length})")
    else:
        print(f"{pi:.6f} (sequence length: {args.sequence_length})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
