"""Hudson Fst from an all-pairs similarity table -- device-backed drop-in for the reference's
scripts/h-fst.py (h-fst.py:18-342: same function names, arguments, return values, messages on
stderr, exit codes, log file and stdout format).

Name handling (population lists -> PanSN prefixes -> sequence names) is host string work; the
means of 1 - identity within and between populations and the Fst arithmetic run in
libimpop_b200 (impop_reduce_identity).  No CPU fallback.
"""
from __future__ import annotations

import argparse
import os
import sys

from .runtime import default_context
from .tables import SimilarityTable, TableFormatError, read_rows, read_table_fast

_HAP_TAGS = (("_hap1", "#1#"), ("_hap2", "#2#"), ("_mat", "#1#"), ("_pat", "#2#"))


def canonicalize_identifier(identifier: str) -> str:
    """Assembly name -> PanSN prefix for str.startswith matching (h-fst.py:18-61).

    `HG00097_hap1_hprc_r2_v1.0.1` -> `HG00097#1#`, `HG01891_mat_...` -> `HG01891#1#`,
    `HG00097` -> `HG00097#`, `HG00097#1` -> `HG00097#1#`; blank or `#...` -> ''."""
    if not identifier:
        return ""
    token = identifier.strip()
    if not token or token[0] == "#":
        return ""
    cut = token.find("_hprc")
    if cut != -1:
        token = token[:cut]
    for suffix, tag in _HAP_TAGS:
        if token.endswith(suffix):
            return token[: -len(suffix)] + tag
    if "#" in token and token.endswith("#"):
        return token
    return token + "#"


def expand_population(raw_ids, all_sequences):
    """(set of sequence names matched by the identifiers' prefixes, identifiers that matched nothing) -- h-fst.py:64-82."""
    names = sorted(all_sequences)
    expanded, missing = set(), []
    for raw_id in raw_ids:
        prefix = canonicalize_identifier(raw_id)
        if not prefix:
            continue
        matches = [s for s in names if s.startswith(prefix)]
        if matches:
            expanded.update(matches)
        else:
            missing.append(raw_id)
    return expanded, missing


def read_similarity_file(filename):
    """(similarity table, set of sequence names) -- h-fst.py:84-119; an unparsable value is skipped with a warning."""
    fast = read_table_fast(filename)        # machine-clean text: native reader; anything else: the csv path below
    if fast is not None:
        return fast[0], set(fast[0].names)
    try:
        with open(filename, newline="") as handle:
            try:
                rows, _, bad = read_rows(handle, on_bad_value="skip")
            except TableFormatError as exc:
                if exc.args[0] == "empty":
                    print(f"Error: Empty file {filename}", file=sys.stderr)
                else:
                    print(f"Error: File must contain columns: {set(['group.a', 'group.b', 'estimated.identity'])}", file=sys.stderr)
                    print(f"Found: {exc.args[0][1]}", file=sys.stderr)
                sys.exit(1)
            for _, text in bad:
                print(f"Warning: Invalid similarity value: {text}", file=sys.stderr)
            table = SimilarityTable.from_rows(rows)
            return table, set(table.names)
    except FileNotFoundError:
        print(f"Error: File not found: {filename}", file=sys.stderr)
        sys.exit(1)


def read_subset_file(filename):
    """Stripped non-blank lines that do not start with '#' (h-fst.py:121-128)."""
    try:
        with open(filename) as handle:
            return {line.strip() for line in handle if line.strip() and not line.startswith("#")}
    except FileNotFoundError:
        print(f"Error: Subset file not found: {filename}", file=sys.stderr)
        sys.exit(1)


def _device_sums(table, set_a, set_b, round_digits, ctx, sequence_length=0):
    """One impop_reduce_identity call: labels A / B (and SUBSET = members of both, see calculate_diversity)."""
    both = set(set_a) & set(set_b) if set_b is not None else set()
    labels = table.labels(ctx, a=set_a, b=set_b, subset=both)
    ident = table.device(ctx, round_digits)
    stats, counts, _ = ctx.reduce_identity(ident, labels, None, length=sequence_length or 0)
    ctx.check()
    return stats.cpu().tolist(), counts.cpu().tolist()


def calculate_diversity(similarities, seq_set1, seq_set2=None, round_digits=None, ctx=None):
    """(mean of 1 - identity over the pairs present in the table, their number, pairs absent) -- h-fst.py:130-171.

    seq_set2 None: all unordered pairs within seq_set1; else every (a in set1, b in set2).  A name
    that is in both sets makes the reference visit its pairs with other shared names twice, and
    the pair with itself is looked up and counts as absent unless the table holds it; both are
    reproduced (sums over the shared names are added once more)."""
    ctx = ctx or default_context()
    table = SimilarityTable.from_mapping(similarities)
    set1 = set(seq_set1)
    if seq_set2 is None:
        st, ct = _device_sums(table, set1, None, round_digits, ctx)
        possible = len(set1) * (len(set1) - 1) // 2
        return (st[2], ct[4], possible - ct[4]) if ct[4] else (0.0, 0, possible)   # st[2]: device mean sum_AA / pairs_AA
    else:
        set2 = set(seq_set2)
        st, ct = _device_sums(table, set1, set2, round_digits, ctx)
        total, count = st[17], ct[6]
        shared = set1 & set2
        if not shared:
            return (st[5], ct[6], len(set1) * len(set2) - ct[6]) if ct[6] else (0.0, 0, len(set1) * len(set2))
        if shared:                       # pairs inside the overlap are visited in both orders ...
            total, count = total + st[14], count + ct[3]
            for s in shared:             # ... and (s, s) is looked up once
                i = table.index.get(s)
                if i is not None and table.matrix[i, i] == table.matrix[i, i]:
                    v = float(table.matrix[i, i])
                    if round_digits is not None:
                        v = round(v, round_digits)
                    total, count = total + (1 - v), count + 1
        possible = len(set1) * len(set2)
    if count == 0:
        return 0.0, 0, possible
    return total / count, count, possible - count


def calculate_fst(similarities, pop_a, pop_b, sequence_length=None, round_digits=None, log_file=None, ctx=None):
    """dict(fst, pi_a, pi_b, pi_xy, dxy, da) by Hudson et al. (1992) -- h-fst.py:173-249."""
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
    st, ct = _device_sums(table, pop_a, pop_b, round_digits, ctx, sequence_length if (sequence_length and sequence_length > 0) else 0)
    per_site = bool(sequence_length and sequence_length > 0)
    L = sequence_length if per_site else 1
    count_a, count_b, count_ab = ct[4], ct[5], ct[6]
    miss_a = len(pop_a) * (len(pop_a) - 1) // 2 - count_a
    miss_b = len(pop_b) * (len(pop_b) - 1) // 2 - count_b
    miss_ab = len(pop_a) * len(pop_b) - count_ab
    # raw (not per-site) values for the log: the device divided by L (h-fst.py:225-240); the log shows both
    raw = {k: (st[c] * L if per_site else st[c]) for k, c in (("pi_a", 2), ("pi_b", 3), ("pi_xy", 4), ("dxy", 5))}
    fst = st[7]
    log_print("FST Calculation")
    log_print("=" * 50)
    log_print(f"Population A: {len(pop_a)} sequences")
    log_print(f"Population B: {len(pop_b)} sequences")
    if round_digits is not None:
        log_print(f"Rounding similarities to {round_digits} decimal places")
    log_print("")
    log_print("Within-population diversity (π):")
    log_print(f"  πA = {raw['pi_a']:.6f} (from {count_a} pairs, {miss_a} missing)")
    log_print(f"  πB = {raw['pi_b']:.6f} (from {count_b} pairs, {miss_b} missing)")
    log_print(f"  πXY = {raw['pi_xy']:.6f} (average of πA and πB)")
    log_print("")
    log_print("Between-population diversity (Dxy):")
    log_print(f"  Dxy = {raw['dxy']:.6f} (from {count_ab} pairs, {miss_ab} missing)")
    log_print("")
    if raw["dxy"] > 0:
        log_print("FST calculation:")
        log_print("  FST = (Dxy - πXY) / Dxy")
        log_print(f"      = ({raw['dxy']:.6f} - {raw['pi_xy']:.6f}) / {raw['dxy']:.6f}")
        log_print(f"      = {fst:.6f}")
    else:
        log_print("FST = 0 (Dxy = 0)")
    if per_site:
        log_print("")
        log_print(f"Per-site values (sequence length = {sequence_length:,}):")
        log_print(f"  πA per site = {st[2]:.8f}")
        log_print(f"  πB per site = {st[3]:.8f}")
        log_print(f"  πXY per site = {st[4]:.8f}")
        log_print(f"  Dxy per site = {st[5]:.8f}")
    return {"fst": fst, "pi_a": st[2], "pi_b": st[3], "pi_xy": st[4], "dxy": st[5], "da": st[6]}


def main(argv=None):
    parser = argparse.ArgumentParser(
        description="Calculate FST from pairwise sequence similarities",
        formatter_class=argparse.RawDescriptionHelpFormatter,
        epilog="Output format:\n  FST<tab>pi_A<tab>pi_B<tab>pi_XY<tab>Dxy<tab>Da\n"
               "FST = (Dxy - pi_XY) / Dxy (Hudson et al. 1992); pi_XY = (pi_A + pi_B) / 2; Da = Dxy - pi_XY")
    parser.add_argument("similarity_file", help="TSV file with columns: group.a, group.b, estimated.identity")
    parser.add_argument("-a", "--pop-a", required=True, help="File listing sequence IDs for population A")
    parser.add_argument("-b", "--pop-b", required=True, help="File listing sequence IDs for population B")
    parser.add_argument("-l", "--length", type=int, default=None, help="Sequence length for per-site calculations")
    parser.add_argument("-r", "--round", type=int, default=None, help="Round similarities to N decimal places")
    parser.add_argument("-d", "--log-dir", default=".", help="Directory for log file (default: current directory)")
    parser.add_argument("-v", "--verbose", action="store_true", help="Print detailed progress to stderr")
    args = parser.parse_args(argv)

    if args.verbose:
        print(f"Reading similarity file: {args.similarity_file}", file=sys.stderr)
    similarities, all_sequences = read_similarity_file(args.similarity_file)
    if args.verbose:
        print("Reading population files...", file=sys.stderr)
    pop_a_raw = read_subset_file(args.pop_a)
    pop_b_raw = read_subset_file(args.pop_b)
    pop_a, missing_a = expand_population(pop_a_raw, all_sequences)
    pop_b, missing_b = expand_population(pop_b_raw, all_sequences)
    if args.verbose:
        print(f"Population A candidates: {len(pop_a_raw)}", file=sys.stderr)
        print(f"Population B candidates: {len(pop_b_raw)}", file=sys.stderr)
        print(f"Population A sequences matched: {len(pop_a)}", file=sys.stderr)
        print(f"Population B sequences matched: {len(pop_b)}", file=sys.stderr)
    if missing_a:
        print(f"Warning: {len(missing_a)} identifiers from population A did not match any sequences", file=sys.stderr)
    if missing_b:
        print(f"Warning: {len(missing_b)} identifiers from population B did not match any sequences", file=sys.stderr)
    if not pop_a or not pop_b:
        print("Error: No valid sequences found in one or both populations", file=sys.stderr)
        sys.exit(1)

    base_name = os.path.splitext(os.path.basename(args.similarity_file))[0]
    log_path = os.path.join(args.log_dir, f"{base_name}_fst.log")
    os.makedirs(args.log_dir, exist_ok=True)
    with open(log_path, "w") as log_file:
        results = calculate_fst(similarities, pop_a, pop_b, sequence_length=args.length, round_digits=args.round,
                                log_file=log_file)
    print("\t".join(f"{results[k]:.8f}" for k in ("fst", "pi_a", "pi_b", "pi_xy", "dxy", "da")))
    if args.verbose:
        print(f"Detailed log saved to: {log_path}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
