"""The native all-pairs table reader (impop_tsv_scan / impop_tsv_fill) against the csv-based reader that mirrors the
reference (tables.read_rows + SimilarityTable.from_rows): identical names and matrix on machine-clean text, and a
refusal (-> general reader) on everything the reference treats specially."""
import io
import json
import os
import time

import numpy as np
import pytest

from impop_b200 import synth, tables
from oracle import similarity

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))


def slow(path):
    with open(path, newline="") as fh:
        rows, count, bad = tables.read_rows(fh, on_bad_value="skip")
    return tables.SimilarityTable.from_rows(rows), count


def same(path):
    fast = tables.read_table_fast(str(path))
    assert fast is not None
    t1, c1 = fast
    t0, c0 = slow(path)
    assert t1.names == t0.names and c1 == c0
    assert np.array_equal(np.isnan(t1.matrix), np.isnan(t0.matrix))
    assert np.array_equal(np.nan_to_num(t1.matrix, nan=-1.0), np.nan_to_num(t0.matrix, nan=-1.0))      # same bits


def test_clean_tables_agree(tmp_path):
    for tag in ("f6", "messy"):                      # messy: extra columns, columns out of order, a repeated pair
        p = tmp_path / f"{tag}.tsv"
        p.write_text(GOLD[tag]["tsv"])
        same(p)
    ws = synth.make_windows(64, 4000, 1, seed=14)
    res = similarity.pairwise(ws.dense(0), ws.node_len[0])
    p = tmp_path / "w.tsv"
    similarity.write_similarity_tsv(str(p), synth.haplotype_names(64, start=1, end=4001), res)
    same(p)
    p = tmp_path / "odd.tsv"                          # CRLF, blank line, exponent / inf / nan tokens, no final newline
    p.write_bytes(b"group.a\tgroup.b\testimated.identity\r\nb\ta\t1e-3\r\n\r\nc\ta\t-INF\r\nc\tb\tnan\r\na\ta\t.5\r\nd\tc\t5.E+0")
    same(p)


@pytest.mark.parametrize("text", [
    "",                                                                   # empty file
    "group.a\tgroup.b\n",                                                 # missing column
    "group.a\tgroup.b\testimated.identity\n\"a\"\tb\t0.5\n",              # csv quoting
    "group.a\tgroup.b\testimated.identity\na\tb\tzero\n",                 # not a number
    "group.a\tgroup.b\testimated.identity\na\tb\t 0.5\n",                 # float() strips blanks; keep that path
    "group.a\tgroup.b\testimated.identity\na\tb\t1_0.5\n",                # float() accepts underscores
    "group.a\tgroup.b\testimated.identity\na\tb\t0x1p-1\n",               # strtod would accept, float() does not
    "group.a\tgroup.b\testimated.identity\na\tb\n",                       # short row
    "group.a\tgroup.b\testimated.identity\n\xe9\tb\t0.5\n",               # non-ASCII name
    "group.a\tgroup.a\tgroup.b\testimated.identity\na\ta\tb\t0.5\n",      # duplicated header name
])
def test_everything_else_goes_to_the_general_reader(tmp_path, text):
    p = tmp_path / "t.tsv"
    p.write_text(text, encoding="utf-8")
    assert tables.read_table_fast(str(p)) is None


def test_readers_of_the_drop_ins_use_it(tmp_path, capsys):
    from impop_b200 import hfst, pica2
    p = tmp_path / "f6.tsv"
    p.write_text(GOLD["f6"]["tsv"])
    t, names, count = pica2.read_similarity_file(str(p))
    t2, names2 = hfst.read_similarity_file(str(p))
    assert count == 15 and names == names2 and len(names) == 6
    assert t[(sorted(names)[0], sorted(names)[1])] == t2[(sorted(names)[0], sorted(names)[1])]


def test_speed_on_a_466_haplotype_table(tmp_path):
    ws = synth.make_windows(466, 20000, 1, seed=15, n_sites_override=60)
    res = similarity.pairwise(ws.dense(0), ws.node_len[0])
    p = tmp_path / "big.tsv"
    similarity.write_similarity_tsv(str(p), synth.haplotype_names(466, start=1, end=20001), res)
    t0 = time.perf_counter(); fast = tables.read_table_fast(str(p)); t1 = time.perf_counter()
    ref, count = slow(p); t2 = time.perf_counter()
    assert fast is not None and fast[1] == count == 466 * 465 // 2
    assert np.array_equal(np.nan_to_num(fast[0].matrix, nan=-1.0), np.nan_to_num(ref.matrix, nan=-1.0))
    print(f"native {1e3 * (t1 - t0):.1f} ms, csv {1e3 * (t2 - t1):.1f} ms")
    assert (t1 - t0) * 3 < (t2 - t1)


def test_scan_memo_is_not_reused_for_other_text():
    """impop_tsv_scan keeps its parse for the impop_tsv_fill that follows; a fill on a buffer whose CONTENT changed at
    the same address, or with no scan before it, must parse the text it is given."""
    import ctypes as C

    from impop_b200 import _native
    lib = _native.lib()
    text = bytearray(b"group.a\tgroup.b\testimated.identity\nx\ty\t0.25\nx\tz\t0.5\ny\tz\t0.75\n")
    buf = (C.c_char * len(text)).from_buffer(text)

    def fill():
        mat = np.empty((3, 3), dtype=np.float64)
        names = np.zeros(16, dtype=np.uint8)
        off = np.zeros(4, dtype=np.int64)
        assert lib.impop_tsv_fill(buf, len(text), mat.ctypes.data, names.ctypes.data, off.ctypes.data) == 0
        return mat

    info = _native.TsvInfo()
    assert lib.impop_tsv_scan(buf, len(text), C.byref(info)) == 0 and info.status == 0 and info.names == 3
    assert fill()[0, 1] == 0.25                                   # from the memo
    assert fill()[0, 1] == 0.25                                   # memo consumed: parsed again
    assert lib.impop_tsv_scan(buf, len(text), C.byref(info)) == 0
    text[text.index(b"0.25"):text.index(b"0.25") + 4] = b"0.35"   # same address, same length, other content
    m = fill()
    assert m[0, 1] == 0.35 and m[1, 0] == 0.35 and m[0, 2] == 0.5 and m[1, 2] == 0.75 and np.isnan(m[0, 0])


def test_af_fast_table_only_when_equivalent(tmp_path):
    """af.py's native reader path (af.load_table_fast) is taken only when it cannot differ from af.py:7-19 + 35-44: one
    row per pair of different samples, names distinct after the cut at ':'; the matrix follows the CUT names' order."""
    from impop_b200 import af
    hdr = "group.a\tgroup.b\testimated.identity\n"
    p = tmp_path / "t.tsv"
    p.write_text(hdr + "b#1#x:9-20\ta#1#y:9-20\t0.5\nb#1#x:9-20\tc:9-20\t0.75\na#1#y:9-20\tc:9-20\t1.0\n")
    t = af.load_table_fast(str(p))
    assert t is not None and t.names == ["a#1#y", "b#1#x", "c"]
    assert t.matrix[0, 1] == 0.5 and t.matrix[1, 2] == 0.75 and t.matrix[0, 2] == 1.0 and t.matrix[2, 0] == 1.0
    rows, samples = af.load_pairs(str(p))
    assert samples == t.names and sorted(r[2] for r in rows) == [0.5, 0.75, 1.0]
    # cutting the coordinates reorders: 'a:5' < 'a#1:5' as full names ('#' < ':' is false for the cut names' first difference)
    p.write_text(hdr + "s10:1-2\ts1#x:1-2\t0.25\n")
    t = af.load_table_fast(str(p))
    assert t is not None and t.names == sorted(["s10", "s1#x"]) and t.matrix[0, 1] == 0.25
    for body in ("a:1-2\tb:1-2\t0.5\na:1-2\tb:1-2\t0.9\n",            # a repeated pair: any row may link (max), not the last
                 "a:1-2\ta:1-2\t1.0\na:1-2\tb:1-2\t0.5\n",            # a self pair
                 "a:1-2\tb:1-2\t0.5\na:3-4\tc:1-2\t0.5\n"):           # two full names of one sample
        p.write_text(hdr + body)
        assert af.load_table_fast(str(p)) is None
