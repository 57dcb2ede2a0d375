"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), table / name handling, sharding, and the world_size-2 gather logic on gloo."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, unhex


def test_library_exports_every_declared_symbol():
    from impop_b200 import _native
    header = open(os.path.join(ROOT, "include", "impop_b200.h")).read()
    declared = set(re.findall(r"^(?:const char \*|int64_t |int )(impop_\w+)\(", header, flags=re.M))
    assert declared == set(_native.SIGNATURES), (declared ^ set(_native.SIGNATURES))
    lib = _native.lib()                       # builds with nvcc if stale; CDLL load needs no GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.impop_version() == 100
    # constants shared by header and binding
    for macro, value in (("IMPOP_NSTATS", _native.NSTATS), ("IMPOP_NCOUNTS", _native.NCOUNTS)):
        assert int(re.search(rf"#define {macro} (\d+)", header).group(1)) == value
    for key, col in _native.ST.items():
        m = re.search(rf"#define IMPOP_ST_{key.upper()} (\d+)", header)
        assert m and int(m.group(1)) == col, key


def test_no_gpu_means_loud_failure():
    """Without a CUDA device there is nothing to fall back to: impop_create fails, Context raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from impop_b200 import _native
    from impop_b200.engine import Context
    h = ctypes.c_void_p()
    assert _native.lib().impop_create(0, ctypes.byref(h)) != 0 and not h.value
    with pytest.raises(RuntimeError):
        Context(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under impop_b200/ or scripts/ may reference it."""
    for base in ("impop_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                    assert "liboracle" not in text


def test_similarity_table_mapping_and_duplicates(gold, tmp_path):
    from impop_b200.tables import SimilarityTable, read_rows
    p = tmp_path / "m.tsv"
    p.write_text(gold["messy"]["tsv"] + "\n")
    with open(p, newline="") as fh:
        rows, count, bad = read_rows(fh, on_bad_value="skip")
    assert count == 7
    tab = SimilarityTable.from_rows(rows)
    assert tab[("a#1#c:1-2", "b#1#c:1-2")] == 0.995 and tab[("b#1#c:1-2", "a#1#c:1-2")] == 0.995   # last row wins
    assert ("zzz", "a#1#c:1-2") not in tab
    assert len(tab) == len({tuple(sorted(r[:2])) for r in rows})
    # order of the duplicate does not matter for 'max' (af.py links on any row)
    t2 = SimilarityTable.from_rows([("a", "b", 0.9), ("b", "a", 1.0), ("a", "b", 0.5)], combine="max")
    assert t2[("a", "b")] == 1.0
    t3 = SimilarityTable.from_rows([("a", "b", 0.9), ("b", "a", 1.0), ("a", "b", 0.5)])
    assert t3[("a", "b")] == 0.5 and t3[("b", "a")] == 0.5
    # a plain dict as the reference returns it, plus elements that occur in no pair
    t4 = SimilarityTable.from_mapping({("a", "b"): 0.25}, elements={"a", "b", "c"})
    assert t4.names == ["a", "b", "c"] and t4[("a", "b")] == 0.25 and np.isnan(t4.matrix[2]).all()
    # python round semantics, not rint(x * 10^r) / 10^r
    t5 = SimilarityTable(["a", "b"], np.array([[np.nan, 0.999985], [0.999985, np.nan]]))
    assert t5.rounded(5)[0, 1] == round(0.999985, 5)


def test_script_error_behaviour_without_gpu(tmp_path, capsys):
    """Input validation happens before any device work and mirrors the reference's messages / exit codes."""
    from impop_b200 import hfst, pica2
    with pytest.raises(SystemExit) as e:
        pica2.read_similarity_file(str(tmp_path / "nope.tsv"))
    assert e.value.code == 1 and "Error: File not found" in capsys.readouterr().out
    empty = tmp_path / "empty.tsv"
    empty.write_text("")
    with pytest.raises(SystemExit):
        pica2.read_similarity_file(str(empty))
    assert "is empty or missing a header" in capsys.readouterr().out
    bad = tmp_path / "bad.tsv"
    bad.write_text("group.a\tgroup.b\tother\nx\ty\t1\n")
    with pytest.raises(SystemExit):
        pica2.read_similarity_file(str(bad))
    assert "File must contain columns: ['estimated.identity', 'group.a', 'group.b']" in capsys.readouterr().out
    with pytest.raises(SystemExit):
        hfst.read_similarity_file(str(bad))
    assert "File must contain columns" in capsys.readouterr().err
    val = tmp_path / "val.tsv"
    val.write_text("group.a\tgroup.b\testimated.identity\nx\ty\tabc\nx\tz\t0.5\n")
    with pytest.raises(SystemExit):
        pica2.read_similarity_file(str(val))
    assert "Invalid similarity value on line 2: abc" in capsys.readouterr().out
    sim, seqs = hfst.read_similarity_file(str(val))                      # h-fst skips the bad row with a warning
    assert seqs == {"x", "z"} and "Warning: Invalid similarity value: abc" in capsys.readouterr().err
    with pytest.raises(SystemExit):
        hfst.read_subset_file(str(tmp_path / "nope.txt"))
    sub = tmp_path / "s.txt"
    sub.write_text("# c\n\n HG1_hap1 \nHG2\n")
    assert hfst.read_subset_file(str(sub)) == {"HG1_hap1", "HG2"}


def test_canonical_prefix_and_expand(gold):
    from impop_b200 import hfst
    for ident, want in gold["canonical"]:
        assert hfst.canonicalize_identifier(ident) == want, ident
    seqs = {"HG1#1#c:1-2", "HG1#2#c:1-2", "HG2#1#c:1-2", "HG10#1#c:1-2"}
    got, missing = hfst.expand_population(["HG1", "HG2_hap2_hprc_r2", "#skip", ""], seqs)
    assert got == {"HG1#1#c:1-2", "HG1#2#c:1-2"} and missing == ["HG2_hap2_hprc_r2"]


def test_tajima_argument_errors_without_gpu():
    from impop_b200 import tj_d
    with pytest.raises(ValueError):
        tj_d.tajimas_d(1, 1.0, 0.1)
    with pytest.raises(ValueError):
        tj_d.tajimas_d(5, 1.0, -0.1)


def test_shard_bounds():
    from impop_b200.distributed import shard_bounds, shard_of
    assert shard_bounds(10, 3).tolist() == [0, 4, 7, 10]
    assert shard_bounds(5, 8).tolist() == [0, 1, 2, 3, 4, 5, 5, 5, 5]
    assert shard_bounds(0, 4).tolist() == [0, 0, 0, 0, 0]
    assert shard_bounds(4, 2, [1, 1, 1, 9]).tolist() == [0, 3, 4]
    b = shard_bounds(1000, 8, np.ones(1000))
    assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) == 125).all()
    assert shard_of(10, 2, 3) == (7, 10)
    with pytest.raises(ValueError):
        shard_bounds(3, 0)


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from impop_b200.distributed import shard_bounds, gather_rows, gather_parts
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
W = 7
full = torch.arange(W * 3, dtype=torch.float64).reshape(W, 3)
b = shard_bounds(W, 2)
got = gather_rows(full[b[rank]:b[rank + 1]].clone(), b)
assert torch.equal(got, full), got
cnt = gather_rows(torch.arange(W, dtype=torch.int64).reshape(W, 1)[b[rank]:b[rank + 1]].clone(), b)
assert cnt.flatten().tolist() == list(range(W))
parts = gather_parts(torch.full((W, 4), float(rank + 1), dtype=torch.float64))
assert parts.shape == (2, W, 4) and parts[0].eq(1).all() and parts[1].eq(2).all()
# fixed-order sum of the gathered parts is identical on both ranks
tot = parts[0] + parts[1]
other = [torch.empty_like(tot) for _ in range(2)]
dist.all_gather(other, tot)
assert torch.equal(other[0], other[1])
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_world_size_2_gloo(tmp_path):
    """The N > 1 result exchange (SURVEY 8 e) on CPU: unequal window shards all-gathered over gloo."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    for r, p in enumerate(procs):
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert f"ok {r}" in out


def test_pooled_fst_text_matches_the_wrapper_snippets():
    """Row a-10: run_fst_impg.sh:199-218's inline python, outputs stored by tests/golden/make_golden_pooled.py."""
    import json
    from impop_b200 import windows
    from oracle import popstats
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pooled_fst.json")))
    assert len(cases) > 50
    for a, b, c, avg, fst in cases:
        assert windows.pooled_fst_text(a, b, c) == (avg, fst)
        assert popstats.pooled_fst_text(a, b, c) == (avg, fst)


def test_bench_workload_forms_agree():
    """bench.make_workload on the host generator: the affine form it hands to the GPU arm, the tight transfer rows of the
    end-to-end leg and the plain form the CPU oracle is given describe the same windows."""
    import bench
    from impop_b200 import synth
    from oracle import similarity
    cfg = dict(bench.CONFIGS[2])
    wl = bench.make_workload(synth.HostGenerator(), cfg, 6, 11, 2, keep_original=6, plain=6)
    n = wl["n"]
    assert wl["x_off"].shape == (7,) and wl["xt_off"].shape == (7,) and (wl["xt_off"] % 32 == 0).all() and (wl["pitch_w"] % 4 == 0).all()
    for w in range(6):
        pw, tw, mo = int(wl["pitch_w"][w]), int(wl["tp_w"][w]), int(wl["m_out"][w])
        assert tw == max(1, (mo + 31) // 32) and pw * 32 >= mo
        a = wl["x"][wl["x_off"][w]:wl["x_off"][w + 1]].reshape(n, pw)
        t = wl["x_tight"][wl["xt_off"][w]:wl["xt_off"][w] + n * tw].reshape(n, tw)
        assert np.array_equal(a[:, :tw], t) and not a[:, tw:].any()
        lens = wl["len"][wl["len_off"][w]:wl["len_off"][w + 1]]
        assert int(wl["heavy"][w]) == int(((lens.astype(np.int64) // 255 + 254) // 255).sum()) and not lens[mo:].any()
        # affine form == original window == plain form (intersections and pi_ij, bit for bit)
        orig = similarity.pairwise(similarity.unpack_bits(wl["orig_x"][w], wl["m_pad_in"]), wl["orig_len"][w])
        aff = similarity.pairwise_affine(similarity.unpack_bits(a, mo), lens[:mo], wl["row_adj"][w * n:(w + 1) * n], wl["win_const"][w])
        pl = similarity.pairwise(similarity.unpack_bits(wl["plain_x"][w], wl["plain_len"].shape[1]), wl["plain_len"][w])
        for other in (aff, pl):
            assert np.array_equal(orig["I"], other["I"]) and np.array_equal(orig["pi"], other["pi"])
    assert wl["bytes_out"] < wl["bytes_in"] // 2
