"""CPU oracle for the impop windowed population-statistics hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``impop_b200/`` (the product) may
import, link or execute anything from this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs use it, and there only as the checker / reported baseline.

Parity status
-------------
* ``oracle.popstats`` (rows a-1 .. a-10 of SURVEY.md section 8) restates the
  reference's four Python scripts.  It is PINNED: ``tests/golden/`` holds
  outputs of the *unmodified* reference scripts (imported from
  /root/reference by ``tests/golden/make_golden.py`` in the build container)
  and ``tests/test_oracle_golden.py`` checks the restatement against them.
* ``oracle.similarity`` / ``oracle/csrc/oracle_impop.c`` (row a-0, the
  all-pairs node-length-weighted intersection / union / estimated.identity)
  restate arithmetic that lives in third-party tools that are NOT under
  /root/reference: pangenome/odgi (``odgi similarity``, C++) and
  pangenome/impg (``impg similarity``, Rust).  The reference pins no version
  of either (README.md:5 only links the repos) and holds no test or golden
  vector at that boundary, so for row a-0 **parity is unpinned**: the
  restatement follows the published definition and
  is anchored on the reference's call sites and on the columns its scripts
  read (pica2.py:22, h-fst.py:94, af.py:13-15).
"""
