"""impop_b200 -- B200-native windowed population statistics (pi / Hudson Fst / Tajima's D / allele
frequencies): the hot path of pangenome/impop as hand-written sm_100a CUDA behind a C ABI
(include/impop_b200.h), with Python drop-ins for scripts/pica2.py, h-fst.py, tj_d.py and af.py.

Importing the package does not touch the GPU; the first kernel call loads libimpop_b200.so and
raises if it (or a CUDA device) is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
