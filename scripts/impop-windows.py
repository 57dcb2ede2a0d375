#!/usr/bin/env python3
"""Matrix-mode driver (extension; no counterpart script in the reference): window graphs (GFA) or a binary batch in,
the TSVs of run_pica2_impg.sh / run_h-fst.sh / run_tajd.sh out, one fused GPU pass for all windows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from impop_b200.windows import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
