#!/usr/bin/env python3
"""Drop-in for the reference's scripts/hudson/hud.py: same command line and output, computed by impop_b200 on the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from impop_b200.hud import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
