"""Process-wide default device context for the drop-in CLIs and function-level API."""
from __future__ import annotations

import os

_ctx = None


def default_context():
    """One Context on cuda:LOCAL_RANK (cuda:0 outside torchrun) for the script drop-ins (TSV mode).  It is a `lite`
    context: plain device buffers over the C ABI, no torch import -- the wrappers start these command lines once per
    window.  Raises without a GPU: no CPU fallback."""
    global _ctx
    if _ctx is None:
        from .engine import Context
        _ctx = Context(int(os.environ.get("LOCAL_RANK", "0")), lite=True)
    return _ctx


def reset():
    global _ctx
    if _ctx is not None:
        _ctx.close()
    _ctx = None
