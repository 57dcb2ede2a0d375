"""Import the UNMODIFIED reference scripts from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing
that runs there may call this; it exists to (re)generate tests/golden/ and to let the
CPU test-suite cross-check the restatement live when the reference tree is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys

REFERENCE_SCRIPTS = os.environ.get("IMPOP_REFERENCE_SCRIPTS", "/root/reference/scripts")

_FILES = {
    "pica2": "pica2.py",
    "hfst": "h-fst.py",
    "tj_d": "tj_d.py",
    "af": "af.py",
    "hud": os.path.join("hudson", "hud.py"),
}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_SCRIPTS, f)) for f in _FILES.values())


def load(name: str):
    """Load one reference script as module `impop_reference_<name>`.

    The module is registered in sys.modules before exec_module: tj_d.py combines
    `from __future__ import annotations` with @dataclass, which resolves the module
    through sys.modules on Python 3.12.
    """
    modname = f"impop_reference_{name}"
    if modname in sys.modules:
        return sys.modules[modname]
    path = os.path.join(REFERENCE_SCRIPTS, _FILES[name])
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod
