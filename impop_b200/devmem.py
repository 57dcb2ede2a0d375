"""Plain device buffers over the C ABI (impop_dev_alloc / impop_dev_copy): what the TSV-mode command lines use instead
of a tensor library.  `scripts/pica2.py`, `h-fst.py`, `af.py`, `tj_d.py` and `hud.py` are started once per BED window by
the reference's wrappers (run_pica2_impg.sh:175, run_h-fst.sh:74-85, run_tajd.sh:166-180): importing torch there costs
seconds per call, so those paths hold their few small arrays in `DevArray`s.  A DevArray offers the handful of tensor
methods the engine and the drop-ins touch (`data_ptr`, `shape`, `stride`, `cpu().tolist()` ...), nothing more.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


class HostArray:
    """Result of DevArray.cpu(): a numpy array with the two accessors the drop-ins use."""

    def __init__(self, a: np.ndarray):
        self._a = a

    def numpy(self) -> np.ndarray:
        return self._a

    def tolist(self):
        return self._a.tolist()

    def __getitem__(self, k):
        return self._a[k]


class DevArray:
    def __init__(self, ctx, shape, dtype):
        self.ctx = ctx
        self.shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize if self.shape else self.dtype.itemsize
        p = C.c_void_p()
        ctx._call("impop_dev_alloc", self.nbytes, C.byref(p))
        self._ptr = p.value or 0
        self.is_cuda = True

    @classmethod
    def from_numpy(cls, ctx, a: np.ndarray):
        a = np.ascontiguousarray(a)
        out = cls(ctx, a.shape, a.dtype)
        if a.nbytes:
            ctx._call("impop_dev_copy", C.c_void_p(out._ptr), C.c_void_p(a.ctypes.data), a.nbytes, 0, None)
        return out

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        if out.nbytes:
            self.ctx._call("impop_dev_copy", C.c_void_p(out.ctypes.data), C.c_void_p(self._ptr), out.nbytes, 1, None)
        return out

    def cpu(self) -> HostArray:
        return HostArray(self.numpy())

    def data_ptr(self) -> int:
        return self._ptr

    def dim(self) -> int:
        return len(self.shape)

    def numel(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    def stride(self, k: int) -> int:
        return int(np.prod(self.shape[k + 1:], dtype=np.int64)) if k + 1 < len(self.shape) else 1

    def contiguous(self):
        return self

    def zero_(self):
        if self.nbytes:
            z = np.zeros(self.shape, dtype=self.dtype)
            self.ctx._call("impop_dev_copy", C.c_void_p(self._ptr), C.c_void_p(z.ctypes.data), z.nbytes, 0, None)
        return self

    def close(self):
        if self._ptr and getattr(self.ctx, "handle", None):
            self.ctx.lib.impop_dev_free(self.ctx.handle, C.c_void_p(self._ptr))
        self._ptr = 0

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass
