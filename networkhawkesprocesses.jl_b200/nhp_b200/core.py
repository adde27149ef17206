"""Device context and device-resident data handles (thin RAII wrappers over the C ABI)."""
import ctypes
import weakref

import numpy as np

from . import _lib
from ._lib import NHPError, check  # noqa: F401


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _fmat(a):
    """Julia-ordered (column-major) flat view of a matrix X[parent, child] -> X[parent + K*child]."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel() if np.ndim(a) >= 2 else _f64(a)


class Context:
    """One GPU + one stream (nhp_ctx)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.nhp_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise NHPError(rc, (self.lib.nhp_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self._fin = weakref.finalize(self, self.lib.nhp_destroy, h)

    def check(self, rc):
        check(self.h, rc)

    @property
    def launches(self):
        return int(self.lib.nhp_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.nhp_last_kernel_ms(self.h))


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


class ContinuousData:
    """Device-resident `(events, nodes, duration)` (nhp_events).  Upload once, evaluate many times."""

    def __init__(self, ctx, events, nodes, duration, K, n_halo=0, index_base=0, flags=1):
        self.ctx = ctx
        ev = _f64(events)
        nd = _i64(nodes)
        if ev.shape != nd.shape:
            raise ValueError("events and nodes must have the same length")
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.nhp_events_upload(ctx.h, _ptr(ev), _ptr(nd), ev.size, float(duration), int(K), int(n_halo), int(index_base),
                                            int(flags), ctypes.byref(h)))
        self.h = h
        self.n = ev.size
        self.n_own = ev.size - int(n_halo)
        self.n_halo = int(n_halo)
        self.index_base = int(index_base)
        self.duration = float(duration)
        self.K = int(K)
        self._fin = weakref.finalize(self, ctx.lib.nhp_events_free, ctx.h, h)

    @classmethod
    def from_handle(cls, ctx, h, K):
        """Wrap an events handle the library created on the device (nhp_cont_rand)."""
        self = cls.__new__(cls)
        self.ctx, self.h, self.K = ctx, h, int(K)
        dur = ctypes.c_double()
        ctx.check(ctx.lib.nhp_events_download(ctx.h, h, None, None, ctypes.byref(dur)))
        self.n = self.n_own = int(ctx.lib.nhp_events_count(h))
        self.n_halo, self.index_base, self.duration = 0, 0, dur.value
        self._fin = weakref.finalize(self, ctx.lib.nhp_events_free, ctx.h, h)
        return self

    def download(self):
        """(events, nodes, duration) on the host, in the reference's conventions (1-based Int64 nodes)."""
        t, c = np.empty(self.n_own, dtype=np.float64), np.empty(self.n_own, dtype=np.int64)
        self.ctx.check(self.ctx.lib.nhp_events_download(self.ctx.h, self.h, _ptr(t), _ptr(c), None))
        return t, c, self.duration

    def free(self):
        self._fin()
