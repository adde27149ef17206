"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see nhp_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_double, c_int, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(_HERE, "nhp_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return SO


class ContModel(Structure):
    _fields_ = [("kind", c_int), ("K", c_int64), ("lambda0", c_void_p), ("W", c_void_p), ("A", c_void_p), ("p1", c_void_p),
                ("p2", c_void_p), ("dtmax", c_double)]


class DiscModel(Structure):
    _fields_ = [("N", c_int64), ("B", c_int64), ("lambda0", c_void_p), ("W", c_void_p), ("A", c_void_p), ("theta", c_void_p), ("dt", c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO)
        _lib.orc_exponential_pdf.restype = c_double
        _lib.orc_exponential_pdf.argtypes = [c_double, c_double]
        _lib.orc_logitnormal_pdf.restype = c_double
        _lib.orc_logitnormal_pdf.argtypes = [c_double, c_double, c_double]
        _lib.orc_disc_loglik.restype = c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def fmat(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel()


def unf(v, K):
    return np.asarray(v).reshape(K, K).T.copy()


class Cont:
    """Continuous model in the oracle's layout.  Matrices are passed as numpy [parent, child]."""

    def __init__(self, kind, lambda0, W, p1, p2=None, A=None, dtmax=np.inf):
        self.K = len(lambda0)
        self.keep = dict(lambda0=np.ascontiguousarray(lambda0, dtype=np.float64), W=fmat(W), A=fmat(A), p1=fmat(p1), p2=fmat(p2))
        k = self.keep
        self.m = ContModel(int(kind), self.K, _p(k["lambda0"]), _p(k["W"]), _p(k["A"]), _p(k["p1"]), _p(k["p2"]), float(dtmax))

    def _ev(self, events, nodes):
        return np.ascontiguousarray(events, dtype=np.float64), np.ascontiguousarray(nodes, dtype=np.int64)

    def loglik(self, events, nodes, duration, recursive=True):
        ev, nd = self._ev(events, nodes)
        ll = c_double()
        rc = lib().orc_cont_loglik(ctypes.byref(self.m), _p(ev), _p(nd), c_int64(ev.size), c_double(duration), int(recursive), ctypes.byref(ll))
        assert rc == 0
        return ll.value

    def loglik_grad(self, events, nodes, duration, recursive=True):
        """(ll, dlambda0[K], dW[K,K], dp1[K,K], dp2[K,K]) -- extension, see orc_cont_loglik_grad"""
        ev, nd = self._ev(events, nodes)
        K = self.K
        ll = c_double()
        dl0, dW, d1, d2 = np.zeros(K), np.zeros(K * K), np.zeros(K * K), np.zeros(K * K)
        rc = lib().orc_cont_loglik_grad(ctypes.byref(self.m), _p(ev), _p(nd), c_int64(ev.size), c_double(duration), int(recursive), ctypes.byref(ll),
                                        _p(dl0), _p(dW), _p(d1), _p(d2))
        assert rc == 0
        f = lambda v: v.reshape(K, K).T.copy()  # [parent, child]
        return ll.value, dl0, f(dW), f(d1), f(d2)

    def event_intensity(self, events, nodes):
        ev, nd = self._ev(events, nodes)
        out = np.empty(ev.size)
        lib().orc_cont_event_intensity(ctypes.byref(self.m), _p(ev), _p(nd), c_int64(ev.size), _p(out))
        return out

    def intensity(self, events, nodes, times):
        ev, nd = self._ev(events, nodes)
        tq = np.ascontiguousarray(times, dtype=np.float64)
        out = np.empty(tq.size * self.K)
        lib().orc_cont_intensity(ctypes.byref(self.m), _p(ev), _p(nd), c_int64(ev.size), _p(tq), c_int64(tq.size), _p(out))
        return out.reshape(self.K, tq.size).T.copy()

    def resample_parents(self, events, nodes, u):
        ev, nd = self._ev(events, nodes)
        uu = np.ascontiguousarray(u, dtype=np.float64)
        par, pn = np.empty(ev.size, np.int64), np.empty(ev.size, np.int64)
        rc = lib().orc_cont_resample_parents(ctypes.byref(self.m), _p(ev), _p(nd), c_int64(ev.size), _p(uu), _p(par), _p(pn))
        if rc != 0:
            raise ValueError("Categorical: the probability vector is invalid")
        return par, pn

    def resample_adjacency(self, A, rho, events, nodes, duration, u, col_begin=0, col_stride=1):
        ev, nd = self._ev(events, nodes)
        Af, rf, uf = fmat(A).copy(), fmat(rho), fmat(u)
        lib().orc_cont_resample_adjacency_cols(ctypes.byref(self.m), _p(Af), _p(rf), _p(ev), _p(nd), c_int64(ev.size), c_double(duration), _p(uf),
                                               c_int64(col_begin), c_int64(col_stride))
        return unf(Af, self.K)


def suffstats(kind, events, nodes, parents, parentnodes, K, dtmax):
    """dict(M0, Mn, Mnm, S1, S2) with the same meaning as nhp_cont_suffstats."""
    L = lib()
    ev = np.ascontiguousarray(events, dtype=np.float64)
    nd = np.ascontiguousarray(nodes, dtype=np.int64)
    par = np.ascontiguousarray(parents, dtype=np.int64)
    pn = np.ascontiguousarray(parentnodes, dtype=np.int64)
    n = c_int64(ev.size)
    Kc = c_int64(K)
    M0, Mn, Mnm = np.empty(K), np.empty(K), np.empty(K * K)
    L.orc_baseline_counts(_p(nd), _p(pn), n, Kc, _p(M0))
    L.orc_node_counts(_p(nd), n, Kc, _p(Mn))
    L.orc_parent_counts(_p(nd), _p(pn), n, Kc, _p(Mnm))
    S1, S2 = np.zeros(K * K), np.zeros(K * K)
    if kind == 1:
        L.orc_log_duration_sum(_p(ev), _p(nd), _p(par), n, Kc, c_double(dtmax), _p(S1))
        with np.errstate(invalid="ignore", divide="ignore"):
            xbar = np.ascontiguousarray(S1 / Mnm)
        L.orc_log_duration_variation(_p(xbar), _p(ev), _p(nd), _p(par), n, Kc, c_double(dtmax), _p(S2))
    else:
        dm = np.empty(K * K)
        L.orc_duration_mean(_p(ev), _p(nd), _p(par), n, Kc, _p(dm))
        S1 = dm * Mnm
    return dict(M0=M0, Mn=Mn, Mnm=unf(Mnm, K), S1=unf(S1, K), S2=unf(S2, K), duration_mean=None if kind == 1 else unf(dm, K))


# ---- discrete ---------------------------------------------------------------------------
def disc_basis(L, B, dt=1.0):
    phi = np.empty(L * B)
    lib().orc_disc_basis(c_int64(L), c_int64(B), c_double(dt), _p(phi))
    return phi.reshape(B, L).T.copy()  # [l, b]


def disc_convolve(data, phi):
    """data [N, T] int64 (numpy row = node) -> conv [T, N, B]."""
    N, T = data.shape
    L, B = phi.shape
    d = np.ascontiguousarray(np.asarray(data, dtype=np.int64).T).ravel()  # data[n + N*t]
    ph = np.ascontiguousarray(phi.T).ravel()
    conv = np.empty(T * N * B)
    lib().orc_disc_convolve(_p(d), c_int64(N), c_int64(T), _p(ph), c_int64(L), c_int64(B), _p(conv))
    return conv.reshape(B, N, T).transpose(2, 1, 0).copy()


class Disc:
    def __init__(self, lambda0, W, theta, dt=1.0, A=None):
        self.N = len(lambda0)
        self.B = theta.shape[2]
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).transpose(2, 1, 0)).ravel()  # theta[p + N*(c + N*b)]
        self.keep = dict(lambda0=np.ascontiguousarray(lambda0, dtype=np.float64), W=fmat(W), A=fmat(A), theta=th)
        k = self.keep
        self.m = DiscModel(self.N, self.B, _p(k["lambda0"]), _p(k["W"]), _p(k["A"]), _p(k["theta"]), float(dt))

    @staticmethod
    def _data(data):
        return np.ascontiguousarray(np.asarray(data, dtype=np.int64).T).ravel()

    @staticmethod
    def _conv(conv):
        return np.ascontiguousarray(conv.transpose(2, 1, 0)).ravel()  # conv[t + T*(n + N*b)]

    def intensity(self, conv):
        T = conv.shape[0]
        lam = np.empty(T * self.N)
        lib().orc_disc_intensity(ctypes.byref(self.m), _p(self._conv(conv)), c_int64(T), _p(lam))
        return lam.reshape(self.N, T).T.copy()

    def loglik(self, data, conv):
        T = conv.shape[0]
        return lib().orc_disc_loglik(ctypes.byref(self.m), _p(self._data(data)), _p(self._conv(conv)), c_int64(T))

    def gibbs_counts(self, data, conv, u):
        T = conv.shape[0]
        NK = 1 + self.N * self.B
        counts = np.empty(self.N * NK)
        uu = np.ascontiguousarray(u, dtype=np.float64)
        rc = lib().orc_disc_gibbs_counts(ctypes.byref(self.m), _p(self._data(data)), _p(self._conv(conv)), c_int64(T), _p(uu), c_int64(uu.size), _p(counts))
        assert rc == 0
        return counts.reshape(NK, self.N).T.copy()  # [c, k]

    def resample_adjacency(self, A, rho, data, conv, u):
        T = conv.shape[0]
        Af = fmat(A).copy()
        lib().orc_disc_resample_adjacency(ctypes.byref(self.m), _p(Af), _p(fmat(rho)), _p(self._data(data)), _p(self._conv(conv)), c_int64(T), _p(fmat(u)))
        return unf(Af, self.N)


def disc_vb_stats(data, conv, e0, E):
    """E [p, c, b] -> dict(alpha_sum[N], kappa_sum[N,N], nu_sum[N,N], gamma_sum[N,N,B])."""
    N, T = data.shape
    B = conv.shape[2]
    Ef = np.ascontiguousarray(np.asarray(E, dtype=np.float64).transpose(2, 1, 0)).ravel()
    a, k, nu, g = np.empty(N), np.empty(N * N), np.empty(N * N), np.empty(N * N * B)
    lib().orc_disc_vb_stats(c_int64(N), c_int64(B), c_int64(T), _p(Disc._data(data)), _p(Disc._conv(conv)), _p(np.ascontiguousarray(e0, dtype=np.float64)),
                            _p(Ef), _p(a), _p(k), _p(nu), _p(g))
    return dict(alpha_sum=a, kappa_sum=unf(k, N), nu_sum=unf(nu, N), gamma_sum=g.reshape(B, N, N).transpose(2, 1, 0).copy())


def set_threads(n):
    return lib().orc_set_threads(int(n))


def max_threads():
    return lib().orc_get_max_threads()
