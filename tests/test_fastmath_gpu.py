"""Accuracy of the table-driven FP64 log/exp the impulse evaluation uses (csrc/fastmath.cuh):
log: absolute error <= 4e-16 * max(1, |log x|); exp: relative error <= 4e-16 -- both far inside the
1e-10 parity tolerance of the intensities they feed."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eval(ctx, which, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    ctx.check(ctx.lib.nhp_test_fastmath(ctx.h, which, x.ctypes.data_as(ctypes.c_void_p), x.size, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def test_fast_log(ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 200000), 1.0 + rng.uniform(-1e-6, 1e-6, 1000),
                        np.array([1.0, 2.0, 0.5, 1e-310, 5e-324, np.nextafter(1.0, 0), np.nextafter(1.0, 2)])])
    got = _eval(ctx, 0, x)
    ref = np.log(x.astype(np.longdouble)).astype(np.float64)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() <= 4e-16, err.max()


def test_fast_exp(ctx):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-707, 709, 300000), rng.uniform(-2, 2, 100000), np.array([0.0, -0.0, 1.0, -1.0, 709.5, -706.99, 1e-20])])
    got = _eval(ctx, 1, x)
    ref = np.exp(x.astype(np.longdouble)).astype(np.float64)
    rel = np.abs(got - ref) / ref
    assert rel.max() <= 4e-16, rel.max()
    # flush below -707, libdevice above 709, NaN propagates
    edge = _eval(ctx, 1, np.array([-708.0, -1e9, -np.inf, 710.0, np.inf, np.nan]))
    assert edge[0] == 0.0 and edge[1] == 0.0 and edge[2] == 0.0 and np.isinf(edge[3]) and np.isinf(edge[4]) and np.isnan(edge[5])
