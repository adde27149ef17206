"""The bound behind the adjacency sweep's speculative batches (csrc/cont_adjacency.cu, k_adj_sweep), checked numerically on
the CPU.  A later bucket's log-intensity sum is S = sum_i log(1 + g_i / b_i) with b_i >= lambda0 the intensity of event i without that
bucket's parent.  Accepted flips of earlier buckets move every b_i by at most `on` upwards (links that went on) and at most `off` downwards
(links that went off, never below lambda0).  Claim:  S / (1 + on / lambda0) <= S' <= S (1 + off / lambda0)  for changes of ANY size --
from log(1 + r x) <= r log(1 + x) for r >= 1 (Bernoulli).  The kernel certifies a decision "off" against the upper end and a decision
"on" against the lower end; this test draws adversarial and random shifts and checks both ends, including the lower end in the form the
kernel uses, S - S' <= S on / (lambda0 + on)."""
import numpy as np
import pytest


def sums(b, g):
    return float(np.sum(np.log1p(g / b)))


@pytest.mark.parametrize("lam0", [1e-3, 0.032, 1.0, 50.0])
@pytest.mark.parametrize("seed", range(8))
def test_one_sided_bounds_hold_for_shifts_of_any_size(lam0, seed):
    rng = np.random.default_rng(1000 * seed + 7)
    m = 2000
    b = lam0 + rng.gamma(0.5, 2.0 * lam0, m) * (rng.random(m) < 0.7)    # many events sit at the baseline rate itself
    g = rng.gamma(0.3, 3.0 * lam0, m) * (rng.random(m) < 0.6)           # contributions of the later bucket, some zero (padding / other events)
    S = sums(b, g)
    for scale in (1e-3, 0.3, 1.0, 30.0, 1e4):                          # largest single contribution of the flipped buckets, in units of lambda0
        on, off = scale * lam0 * rng.uniform(0.5, 1.0), scale * lam0 * rng.uniform(0.5, 1.0)
        for mode in ("extreme", "random", "mixed"):
            if mode == "extreme":
                up, dn = np.full(m, on), np.full(m, off)
            elif mode == "random":
                up, dn = on * rng.random(m), off * rng.random(m)
            else:
                up, dn = on * (rng.random(m) < 0.5), off * (rng.random(m) < 0.5)
            # links that went on only / off only / both; an intensity never falls below lambda0 (it is lambda0 + non-negative contributions)
            b_on = b + up
            b_off = np.maximum(b - dn, lam0)
            b_mix = np.maximum(b + up - dn, lam0)
            tol = 1e-12 * max(S, 1.0)
            for bp in (b_on, b_off, b_mix):
                Sp = sums(bp, g)
                assert Sp <= S * (1.0 + off / lam0) + tol
                assert Sp >= S / (1.0 + on / lam0) - tol
                assert S - Sp <= S * (on / (lam0 + on)) + tol            # the kernel's form of the lower end
            assert sums(b_on, g) <= S + tol                              # links going on can only lower a later sum ...
            assert sums(b_off, g) >= S - tol                             # ... and links going off can only raise it


def test_bounds_are_tight_in_the_small_contribution_limit():
    """With g << b the terms are ~ g / b, so a uniform shift of every b = lambda0 by +on scales S by exactly 1 / (1 + on / lambda0):
    the lower end cannot be improved without knowing which events the buckets share."""
    lam0, on = 0.05, 0.2
    b = np.full(500, lam0)
    g = np.full(500, 1e-9)
    S, Sp = sums(b, g), sums(b + on, g)
    assert abs(Sp / S - 1.0 / (1.0 + on / lam0)) < 1e-6
