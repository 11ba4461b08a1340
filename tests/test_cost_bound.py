"""CPU check of the error analysis behind the kernels' exact-cost pruning.

multibox_b200/csrc/mbx_bound.h defines a 4-FMA approximation a(i,j) + G_i of the
reference's cost entry (loss.py:35) and a margin m_j + mg_i with the claim
|c_exact - (a + G)| <= m_j + mg_i.  The same header is compiled as plain C into the
oracle library (oracle/c/mbx_oracle.c: orc_bound_max_ratio) next to the bit-exact
cost_entry(); this test evaluates error / margin over random, structured and
adversarial inputs.  The ratio must stay <= 1 (it stays below ~0.4 in practice)."""
import ctypes

import numpy as np
import pytest

from multibox_b200 import synth
from oracle import c_oracle


def _approx_log(lc, rng=None):
    """A log(c) as wrong as the kernels' fast log may be: |lc' - lc| = 2^-21 + 2^-20 |lc| (the documented bound
    of __logf against the true log plus numpy's own 4 ulp), random sign."""
    lc = np.asarray(lc, np.float32)
    rng = np.random.default_rng(0) if rng is None else rng
    sign = rng.choice([-1.0, 1.0], size=lc.shape)
    with np.errstate(all="ignore"):
        return (lc.astype(np.float64) + sign * (2.0 ** -21 + 2.0 ** -20 * np.abs(lc.astype(np.float64)))).astype(np.float32)


def _ratio(loc, lc, l1, gt, alpha, lc_w=None):
    L = c_oracle.lib()
    L.orc_bound_max_ratio.restype = ctypes.c_double
    loc = np.ascontiguousarray(loc, np.float32)
    lc = np.ascontiguousarray(lc, np.float32)
    lc_w = np.ascontiguousarray(_approx_log(lc) if lc_w is None else lc_w, np.float32)
    l1 = np.ascontiguousarray(l1, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    unb = ctypes.c_int64(0)
    r = L.orc_bound_max_ratio(c_oracle._p(loc), c_oracle._p(lc), c_oracle._p(lc_w), c_oracle._p(l1), c_oracle._p(gt),
                              ctypes.c_int64(loc.shape[0]), ctypes.c_int64(gt.shape[0]), ctypes.c_float(alpha),
                              ctypes.byref(unb))
    return r, unb.value


def _logs(c):
    c = np.ascontiguousarray(c, np.float32)
    with np.errstate(all="ignore"):
        lc = c_oracle.nplog(c)
        v = (np.float32(1.0) - c).astype(np.float32)
        v[v > 1] = 1
        v[v <= 0] = np.float32(1e-10)
        return lc, c_oracle.nplog(v)


@pytest.mark.parametrize("seed", range(6))
def test_bound_holds_on_random_scales(seed):
    rng = np.random.default_rng(seed)
    worst = 0.0
    for trial in range(60):
        P, n = 512, 24
        scale = 10.0 ** rng.uniform(-8, 8)
        gscale = scale * 10.0 ** rng.uniform(-3, 3)
        loc = (rng.standard_normal((P, 4)) * scale).astype(np.float32)
        gt = (rng.standard_normal((n, 4)) * gscale).astype(np.float32)
        if trial % 4 == 0:      # saturated sigmoids
            c = (1 / (1 + np.exp(-rng.standard_normal(P) * 12))).astype(np.float32)
        else:
            c = rng.random(P).astype(np.float32)
        c = np.clip(c, 1e-38, 1).astype(np.float32)
        if trial % 5 == 0:      # GT boxes that nearly coincide with predictions: the expanded square cancels
            gt = (loc[:n].astype(np.float64) * (1 + rng.standard_normal((n, 4)) * 1e-6)).astype(np.float32)
        if trial % 7 == 0:      # one coordinate dominates
            loc[:, 0] *= np.float32(1e4)
        alpha = float(10.0 ** rng.uniform(-6, 8))
        lc, l1 = _logs(c)
        r, _ = _ratio(loc, lc, l1, gt, alpha)
        worst = max(worst, r)
    assert worst <= 1.0, worst


def test_bound_holds_on_the_benchmark_configs():
    for name in ("cfg2", "cfg4"):
        cfg = dict(synth.TRAIN_CONFIGS[name])
        cfg["B"] = 8
        d = synth.make_train_inputs(edge_cases=True, **cfg)
        B, P = d["B"], d["P"]
        loc = (d["locations"].reshape(B, P, 4) + d["priors"][None]).astype(np.float32)
        conf = (d["confidences"].reshape(B, P) + np.float32(1e-10)).astype(np.float32)
        for b in range(B):
            n = int(d["num_gt"][b])
            if n == 0:
                continue
            lc, l1 = _logs(conf[b])
            r, unb = _ratio(loc[b], lc, l1, d["gt"][b, :n], d["alpha"])
            assert r <= 1.0 and unb == 0, (name, b, r, unb)


def test_non_finite_inputs_disable_the_bound():
    """inf / NaN coordinates or log terms must give an infinite margin (nothing is pruned with it),
    never a finite margin around a meaningless approximation."""
    rng = np.random.default_rng(3)
    P, n = 64, 4
    loc = rng.random((P, 4)).astype(np.float32)
    gt = rng.random((n, 4)).astype(np.float32)
    c = rng.uniform(0.1, 0.9, P).astype(np.float32)
    lc, l1 = _logs(c)
    loc[3, 1] = np.inf
    loc[5, 2] = np.nan
    lc[7] = -np.inf          # confidence 0: cost +inf
    lc[9] = np.nan           # negative confidence
    gt[2, 0] = np.inf
    r, unb = _ratio(loc, lc, l1, gt, 1000.0)
    assert r <= 1.0
    assert unb == 4 * (n - 1) + P      # the four bad priors against the good rows + every prior against the bad row
    r, unb = _ratio(loc[10:], lc[10:], l1[10:], gt[:2], float("inf"))
    assert unb == (P - 10) * 2
    r, unb = _ratio((loc[10:] * np.float32(1e25)), lc[10:], l1[10:], gt[:2], 1000.0)   # squares overflow
    assert r <= 1.0 and unb == (P - 10) * 2
