"""Host-side check of the arithmetic claim behind the detect kernel's division-free IoU
decision (multibox_b200/csrc/mbx_detect.cu: iou_fast): whenever the fast test does NOT flag a
pair as `near`, sign(inter - fl(thr*den)) equals the exact decision fl(inter/den) > thr that
oracle/np_oracle.greedy_nms makes.  Emulated in numpy float32, operation for operation, on
random pairs, lattice boxes whose IoUs hit simple fractions exactly, and pairs pushed to within
a few ulp of the threshold."""
import numpy as np
import pytest

F = np.float32
TOL = F(9.5367431640625e-07)     # 2^-20
DEN_MIN = F(1e-20)


def _pairs(a, b):
    """fp32 inter / den exactly as the kernel and the oracle form them."""
    zero = F(0)
    w = np.maximum(zero, np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0]))
    h = np.maximum(zero, np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1]))
    inter = (w * h).astype(F)
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])).astype(F)
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(F)
    den = ((area_a + area_b).astype(F) - inter).astype(F)
    return inter, den


def _fast(inter, den, thr):
    thr = F(thr)
    with np.errstate(over="ignore", invalid="ignore"):
        t = (thr * den).astype(F)
        d = (inter - t).astype(F)
        near = (den > 0) & (~(den >= DEN_MIN) | ~(np.abs(d) > (TOL * t).astype(F)))
    return (d > 0) & (den >= 0), near


def _exact(inter, den, thr):
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / den).astype(F) > F(thr)


def _boxes(rng, n, lattice=None):
    if lattice:
        p = rng.integers(0, lattice + 1, size=(n, 4)).astype(F) / F(lattice)
    else:
        p = rng.random((n, 4)).astype(F)
    x1, x2 = np.minimum(p[:, 0], p[:, 2]), np.maximum(p[:, 0], p[:, 2])
    y1, y2 = np.minimum(p[:, 1], p[:, 3]), np.maximum(p[:, 1], p[:, 3])
    return np.stack([x1, y1, x2, y2], 1).astype(F)


@pytest.mark.parametrize("thr", [0.5, 0.3, 1.0 / 3.0, 0.25, 0.75, 0.05, 1e-5, 0.999])
def test_fast_decision_agrees_when_not_near(thr):
    rng = np.random.default_rng(int(thr * 1e6) + 1)
    n = 400_000
    sets = [(_boxes(rng, n), _boxes(rng, n)),
            (_boxes(rng, n, 8), _boxes(rng, n, 8)),            # IoUs are exact small fractions
            (_boxes(rng, n, 16), _boxes(rng, n, 16)),
            (_boxes(rng, n, 3), _boxes(rng, n, 3))]
    a = _boxes(rng, n)
    sets.append((a, (a + rng.normal(0, 0.02, a.shape)).astype(F)))   # heavy overlap
    n_near = 0
    for a, b in sets:
        inter, den = _pairs(a, b)
        sup, near = _fast(inter, den, thr)
        ex = _exact(inter, den, thr)
        assert np.array_equal(sup[~near], ex[~near])
        n_near += int(near.sum())
        # the near band is narrow: it only ever holds pairs within ~2^-19 of the threshold
        ok = den >= DEN_MIN
        q = inter[ok & near].astype(np.float64) / den[ok & near].astype(np.float64)
        assert np.all(np.abs(q - thr) <= 4e-6 * thr)
    if thr in (0.5, 0.25, 0.75):
        assert n_near > 0      # lattice boxes do hit these thresholds exactly


def test_pairs_forced_to_the_threshold():
    """inter chosen within +-8 ulp of thr*den: the fast test must either agree or say `near`."""
    rng = np.random.default_rng(7)
    for thr in (0.5, 0.3, 0.7):
        den = (rng.random(200_000) * 1.9 + 1e-3).astype(F)
        t = (F(thr) * den).astype(F)
        inter = t.copy()
        for k in range(-8, 9):
            x = inter.copy()
            for _ in range(abs(k)):
                x = np.nextafter(x, F(np.inf) if k > 0 else F(-np.inf)).astype(F)
            sup, near = _fast(x, den, thr)
            ex = _exact(x, den, thr)
            assert np.array_equal(sup[~near], ex[~near])
            assert near.all()      # every such pair is inside the band and goes to the exact path


def test_degenerate_denominators():
    inter = np.array([1e-30, np.nan, 0.1, 1e-25], dtype=F)
    den = np.array([1e-30, 1.0, np.inf, 3e-21], dtype=F)
    _, near = _fast(inter, den, 0.5)
    assert near.all()                      # tiny / NaN / inf: decided by the IEEE division
    # den <= 0 or NaN is decided on the fast path, identically to fl(inter/den) > thr
    inter = np.array([0, 0.25, np.nan, 1e-40, 0, 0.3, np.nan, 0.1], dtype=F)
    den = np.array([0, 0, 0, 0, -0.5, -1e-30, -1.0, np.nan], dtype=F)
    for thr in (0.5, 0.01, 0.99):
        sup, near = _fast(inter, den, thr)
        assert not near.any()
        assert np.array_equal(sup, _exact(inter, den, thr))


def test_invalid_boxes_negative_area():
    """x2 < x1 boxes (the reference never fixes them up) give negative areas and denominators."""
    rng = np.random.default_rng(3)
    n = 300_000
    a = rng.random((n, 4)).astype(F)          # corners NOT ordered
    b = rng.random((n, 4)).astype(F)
    inter, den = _pairs(a, b)
    assert (den < 0).mean() > 0.05
    for thr in (0.5, 0.3):
        sup, near = _fast(inter, den, thr)
        ex = _exact(inter, den, thr)
        assert np.array_equal(sup[~near], ex[~near])
        assert near.mean() < 1e-3
