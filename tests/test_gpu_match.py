"""-m gpu parity tests of the matching kernel (through the C ABI) against the
oracle and the golden vectors generated from the reference's own code.
Matched-prior indices, GT indices and stacked GT rows must be BIT-EXACT."""
import hashlib
import os

import numpy as np
import pytest
import torch

from multibox_b200 import loss, synth
from oracle import c_oracle, np_oracle
from gpu_util import boundary_inputs, dev, gpu_cost_matrix, gpu_nplog

pytestmark = pytest.mark.gpu


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _gpu_assign(loc, conf, gt, ng, B, alpha, warps=0, cols=0, generic=False):
    if warps or cols or generic:
        out = loss.match_loss_raw(dev(loc).view(B, -1, 4), dev(conf).view(B, -1), dev(gt), dev(ng), None, alpha,
                                  flags=2 | (4 if generic else 0), want_mask=True, want_gt_idx=True,
                                  want_stacked=True, want_grads=False, warps=warps, cols=cols)
        loss.raise_for_status(out["results"][2].item())
        n = int(out["n_stacked"].item())
        return out["mask"].cpu().numpy(), out["stacked_gt"][:n].cpu().numpy(), out["matched_gt_idx"].cpu().numpy()
    m, s, g = loss.compute_assignments(dev(loc), dev(conf), dev(gt), dev(ng), B, alpha, return_indices=True)
    return m.cpu().numpy(), s.cpu().numpy(), g.cpu().numpy()


def test_native_library_is_loaded(cuda_device):
    from multibox_b200 import _lib
    lib = _lib.load()
    assert lib.mbx_version() == 100
    assert any("libmultibox_b200.so" in line for line in open("/proc/self/maps"))


def test_nplog_bitwise_vs_numpy(cuda_device):
    bits = np.arange(1, 0x7f800000, 251, dtype=np.uint32)
    x = bits.view(np.float32)
    assert np.array_equal(gpu_nplog(x).view(np.uint32), np.log(x).view(np.uint32))
    with np.errstate(all="ignore"):
        sp = np.array([0.0, np.inf, -1.0, np.nan, 1.0, 1e-10, 1e-45], dtype=np.float32)
        a, b = gpu_nplog(sp), np.log(sp)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_sqrt_is_correctly_rounded(cuda_device):
    """The kernels' branch-free sqrt equals sqrt.rn on EVERY float32 bit pattern (4 x 2^30)."""
    from multibox_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for first in (0x00000000, 0x40000000, 0x80000000, 0xC0000000):
        _lib.check(lib.mbx_debug_sqrt_mismatches(first, 0x40000000, bad.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "sqrt check")
    assert bad.item() == 0
    x = np.array([0.0, 1e-45, 1e-38, 7.9e-31, 2.0, 3.4e38, np.inf], dtype=np.float32)
    assert np.array_equal(np.sqrt(x), np.sqrt(x.astype(np.float64)).astype(np.float32))   # host sqrt is rn too


def test_cost_matrix_bitwise_vs_numpy(cuda_device):
    d = synth.make_train_inputs(K=5, B=3, M=20, seed=21, edge_cases=True)
    loc, conf = boundary_inputs(d)
    P = d["P"]
    for b, alpha in ((1, 1000.0), (2, 1.0)):
        sl = slice(b * P, (b + 1) * P)
        gt = d["gt"][b][:max(1, d["num_gt"][b])]
        lc, l1 = np_oracle.log_terms(conf[sl].copy())
        Cn = np_oracle.cost_matrix(loc[sl], lc, l1, gt, np.float32(alpha))
        Cg = gpu_cost_matrix(loc[sl], conf[sl], gt, alpha)
        assert np.array_equal(Cn.view(np.uint64), Cg.view(np.uint64))


def test_match_small_golden(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "match_small.npz"))
    B = g["locations"].shape[0]
    d = dict(B=B, locations=g["locations"], confidences=g["confidences"], priors=g["priors"])
    loc, conf = boundary_inputs(d)
    m, s, gi = _gpu_assign(loc, conf, g["gt"], g["num_gt"], B, float(g["alpha"]))
    assert m.dtype == np.int32 and s.dtype == np.float32
    assert np.array_equal(m, g["mask"])
    assert np.array_equal(s, g["stacked_gt"])
    assert np.array_equal(gi, g["matched_gt_idx"])


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_match_config_golden(cuda_device, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "match_%s.npz" % name))
    d = synth.make_train_inputs(**synth.TRAIN_CONFIGS[name])
    assert _sha(d["priors"], d["locations"], d["confidences"], d["gt"], d["num_gt"]) == str(g["inputs_sha256"])
    loc, conf = boundary_inputs(d)
    m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], d["B"], d["alpha"])
    assert np.array_equal(np.nonzero(m)[0], g["matched_flat_idx"])
    assert np.array_equal(gi[m == 1], g["matched_gt_idx"])
    assert np.array_equal(s, g["stacked_gt"])


@pytest.mark.parametrize("K,B,M,dist,alpha", [(5, 16, 20, "full", 1.0), (7, 24, 100, "coco_person", 1000.0),
                                              (7, 8, 100, "full", 1000.0), (11, 8, 200, "uniform", 1000.0),
                                              (11, 4, 200, "full", 10.0)])
def test_match_vs_c_oracle(cuda_device, K, B, M, dist, alpha):
    d = synth.make_train_inputs(K=K, B=B, M=M, dist=dist, seed=100 + K + M, alpha=alpha, edge_cases=True)
    loc, conf = boundary_inputs(d)
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, alpha)
    m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], B, alpha)
    assert np.array_equal(m, m0)
    assert np.array_equal(gi, g0)
    assert np.array_equal(s, s0)


# (warps, cols, generic): the generic shared-memory kernel and the register-resident family
VARIANTS = [(1, 0, True), (2, 0, True), (4, 0, True), (8, 0, True),
            (4, 6, False), (4, 8, False), (8, 3, False), (8, 0, False), (16, 2, False), (16, 0, False),
            (2, 0, False)]        # (2, 0): 11 columns per thread -> no instantiation -> generic fallback


@pytest.mark.parametrize("warps,cols,generic", VARIANTS)
def test_match_kernel_variants_agree(cuda_device, warps, cols, generic):
    d = synth.make_train_inputs(K=5, B=12, M=20, dist="uniform", seed=77, edge_cases=True)
    loc, conf = boundary_inputs(d)
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], 12, d["alpha"])
    m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], 12, d["alpha"], warps=warps, cols=cols, generic=generic)
    assert np.array_equal(m, m0) and np.array_equal(gi, g0) and np.array_equal(s, s0)


@pytest.mark.parametrize("P,M,warps", [(33, 5, 0), (33, 5, 1), (96, 30, 1), (200, 64, 0), (1024, 8, 0), (2500, 40, 0),
                                       (4500, 16, 0)])
def test_match_odd_shapes(cuda_device, P, M, warps):
    """Prior counts that are not 129K+1 (single-warp CTAs, partially filled last column, P too
    large for the register family -> generic kernel)."""
    rng = np.random.default_rng(P)
    B = 6
    loc = rng.uniform(0, 1, size=(B * P, 4)).astype(np.float32)
    conf = rng.uniform(0.01, 0.99, size=B * P).astype(np.float32)
    ng = rng.integers(0, M + 1, size=B).astype(np.int32)
    ng[0] = M
    gt = synth.gt_boxes(rng, ng, M)
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, gt, ng, B, 100.0)
    m, s, gi = _gpu_assign(loc, conf, gt, ng, B, 100.0, warps=warps)
    assert np.array_equal(m, m0) and np.array_equal(gi, g0) and np.array_equal(s, s0)


def test_tie_rule_matches_scipy(cuda_device):
    """Exact cost ties: scipy's scan order / tie rule decides; the kernel must
    agree with scipy (via the C oracle that is pinned to scipy on tie-heavy inputs)."""
    rng = np.random.default_rng(3)
    P, M, B = 646, 20, 10
    loc = np.zeros((B, P, 4), np.float32)
    conf = np.full((B, P), 0.5, np.float32)
    gt = np.zeros((B, M, 4), np.float32)
    ng = np.zeros(B, np.int32)
    for b in range(B):
        n = int(rng.integers(1, M + 1))
        ng[b] = n
        gt[b, :n] = rng.integers(0, 3, size=(n, 4)).astype(np.float32) * 0.25      # duplicate GT rows
        if b % 3 == 0:
            loc[b] = 0.25                                                            # every prior identical
        elif b % 3 == 1:
            loc[b] = rng.integers(0, 3, size=(P, 4)).astype(np.float32) * 0.25      # heavy duplication
            conf[b] = rng.choice([0.25, 0.5, 0.75], size=P).astype(np.float32)
        else:
            loc[b] = rng.integers(0, 2, size=(P, 1)).astype(np.float32) * 0.5
    m0, s0, g0 = c_oracle.compute_assignments(loc.reshape(-1, 4), conf.reshape(-1), gt, ng, B, 8.0)
    # the C oracle itself equals scipy on these (checked here once more, end to end)
    m1, s1, g1 = np_oracle.compute_assignments(loc.reshape(-1, 4), conf.reshape(-1).copy(), gt, ng, np.int32(B),
                                               np.float32(8.0), return_indices=True)
    assert np.array_equal(m0, m1) and np.array_equal(g0, g1)
    for warps, cols, generic in [(0, 0, False)] + VARIANTS:
        m, s, gi = _gpu_assign(loc.reshape(-1, 4), conf.reshape(-1), gt, ng, B, 8.0, warps=warps, cols=cols,
                               generic=generic)
        assert np.array_equal(m, m0), (warps, cols, generic)
        assert np.array_equal(gi, g0), (warps, cols, generic)
        assert np.array_equal(s, s0), (warps, cols, generic)


def test_errors_like_scipy(cuda_device):
    d = synth.make_train_inputs(K=5, B=2, M=20, dist="full", seed=5)
    loc, conf = boundary_inputs(d)
    bad = loc.copy()
    bad[700, 2] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        loss.compute_assignments(dev(bad), dev(conf), dev(d["gt"]), dev(d["num_gt"]), 2, 1000.0)
    with pytest.raises(ValueError):
        np_oracle.compute_assignments(bad, conf.copy(), d["gt"], d["num_gt"], np.int32(2), np.float32(1000.0))
    zero = np.zeros_like(conf)         # log(0) = -inf => every cost +inf => infeasible
    with pytest.raises(ValueError, match="infeasible"):
        loss.compute_assignments(dev(loc), dev(zero), dev(d["gt"]), dev(d["num_gt"]), 2, 1000.0)
    # the workspace must be reusable after a failed batch
    m, s = loss.compute_assignments(dev(loc), dev(conf), dev(d["gt"]), dev(d["num_gt"]), 2, 1000.0)
    m0, s0, _ = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], 2, 1000.0)
    assert np.array_equal(m.cpu().numpy(), m0) and np.array_equal(s.cpu().numpy(), s0)


def test_too_many_priors_is_an_argument_error(cuda_device):
    from multibox_b200 import _lib
    P, B = 70000, 1          # beyond both kernel families: reported, never silently mis-solved
    with pytest.raises(_lib.MultiboxLibraryError, match="shared memory"):
        loss.compute_assignments(torch.zeros(B * P, 4, device="cuda"), torch.full((B * P,), 0.5, device="cuda"),
                                 torch.zeros(B, 2, 4, device="cuda"), torch.ones(B, dtype=torch.int32, device="cuda"),
                                 B, 1.0)


def test_empty_and_full_batches(cuda_device):
    d = synth.make_train_inputs(K=5, B=4, M=20, dist="uniform", seed=8)
    loc, conf = boundary_inputs(d)
    ng = np.zeros(4, np.int32)
    m, s = loss.compute_assignments(dev(loc), dev(conf), dev(d["gt"]), dev(ng), 4, 1000.0)
    assert m.sum().item() == 0 and tuple(s.shape) == (0, 4)


def test_full_size_properties(cuda_device):
    """BASELINE configs[4]-shaped images (K=11, P=1420, M=200), 512 of them:
    size-independent properties on all, exact comparison on a 24-image sample."""
    d = synth.make_train_inputs(K=11, B=512, M=200, dist="uniform", seed=1005)
    loc, conf = boundary_inputs(d)
    B, P, M = 512, d["P"], 200
    m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], B, 1000.0)
    m2, gi2 = m.reshape(B, P), gi.reshape(B, P)
    assert np.array_equal(m2.sum(1), d["num_gt"])                   # every GT matched exactly once
    assert s.shape[0] == int(d["num_gt"].sum())
    off = 0
    for b in range(B):
        n = int(d["num_gt"][b])
        idx = gi2[b][m2[b] == 1]
        assert np.array_equal(np.sort(idx), np.arange(n))           # a permutation of the GT rows
        assert np.array_equal(s[off:off + n], d["gt"][b][idx])      # stacked rows in ascending prior order
        off += n
    assert (gi2[m2 == 0] == -1).all()
    sample = np.arange(0, B, 22)[:24]
    for b in sample:
        sl = slice(b * P, (b + 1) * P)
        m0, s0, g0 = c_oracle.compute_assignments(loc[sl], conf[sl], d["gt"][b:b + 1], d["num_gt"][b:b + 1], 1, 1000.0)
        assert np.array_equal(m2[b], m0) and np.array_equal(gi2[b], g0)


@pytest.mark.parametrize("warps", [0, 4, 8, 16, 1])
def test_large_batch_stress(cuda_device, warps):
    """Many images per persistent CTA (B=4096, K=5): every image checked against the C oracle
    (fast port), twice, to flush out intra-CTA races."""
    d = synth.make_train_inputs(K=5, B=4096, M=20, dist="uniform", seed=99)
    loc, conf = boundary_inputs(d)
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], 4096, d["alpha"])
    for _ in range(2):
        m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], 4096, d["alpha"], warps=warps)
        assert np.array_equal(m, m0)
        assert np.array_equal(gi, g0)
        assert np.array_equal(s, s0)


def test_dynamic_schedule_wide_gt_range(cuda_device):
    """MAX_NUM_BBOXES >= 256 (the order kernel buckets counts by n >> shift) with more images than
    resident CTAs: every image against the C oracle."""
    B, M = 700, 300
    d = synth.make_train_inputs(K=5, B=B, M=M, dist="uniform", seed=123)
    assert d["num_gt"].max() > 256
    loc, conf = boundary_inputs(d)
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
    m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
    assert np.array_equal(m, m0) and np.array_equal(gi, g0) and np.array_equal(s, s0)


def test_single_image_and_tiny_batches(cuda_device):
    """B = 1, 2, 3 (fewer images than any scheduling granularity), with and without GT."""
    for B in (1, 2, 3):
        d = synth.make_train_inputs(K=5, B=B, M=20, dist="full" if B != 2 else "uniform", seed=50 + B)
        if B == 1:
            d["num_gt"][:] = 0
        loc, conf = boundary_inputs(d)
        m0, s0, g0 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
        m, s, gi = _gpu_assign(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
        assert np.array_equal(m, m0) and np.array_equal(gi, g0) and np.array_equal(s.reshape(-1, 4), s0.reshape(-1, 4))


# ----------------------------------------------------------------------------- exact-cost pruning
def _assert_same(loc, conf, gt, ng, B, alpha, **kw):
    m0, s0, g0 = c_oracle.compute_assignments(loc, conf, gt, ng, B, alpha)
    m, s, gi = _gpu_assign(loc, conf, gt, ng, B, alpha, **kw)
    assert np.array_equal(m, m0) and np.array_equal(gi, g0) and np.array_equal(s.reshape(-1, 4), s0.reshape(-1, 4))


@pytest.mark.parametrize("warps,cols", [(0, 0), (4, 6), (16, 2)])
def test_pruning_near_duplicate_priors(cuda_device, warps, cols):
    """Many priors whose costs differ by far less than the margin of the cheap bound (the window
    of candidate columns holds tens of priors per row; near-ties and exact ties among them): the
    kernel must still select exactly what the full evaluation selects."""
    rng = np.random.default_rng(17)
    B, P, M = 10, 646, 20
    ng = rng.integers(1, M + 1, size=B).astype(np.int32)
    gt = synth.gt_boxes(rng, ng, M)
    loc = np.zeros((B, P, 4), np.float32)
    conf = np.zeros((B, P), np.float32)
    for b in range(B):
        base = rng.random((8, 4)).astype(np.float32)                         # 8 clusters of ~80 near-identical priors
        pick = rng.integers(0, 8, size=P)
        jitter = (rng.standard_normal((P, 4)) * 10.0 ** rng.uniform(-8, -5)).astype(np.float32)
        loc[b] = base[pick] + jitter
        conf[b] = np.float32(0.3) + (rng.integers(0, 3, size=P) * np.float32(1e-7)).astype(np.float32)
        if b % 3 == 0:
            gt[b, :ng[b]] = loc[b, rng.integers(0, P, size=ng[b])]            # GT exactly on a prior: cost ~ log terms only
    _assert_same(loc.reshape(-1, 4), conf.reshape(-1), gt, ng, B, 1000.0, warps=warps, cols=cols)


@pytest.mark.parametrize("scale,alpha", [(1.0, 1e-3), (1e3, 1e6), (1e-4, 1e9), (1e6, 1.0), (1.0, 0.0), (3e9, 2.0)])
def test_pruning_extreme_scales(cuda_device, scale, alpha):
    """Coordinates / alpha far from the unit square: the margins grow with the magnitudes (up to
    'nothing is pruned'); results stay those of the full evaluation."""
    rng = np.random.default_rng(int(np.log10(scale) * 7 + 100))
    B, P, M = 6, 646, 20
    ng = rng.integers(0, M + 1, size=B).astype(np.int32)
    ng[0] = M
    gt = (synth.gt_boxes(rng, ng, M) * np.float32(scale)).astype(np.float32)
    loc = (rng.random((B * P, 4)) * scale).astype(np.float32)
    conf = rng.uniform(0.01, 0.99, size=B * P).astype(np.float32)
    _assert_same(loc, conf, gt, ng, B, alpha)


def test_invalid_entries_far_from_every_gt_still_raise(cuda_device):
    """scipy validates the WHOLE cost matrix: a NaN / -inf entry raises even when it belongs to a
    prior no GT box would ever consider (a column the cheap bound would prune)."""
    d = synth.make_train_inputs(K=5, B=3, M=20, dist="full", seed=6)
    loc, conf = boundary_inputs(d)
    P = d["P"]
    # the prior farthest from every GT box of image 1
    dist = ((loc[P:2 * P, None, :] - d["gt"][1][None, :, :]) ** 2).sum(-1).min(1)
    far = P + int(np.argmax(dist))
    for bad_value in (-0.5, np.inf, np.nan):          # log(-0.5) = NaN; log(inf) = inf -> cost -inf; NaN
        bad = conf.copy()
        bad[far] = bad_value
        with pytest.raises(ValueError, match="invalid numeric"):
            loss.compute_assignments(dev(loc), dev(bad), dev(d["gt"]), dev(d["num_gt"]), 3, 1000.0)
        with np.errstate(all="ignore"), pytest.raises(ValueError):
            c_oracle.compute_assignments(loc, bad, d["gt"], d["num_gt"], 3, 1000.0)
    badl = loc.copy()
    badl[far, 3] = -np.inf                            # an infinite coordinate: cost +inf, legal, never selected
    _assert_same(badl, conf, d["gt"], d["num_gt"], 3, 1000.0)
    badg = d["gt"].copy()
    badg[2, 5, 1] = np.inf                            # an infinite GT box: its whole row costs +inf -> infeasible
    with pytest.raises(ValueError, match="infeasible"):
        loss.compute_assignments(dev(loc), dev(conf), dev(badg), dev(d["num_gt"]), 3, 1000.0)
    with np.errstate(all="ignore"), pytest.raises(ValueError, match="infeasible"):
        c_oracle.compute_assignments(loc, conf, badg, d["num_gt"], 3, 1000.0)
    # zero confidence far away: +inf cost, legal, never selected
    okc = conf.copy()
    okc[far] = 0.0
    _assert_same(loc, okc, d["gt"], d["num_gt"], 3, 1000.0)


def test_full_size_config3_every_image(cuda_device):
    """BASELINE configs[3] as stated (K=7, P=904, B=1024, MAX_NUM_BBOXES=100, COCO-person-like GT counts):
    every image against the C oracle; dynamic heavy-first scheduling is active at this size."""
    d = synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg4"])
    assert d["B"] == 1024 and d["P"] == 904 and d["M"] == 100
    loc, conf = boundary_inputs(d)
    _assert_same(loc, conf, d["gt"], d["num_gt"], d["B"], d["alpha"])


def test_full_size_config4_shard_every_image(cuda_device):
    """A BASELINE configs[4] per-GPU shard (K=11, P=1420, MAX_NUM_BBOXES=200, n ~ U{0..200}, 1024 images =
    8192 / 8 GPUs): every image against the C oracle (the first shard and, sampled, the other seven)."""
    d = synth.make_train_inputs(K=11, B=1024, M=200, dist="uniform", seed=1005)
    loc, conf = boundary_inputs(d)
    _assert_same(loc, conf, d["gt"], d["num_gt"], 1024, 1000.0)


def test_nplog_exhaustive_on_gpu(cuda_device):
    """Every positive finite float32 (2^31 - 2^23 values): the kernels' log equals the C restatement of
    numpy's float32 log bit for bit (the C twin equals np.log on all of them: gen_golden.py --exhaustive)."""
    step = 1 << 24
    for first in range(0, 0x7f800000, step):
        bits = np.arange(max(first, 1), min(first + step, 0x7f800000), dtype=np.uint32)
        x = bits.view(np.float32)
        got = gpu_nplog(x).view(np.uint32)
        want = c_oracle.nplog(x).view(np.uint32)
        assert np.array_equal(got, want), hex(first)


def test_fast_log_within_the_bound_on_every_float(cuda_device):
    """The cheap cost bound is fed the hardware's fast log of the confidence; its error analysis
    (mbx_bound.h) assumes |__logf(x) - numpy log(x)| <= 2^-21 + 2^-19 |__logf(x)|.  Checked here on EVERY
    float32 bit pattern (zeros, denormals, negatives, infinities and NaNs included)."""
    from multibox_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for first in (0x00000000, 0x40000000, 0x80000000, 0xC0000000):
        _lib.check(lib.mbx_debug_fastlog_violations(first, 0x40000000, bad.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "fast log check")
    assert bad.item() == 0
