"""CPU tests: the oracle (numpy + C restatements) against the golden vectors that
oracle/gen_golden.py wrote from the reference's own code, and against the
properties the reference's model_tests.py pins."""
import hashlib
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from multibox_b200 import synth
from oracle import c_oracle, np_oracle, ref_slices


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _boundary(d):
    B = d["B"]
    loc = d["locations"].reshape(-1, 4) + np.tile(d["priors"], (B, 1))
    conf = d["confidences"].reshape(-1) + np.float32(1e-10)
    return loc, conf


# ---------------------------------------------------------------- priors
@pytest.mark.parametrize("K", [5, 7, 11])
def test_priors_oracle_matches_reference_golden(golden_dir, K):
    g = np.load(os.path.join(golden_dir, "priors.npz"))
    ours = np.array(np_oracle.generate_priors(g["ratios%d" % K].tolist()), dtype=np.float64)
    assert ours.shape == (129 * K + 1, 4)
    assert np.array_equal(ours, g["K%d" % K])


def test_priors_count_646():
    # the only known-answer value in the reference: model_tests.py:15
    assert len(np_oracle.generate_priors([1, 2, 3, 1 / 2., 1 / 3.])) == 646


# ---------------------------------------------------------------- matching
def test_match_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "match_small.npz"))
    B = g["locations"].shape[0]
    d = dict(B=B, locations=g["locations"], confidences=g["confidences"], priors=g["priors"])
    loc, conf = _boundary(d)
    for impl in ("np", "c"):
        if impl == "np":
            m, s, gi = np_oracle.compute_assignments(loc, conf.copy(), g["gt"], g["num_gt"], np.int32(B),
                                                     np.float32(g["alpha"]), return_indices=True)
        else:
            m, s, gi = c_oracle.compute_assignments(loc, conf, g["gt"], g["num_gt"], B, float(g["alpha"]))
        assert np.array_equal(m, g["mask"]), impl
        assert np.array_equal(s, g["stacked_gt"]), impl
        assert np.array_equal(gi, g["matched_gt_idx"]), impl


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_match_config_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "match_%s.npz" % name))
    d = synth.make_train_inputs(**synth.TRAIN_CONFIGS[name])
    assert _sha(d["priors"], d["locations"], d["confidences"], d["gt"], d["num_gt"]) == str(g["inputs_sha256"]), \
        "synthetic input generator drifted from the one the golden file was made with"
    loc, conf = _boundary(d)
    m, s, gi = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], d["B"], d["alpha"])
    assert np.array_equal(np.nonzero(m)[0], g["matched_flat_idx"])
    assert np.array_equal(gi[m == 1], g["matched_gt_idx"])
    assert np.array_equal(s, g["stacked_gt"])
    m2, s2 = np_oracle.compute_assignments(loc, conf.copy(), d["gt"], d["num_gt"], np.int32(d["B"]),
                                           np.float32(d["alpha"]))
    assert np.array_equal(m, m2) and np.array_equal(s, s2)


@pytest.mark.skipif(not ref_slices.available(), reason="reference checkout not present")
def test_restatement_equals_reference_slices_fresh_seed():
    ref = ref_slices.load()
    d = synth.make_train_inputs(K=5, B=8, M=20, seed=4242, alpha=1.0, edge_cases=True)   # alpha=1: model_tests.py:98
    loc, conf = _boundary(d)
    m0, s0 = ref["compute_assignments"](loc, conf.copy(), d["gt"], d["num_gt"], np.int32(8), np.float32(1.0))
    m1, s1, _ = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], 8, 1.0)
    assert np.array_equal(m0, m1) and np.array_equal(s0, s1)


# ---------------------------------------------------------------- third-party arithmetic
def test_c_lsap_equals_scipy_on_ties():
    rng = np.random.default_rng(5)
    for trial in range(600):
        nr, nc = rng.integers(1, 25, size=2)
        C = rng.integers(0, rng.integers(1, 5) + 1, size=(nr, nc)).astype(np.float64)
        if trial % 5 == 0:
            C[:] = 1.0     # constant matrix: scipy returns the identity (its reverse-filled scan order)
        a, b = linear_sum_assignment(C)
        a2, b2 = c_oracle.lsap(C)
        assert np.array_equal(a, a2) and np.array_equal(b, b2), (trial, C)


def test_c_lsap_errors_like_scipy():
    C = np.ones((4, 3))
    C[1, 1] = np.nan
    with pytest.raises(ValueError):
        linear_sum_assignment(C)
    with pytest.raises(ValueError):
        c_oracle.lsap(C)
    C = np.full((3, 3), np.inf)
    with pytest.raises(ValueError):
        linear_sum_assignment(C)
    with pytest.raises(ValueError):
        c_oracle.lsap(C)


def test_c_nplog_is_bitwise_numpy_log():
    # sampled here (every 997th float32); oracle/gen_golden.py --exhaustive walks all of them
    bits = np.arange(1, 0x7f800000, 997, dtype=np.uint32)
    x = bits.view(np.float32)
    assert np.array_equal(np.log(x).view(np.uint32), c_oracle.nplog(x).view(np.uint32))
    rng = np.random.default_rng(0)
    c = rng.uniform(0, 1, size=200000).astype(np.float32)
    assert np.array_equal(np.log(c).view(np.uint32), c_oracle.nplog(c).view(np.uint32))


def test_c_cost_matrix_is_bitwise_numpy():
    d = synth.make_train_inputs(K=5, B=2, M=20, seed=3, edge_cases=False)
    loc, conf = _boundary(d)
    P = d["P"]
    lc, l1 = np_oracle.log_terms(conf[:P].copy())
    gt = d["gt"][0][:max(1, d["num_gt"][0])]
    Cn = np_oracle.cost_matrix(loc[:P], lc, l1, gt, np.float32(1000.0))
    Cc = c_oracle.cost_matrix(loc[:P], conf[:P], gt, 1000.0)
    assert np.array_equal(Cn.view(np.uint64), Cc.view(np.uint64))


# ---------------------------------------------------------------- loss properties (model_tests.py)
def _loss(locs, confs, gt, ng, priors, alpha):
    return np_oracle.add_loss(locs, confs, gt, ng, priors, alpha)


def test_single_bounding_box_properties():
    # model_tests.py:104-156: one GT box, random priors, alpha = 1
    rng = np.random.default_rng(0)
    priors = rng.uniform(size=(646, 4)).astype(np.float32)
    locs = rng.normal(0, 0.1, size=(1, 646, 4)).astype(np.float32)
    confs = rng.uniform(0.01, 0.99, size=(1, 646, 1)).astype(np.float32)
    gt = np.zeros((1, 5, 4), np.float32)
    gt[0, 0] = [0.1, 0.1, 0.9, 0.9]
    out = _loss(locs, confs, gt, np.array([1], np.int32), priors, 1.0)
    assert out["location_loss"] > 0 and out["confidence_loss"] > 0
    assert out["mask"].sum() == 1


def test_no_gt_bounding_box_properties():
    # model_tests.py:158-209: zero GT => location loss exactly 0, confidence loss > 0
    rng = np.random.default_rng(1)
    priors = rng.uniform(size=(646, 4)).astype(np.float32)
    locs = rng.normal(0, 0.1, size=(1, 646, 4)).astype(np.float32)
    confs = rng.uniform(0.01, 0.99, size=(1, 646, 1)).astype(np.float32)
    out = _loss(locs, confs, np.zeros((1, 5, 4), np.float32), np.array([0], np.int32), priors, 1.0)
    assert out["location_loss"] == 0.0 and out["confidence_loss"] > 0
    assert out["mask"].sum() == 0 and out["stacked_gt"].shape == (0, 4)


def test_box_with_no_box_properties():
    # model_tests.py:211-263: batch of two, counts [1, 0]
    rng = np.random.default_rng(2)
    priors = rng.uniform(size=(646, 4)).astype(np.float32)
    locs = rng.normal(0, 0.1, size=(2, 646, 4)).astype(np.float32)
    confs = rng.uniform(0.01, 0.99, size=(2, 646, 1)).astype(np.float32)
    gt = np.zeros((2, 5, 4), np.float32)
    gt[0, 0] = [0.1, 0.1, 0.9, 0.9]
    out = _loss(locs, confs, gt, np.array([1, 0], np.int32), priors, 1.0)
    assert out["location_loss"] > 0 and out["confidence_loss"] > 0
    assert out["mask"][:646].sum() == 1 and out["mask"][646:].sum() == 0


def test_loss_gradients_finite_difference():
    d = synth.make_train_inputs(K=5, B=2, M=20, seed=9)
    out = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 10.0)
    assign = (out["mask"], out["stacked_gt"])
    rng = np.random.default_rng(0)
    for _ in range(6):
        b, p, k = rng.integers(0, 2), rng.integers(0, 646), rng.integers(0, 4)
        for which, eps in (("locations", 1e-2), ("confidences", 1e-3)):
            x = d[which].astype(np.float64).copy()
            idx = (b, p, k) if which == "locations" else (b, p, 0)
            xp, xm = x.copy(), x.copy()
            xp[idx] += eps
            xm[idx] -= eps
            args = lambda t: (t.astype(np.float32), d["confidences"]) if which == "locations" \
                else (d["locations"], t.astype(np.float32))
            fp = np_oracle.add_loss(*args(xp), d["gt"], d["num_gt"], d["priors"], 10.0, assignments=assign)
            fm = np_oracle.add_loss(*args(xm), d["gt"], d["num_gt"], d["priors"], 10.0, assignments=assign)
            num = ((fp["location_loss_f64"] + fp["confidence_loss_f64"]) -
                   (fm["location_loss_f64"] + fm["confidence_loss_f64"])) / (2 * eps)
            ana = out["d_locations"][idx] if which == "locations" else out["d_confidences"][idx]
            assert abs(num - ana) <= 2e-2 * max(1.0, abs(ana)), (which, idx, num, ana)


# ---------------------------------------------------------------- detect
def _detect_rows(post, image_ids):
    ids, boxes, scores = [], [], []
    for b, m in enumerate(post):
        for k in range(m["boxes"].shape[0]):
            ids.append(int(image_ids[b]))
            boxes.append(m["boxes"][k])
            scores.append(float(m["scores"][k]))
    return np.array(ids), np.array(boxes).reshape(-1, 4), np.array(scores)


def test_detect_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "detect_small.npz"))
    post = np_oracle.postprocess(g["locations"], g["confidences"], g["priors"], g["restrictions"],
                                 g["max_to_keep"], g["offsets"], g["patch_dims"], g["image_dims"],
                                 g["is_flipped"])
    ids, boxes, scores = _detect_rows(post, g["image_ids"])
    assert np.array_equal(ids, g["out_image_id"])
    assert np.array_equal(boxes, g["out_bbox"])
    assert np.array_equal(scores, g["out_score"])
    assert np.array_equal(np.array([m["boxes"].shape[0] for m in post]), g["out_count"])
    assert np.array_equal(np.concatenate([m["prior_idx"] for m in post]), g["out_prior_idx"])


def test_nms_spec_matches_torchvision_cpu():
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(0)
    for trial in range(40):
        k = int(rng.integers(1, 200))
        c = rng.uniform(size=(k, 2, 2))
        boxes = np.concatenate([c.min(1), c.max(1)], 1).astype(np.float32)
        if trial % 4 == 0:
            boxes[k // 2:] = boxes[:k - k // 2]          # exact duplicates
        scores = np.sort(rng.uniform(size=k).astype(np.float32))[::-1].copy()
        ours = np_oracle.greedy_nms(boxes, 0.5)
        theirs = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
        # with strictly descending (or stably ordered) scores both keep the same positions
        assert np.array_equal(np.sort(ours), np.sort(theirs)), trial


def test_eval_topk_shape():
    d = synth.make_detect_inputs(K=5, B=2, keep=100, seed=3)
    rows = np_oracle.eval_topk(d["locations"], d["confidences"], d["priors"], 299, d["image_ids"], k=100)
    assert len(rows) == 200 and len(rows[0]) == 7 and rows[0][6] == 1
    sc = [r[5] for r in rows[:100]]
    assert all(sc[i] >= sc[i + 1] for i in range(99))


def test_eval_small_golden(golden_dir):
    """eval_small.npz was written by executing reference eval.py:142-175 VERBATIM (oracle/ref_slices.py
    eval_loop_body); the numpy restatement must reproduce its rows bit for bit -- and, where the
    reference checkout is present, so must a fresh execution of the slice."""
    import os
    g = np.load(os.path.join(golden_dir, "eval_small.npz"))
    rows = np_oracle.eval_topk(g["locations"], g["confidences"], g["priors"], 299, g["image_ids"], k=100)
    assert np.array_equal(np.array(rows, dtype=np.float64), g["rows"])
    from oracle import ref_slices
    if ref_slices.available():
        fresh = ref_slices.eval_loop_body(g["locations"], g["confidences"], g["priors"], 299, g["image_ids"])
        a = np.array([[float(np.asarray(v).reshape(-1)[0]) for v in r] for r in fresh], dtype=np.float64)
        assert np.array_equal(a, g["rows"])


def test_loss_graph_vs_torch_autograd():
    """Independent check of the oracle's restatement of the TF graph part (loss.py:67-74,88-101) and
    of the hand-derived gradients: the same statements written with torch CPU
    ops -- index_select for tf.dynamic_partition, the four sentinel rows of loss.py:94-97 kept,
    sum(t**2)/2 for tf.nn.l2_loss -- and differentiated by autograd, as TF's autodiff would."""
    import torch
    d = synth.make_train_inputs(K=5, B=5, M=20, seed=77, edge_cases=True)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], d["alpha"])
    B, P = d["B"], d["P"]
    loc_in = torch.from_numpy(d["locations"]).requires_grad_(True)
    conf_in = torch.from_numpy(d["confidences"]).requires_grad_(True)
    priors = torch.from_numpy(d["priors"])
    locations = loc_in.reshape(-1, 4) + priors.repeat(B, 1)                      # loss.py:67,70-71
    confidences = conf_in.reshape(-1) + np.float32(np_oracle.SMALL_EPSILON)      # loss.py:68,74
    matching = torch.from_numpy(ref["mask"].astype(np.int64))                    # the py_func's outputs are constants
    stacked = torch.from_numpy(ref["stacked_gt"])
    i0, i1 = torch.nonzero(matching == 0).flatten(), torch.nonzero(matching == 1).flatten()
    unmatched_locations, matched_locations = locations.index_select(0, i0), locations.index_select(0, i1)   # :88
    unmatched_confidences, matched_confidences = confidences.index_select(0, i0), confidences.index_select(0, i1)
    matched_locations = torch.cat([matched_locations, torch.zeros(1, 4)], 0)     # :94-97
    stacked = torch.cat([stacked, torch.zeros(1, 4)], 0)
    matched_confidences = torch.cat([matched_confidences, torch.ones(1)], 0)
    unmatched_confidences = torch.cat([unmatched_confidences, torch.zeros(1)], 0)
    diff = matched_locations - stacked
    location_loss = d["alpha"] * (diff.double() ** 2).sum() / 2.                 # :100 (fp64 sum: order-free)
    confidence_loss = -1. * torch.log(matched_confidences).double().sum() \
        - torch.log((1. - unmatched_confidences) + np.float32(np_oracle.SMALL_EPSILON)).double().sum()   # :101
    assert unmatched_locations.shape[0] == B * P - int(ref["mask"].sum())
    (location_loss + confidence_loss).backward()
    np.testing.assert_allclose(location_loss.item(), ref["location_loss_f64"], rtol=1e-7)
    np.testing.assert_allclose(confidence_loss.item(), ref["confidence_loss_f64"], rtol=1e-6)
    np.testing.assert_allclose(loc_in.grad.numpy(), ref["d_locations"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(conf_in.grad.numpy(), ref["d_confidences"], rtol=1e-5, atol=0)
