"""-m gpu parity tests of detection post-processing (through the C ABI) against
the oracle and the golden rows generated from the reference's own detect.py.
Kept-box indices and float64 converted boxes must be BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from multibox_b200 import detect, synth
from oracle import np_oracle
from gpu_util import dev

pytestmark = pytest.mark.gpu


def _run(d, nms_iou=None, k_max=None, logits=False, warps=0):
    out = detect.postprocess(dev(d["locations"]), dev(d["logits"] if logits else d["confidences"]),
                             dev(d["priors"]),
                             restrictions=dev(d["restrictions"]), max_to_keep=dev(d["max_to_keep"]),
                             offsets=dev(d["offsets"]), patch_dims=dev(d["patch_dims"]),
                             image_dims=dev(d["image_dims"]), is_flipped=dev(d["is_flipped"]),
                             nms_iou=nms_iou, k_max=k_max, logits=logits, warps=warps)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _compare(out, post):
    counts = np.array([m["boxes"].shape[0] for m in post], dtype=np.int32)
    assert np.array_equal(out["count"], counts)
    for b, m in enumerate(post):
        c = counts[b]
        assert np.array_equal(out["prior_idx"][b, :c], m["prior_idx"]), b
        assert np.array_equal(out["boxes"][b, :c], m["boxes"]), b
        assert np.array_equal(out["patch_boxes"][b, :c], m["patch_boxes"]), b
        assert np.array_equal(out["scores"][b, :c], m["scores"]), b
        assert (out["prior_idx"][b, c:] == -1).all() and (out["boxes"][b, c:] == 0).all()


def test_detect_small_golden(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "detect_small.npz"))
    d = {k: g[k] for k in ("priors", "locations", "confidences", "restrictions", "max_to_keep", "offsets",
                           "patch_dims", "image_dims", "is_flipped", "image_ids")}
    out = _run(d)
    assert np.array_equal(out["count"], g["out_count"])
    rows = detect.detection_results({k: torch.from_numpy(v) for k, v in out.items()}, d["image_ids"])
    assert np.array_equal(np.array([r["image_id"] for r in rows]), g["out_image_id"])
    assert np.array_equal(np.array([r["bbox"] for r in rows]).reshape(-1, 4), g["out_bbox"])
    assert np.array_equal(np.array([r["score"] for r in rows]), g["out_score"])
    idx = np.concatenate([out["prior_idx"][b, :c] for b, c in enumerate(out["count"])])
    assert np.array_equal(idx, g["out_prior_idx"])


def test_detect_cfg3_head_golden(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "detect_cfg3_head.npz"))
    cfg = dict(synth.DETECT_CONFIGS["cfg3"])
    cfg.pop("nms_iou")
    d = synth.make_detect_inputs(**cfg)
    for k in ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims",
              "is_flipped", "image_ids"):
        d[k] = d[k][:16]
    out = _run(d)
    assert np.array_equal(out["count"], g["out_count"])
    boxes = np.concatenate([out["boxes"][b, :c] for b, c in enumerate(out["count"])])
    scores = np.concatenate([out["scores"][b, :c] for b, c in enumerate(out["count"])])
    idx = np.concatenate([out["prior_idx"][b, :c] for b, c in enumerate(out["count"])])
    assert np.array_equal(boxes, g["out_bbox"])
    assert np.array_equal(scores.astype(np.float64), g["out_score"])
    assert np.array_equal(idx, g["out_prior_idx"])


@pytest.mark.parametrize("K,B,keep,patches,nms", [(5, 48, 200, False, None), (5, 48, 200, False, 0.5),
                                                  (5, 40, 100, True, 0.5), (11, 24, 200, True, None),
                                                  (11, 24, 200, True, 0.3), (7, 16, 50, True, 0.7)])
def test_detect_vs_oracle(cuda_device, K, B, keep, patches, nms):
    d = synth.make_detect_inputs(K=K, B=B, keep=keep, seed=400 + K + B, patches=patches)
    d["confidences"][1, 5:300:3, 0] = d["confidences"][1, 5, 0]     # confidence ties
    if patches:
        d["restrictions"][3] = np.array([0.49, 0.49, 0.51, 0.51], np.float32)   # nothing survives
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=nms)
    for warps in (0, 4):
        _compare(_run(d, nms_iou=nms, warps=warps), post)


def test_detect_nms_heavy_overlap(cuda_device):
    d = synth.make_detect_inputs(K=5, B=8, keep=200, seed=9)
    d["locations"] *= 0.0                       # boxes == priors: dense, heavily overlapping grid
    d["locations"] += np.random.default_rng(0).normal(0, 0.01, d["locations"].shape).astype(np.float32)
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=0.5)
    assert sum(m["boxes"].shape[0] for m in post) < 8 * 200     # NMS really suppresses here
    _compare(_run(d, nms_iou=0.5), post)


@pytest.mark.parametrize("thr", [0.5, 1.0 / 3.0, 0.25, 0.75, 0.0, 1e-8, 2.0])
def test_detect_nms_threshold_ties(cuda_device, thr):
    """Lattice-valued boxes (zero priors, so decode is the identity): many IoUs equal the
    threshold EXACTLY or sit within an ulp of it, zero-area boxes give 0/0.  The kernel's
    division-free decision must hand every such pair to the exact path (strict >, NaN false)."""
    d = synth.make_detect_inputs(K=5, B=8, keep=200, seed=21)
    rng = np.random.default_rng(22)
    B, P = d["B"], d["P"]
    for b in range(B):
        lat = (4, 8, 16, 3)[b % 4]
        p = rng.integers(0, lat + 1, size=(P, 4)).astype(np.float32) / np.float32(lat)
        box = np.stack([np.minimum(p[:, 0], p[:, 2]), np.minimum(p[:, 1], p[:, 3]),
                        np.maximum(p[:, 0], p[:, 2]), np.maximum(p[:, 1], p[:, 3])], 1)
        if b in (3, 6):   # corners NOT ordered: x2 < x1 boxes, negative areas / denominators (no fix-up in the reference)
            box = p
        if b >= 4:   # a jittered copy: IoUs a few ulp either side of simple fractions
            box = (box * np.float32(1.0 + 3e-7)).astype(np.float32)
        d["locations"][b] = box
    d["priors"] = np.zeros_like(d["priors"])
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=thr)
    for warps in (0, 4, 16):
        _compare(_run(d, nms_iou=thr, warps=warps), post)


def test_detect_logits(cuda_device):
    d = synth.make_detect_inputs(K=5, B=6, keep=100, seed=5)
    out = _run(d, logits=True)
    s = torch.sigmoid(torch.from_numpy(d["logits"])).numpy()
    top = np.sort(s.reshape(6, -1), axis=1)[:, ::-1][:, :100]
    np.testing.assert_allclose(out["scores"], top, rtol=2e-6)


def test_filter_and_convert_single_image_mirrors(cuda_device):
    d = synth.make_detect_inputs(K=5, B=4, keep=100, seed=6, patches=True)
    boxes = np.clip(d["locations"][1] + d["priors"], 0., 1.)
    fb0, fc0 = np_oracle.filter_proposals(boxes, d["confidences"][1], d["restrictions"][1])
    fb, fc = detect.filter_proposals(dev(boxes), dev(d["confidences"][1]), d["restrictions"][1])
    assert np.array_equal(fb.cpu().numpy(), fb0) and np.array_equal(fc.cpu().numpy(), fc0)
    fb1, fc1 = detect.filter_proposals(dev(boxes), dev(d["confidences"][1]))        # default [.1,.1,.9,.9]
    fb2, fc2 = np_oracle.filter_proposals(boxes, d["confidences"][1])
    assert np.array_equal(fb1.cpu().numpy(), fb2) and np.array_equal(fc1.cpu().numpy(), fc2)
    e, _ = detect.filter_proposals(dev(boxes), dev(d["confidences"][1]), [0.5, 0.5, 0.5, 0.5])
    assert tuple(e.shape) == (0,)                                                   # reference: np.array([])
    for flip in (0, 1):
        c0 = np_oracle.convert_proposals(fb0, d["offsets"][1], d["patch_dims"][1], d["image_dims"][1], flip)
        c1 = detect.convert_proposals(dev(fb0), d["offsets"][1], d["patch_dims"][1], d["image_dims"][1], flip)
        assert c1.dtype == torch.float64 and np.array_equal(c1.cpu().numpy(), c0)


def test_eval_topk_rows(cuda_device):
    d = synth.make_detect_inputs(K=5, B=5, keep=100, seed=8)
    rows0 = np_oracle.eval_topk(d["locations"], d["confidences"], d["priors"], 299, d["image_ids"], k=100)
    rows1 = detect.eval_topk(dev(d["locations"]), dev(d["confidences"]), dev(d["priors"]), 299, d["image_ids"], k=100)
    assert len(rows0) == len(rows1) == 500
    assert np.array_equal(np.array(rows0, dtype=np.float64), np.array(rows1, dtype=np.float64))


def test_eval_small_golden(cuda_device, golden_dir):
    """detect.eval_topk against the rows the reference's own eval loop (eval.py:142-175, executed
    verbatim by oracle/gen_golden.py) produced -- exact ties included."""
    g = np.load(os.path.join(golden_dir, "eval_small.npz"))
    rows = detect.eval_topk(dev(g["locations"]), dev(g["confidences"]), dev(g["priors"]), 299, g["image_ids"], k=100)
    assert np.array_equal(np.array(rows, dtype=np.float64), g["rows"])


def test_full_size_config2_with_nms_every_image(cuda_device):
    """BASELINE configs[2] as stated (K=5, B=256, keep 200, NMS IoU 0.5): every image, exact comparison
    with the oracle (kept prior indices, scores, float32 and float64 boxes)."""
    cfg = dict(synth.DETECT_CONFIGS["cfg3"])
    nms = cfg.pop("nms_iou")
    d = synth.make_detect_inputs(**cfg)
    assert d["B"] == 256
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=nms)
    _compare(_run(d, nms_iou=nms), post)


def test_full_size_config4_detect_leg(cuda_device):
    """BASELINE configs[4] detect leg (K=11, P=1420, keep 200, NMS IoU 0.5): 256 images of a per-GPU
    shard, every one compared exactly with the oracle."""
    cfg = dict(synth.DETECT_CONFIGS["cfg5d"])
    nms = cfg.pop("nms_iou")
    cfg["B"] = 256
    d = synth.make_detect_inputs(**cfg)
    assert d["P"] == 1420
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=nms)
    _compare(_run(d, nms_iou=nms), post)


def test_full_size_properties(cuda_device):
    """BASELINE configs[2] at full size (B=256): sortedness, bounds, idempotence."""
    cfg = dict(synth.DETECT_CONFIGS["cfg3"])
    nms = cfg.pop("nms_iou")
    d = synth.make_detect_inputs(**cfg)
    out = _run(d, nms_iou=nms)
    assert (out["count"] <= 200).all() and (out["count"] > 0).all()
    for b in range(0, 256, 17):
        c = out["count"][b]
        sc = out["scores"][b, :c]
        assert (np.diff(sc) <= 0).all()
        assert len(set(out["prior_idx"][b, :c].tolist())) == c
        # kept set is NMS-stable: running NMS on the kept boxes keeps all of them
        assert len(np_oracle.greedy_nms(out["patch_boxes"][b, :c], nms)) == c
    out2 = _run(d, nms_iou=nms)
    assert np.array_equal(out["prior_idx"], out2["prior_idx"])


@pytest.mark.parametrize("use_graph,zero_copy", [(False, False), (True, False), (True, True), (False, True)])
def test_detect_step_host_path(cuda_device, use_graph, zero_copy):
    d = synth.make_detect_inputs(K=5, B=9, keep=60, seed=21, patches=True)      # odd B: section alignment
    names = ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims",
             "is_flipped")
    step = detect.DetectStep(9, d["P"], 60, d["priors"], nms_iou=0.5, use_graph=use_graph, zero_copy=zero_copy)
    post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=0.5)
    for _ in range(3):
        out = step.run_host(**{k: d[k] for k in names})
        for b, m in enumerate(post):
            c = m["boxes"].shape[0]
            assert out["count"][b].item() == c
            assert np.array_equal(out["prior_idx"][b, :c].numpy(), m["prior_idx"])
            assert np.array_equal(out["boxes"][b, :c].numpy(), m["boxes"])
            assert np.array_equal(out["scores"][b, :c].numpy(), m["scores"])


def test_standalone_nms_vs_torchvision(cuda_device):
    """detect.nms (the detect kernel with priors=None) against torchvision's CPU NMS and the oracle's
    greedy loop: same kept indices in the same order.  Ragged counts via -inf padding."""
    torchvision = pytest.importorskip("torchvision")
    rng = np.random.default_rng(5)
    B, n = 6, 300
    ctr = rng.random((B, n, 2)).astype(np.float32) * 0.8 + 0.1
    wh = rng.random((B, n, 2)).astype(np.float32) * 0.25 + 0.02
    boxes = np.clip(np.concatenate([ctr - wh / 2, ctr + wh / 2], -1), 0, 1).astype(np.float32)
    scores = rng.permutation(B * n).reshape(B, n).astype(np.float32) / np.float32(B * n)     # distinct
    counts = np.array([300, 250, 1, 0, 300, 77], np.int32)
    keep, cnt = detect.nms(dev(boxes), dev(scores), 0.4, counts=dev(counts))
    torch.cuda.synchronize()
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        c = counts[b]
        ref = torchvision.ops.nms(torch.from_numpy(boxes[b, :c]), torch.from_numpy(scores[b, :c]), 0.4).numpy()
        order = np.argsort(scores[b, :c], kind="stable")[::-1]
        ref2 = order[np_oracle.greedy_nms(boxes[b, :c][order], 0.4)] if c else np.zeros(0, np.int64)
        assert np.array_equal(ref, ref2)
        assert cnt[b] == len(ref) and np.array_equal(keep[b, :cnt[b]], ref), b
        assert (keep[b, cnt[b]:] == -1).all()


@pytest.mark.parametrize("warps", [0, 4, 16])
def test_detect_nms_many_chunks(cuda_device, warps):
    """max_to_keep = 1024 (32 chunks of 32 survivors: more later chunks than warps), clustered boxes so
    that whole chunks are suppressed, and plain random boxes."""
    d = synth.make_detect_inputs(K=11, B=6, keep=1024, seed=31)
    rng = np.random.default_rng(32)
    centres = rng.random((6, 12, 4)).astype(np.float32)
    for b in range(3):      # images 0-2: every prior snaps to one of 12 cluster boxes (+ jitter)
        c = centres[b][rng.integers(0, 12, d["P"])]
        x1, x2 = np.minimum(c[:, 0], c[:, 2]) * 0.8, np.maximum(c[:, 0], c[:, 2]) * 0.8 + 0.1
        y1, y2 = np.minimum(c[:, 1], c[:, 3]) * 0.8, np.maximum(c[:, 1], c[:, 3]) * 0.8 + 0.1
        box = np.stack([x1, y1, x2, y2], 1) + rng.normal(0, 0.004, (d["P"], 4)).astype(np.float32)
        d["locations"][b] = box.astype(np.float32) - d["priors"]
    for thr in (0.5, 0.8):
        post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                     d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                     d["is_flipped"], nms_iou=thr)
        counts = [m["boxes"].shape[0] for m in post]
        assert min(counts) < 200 and max(counts) > 600
        _compare(_run(d, nms_iou=thr, k_max=1024, warps=warps), post)


@pytest.mark.parametrize("P,keep,B", [(5, 200, 3), (33, 64, 1), (130, 200, 7), (257, 1024, 2)])
def test_detect_small_prior_sets(cuda_device, P, keep, B):
    """Prior sets smaller than max_to_keep / than one sort tile, single-image batches: every valid
    prior survives the top-k, NMS and conversion still match the oracle."""
    rng = np.random.default_rng(P)
    pri = rng.random((P, 4)).astype(np.float32) * 0.5
    pri[:, 2:] = pri[:, :2] + 0.05 + rng.random((P, 2)).astype(np.float32) * 0.4
    d = dict(priors=pri, locations=rng.normal(0, 0.02, (B, P, 4)).astype(np.float32),
             confidences=rng.random((B, P, 1)).astype(np.float32),
             restrictions=np.tile(np.array([0, 0, 1, 1], np.float32), (B, 1)),
             max_to_keep=np.full((B, 1), keep, np.int32), offsets=np.zeros((B, 2), np.int32),
             patch_dims=np.full((B, 2), 300, np.int32), image_dims=np.full((B, 2), 300, np.int32),
             is_flipped=np.zeros((B, 1), np.int32))
    for nms in (None, 0.3):
        post = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                     d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                     d["is_flipped"], nms_iou=nms)
        if nms is None:
            assert all(m["boxes"].shape[0] == min(P, keep) for m in post)
        _compare(_run(d, nms_iou=nms, k_max=keep), post)
