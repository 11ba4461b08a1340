"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz from the reference's OWN code.

Run in the build container (needs the read-only reference checkout):

    python -m oracle.gen_golden [--exhaustive]

For every fixture the reference's functions are executed from line slices
(oracle/ref_slices.py), the numpy restatement (oracle/np_oracle.py) and the C
restatement (oracle/c/mbx_oracle.c) are required to reproduce them bit for bit,
and only then is the fixture written.  The fixtures travel to the GPU box; the
reference checkout does not.

Fixtures
  priors.npz         generate_priors for K=5/7/11 (+ unrestricted K=5), float64
  match_small.npz    full inputs + reference outputs, B=6 (edge cases), K=5, M=20
  match_cfg1.npz     BASELINE configs[0]: seed-regenerated inputs (sha256 pinned)
  match_cfg2.npz     BASELINE configs[1] inputs, reference mask / stacked GT
  detect_small.npz   detect loop body on B=12 mixed patches (whole / crop / flip)
  eval_small.npz     eval loop body (eval.py:142-175 executed verbatim) on B=6 images, top-100 rows
  detect_cfg3_head.npz  first 16 images of configs[2] (no NMS: reference has none)
  patches.npz        extract_patches geometry (detect.py:20-72) for six image / crop shapes
"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle, np_oracle, ref_slices  # noqa: E402
from multibox_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def boundary_inputs(d):
    """What crosses the reference's py_func boundary (loss.py:67-74,81)."""
    B = d["B"]
    loc = d["locations"].reshape(-1, 4) + np.tile(d["priors"], (B, 1))
    conf = d["confidences"].reshape(-1) + np.float32(np_oracle.SMALL_EPSILON)
    assert loc.dtype == np.float32 and conf.dtype == np.float32
    return loc, conf


def ref_match(d):
    ref = ref_slices.load()
    loc, conf = boundary_inputs(d)
    mask, stacked = ref["compute_assignments"](loc, conf.copy(), d["gt"], d["num_gt"],
                                               np.int32(d["B"]), np.float32(d["alpha"]))
    m1, s1, g1 = np_oracle.compute_assignments(loc, conf.copy(), d["gt"], d["num_gt"],
                                               np.int32(d["B"]), np.float32(d["alpha"]),
                                               return_indices=True)
    m2, s2, g2 = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], d["B"], d["alpha"])
    assert mask.dtype == np.int32 and stacked.dtype == np.float32
    assert np.array_equal(mask, m1) and np.array_equal(stacked, s1), "numpy restatement != reference"
    assert np.array_equal(mask, m2) and np.array_equal(stacked, s2), "C restatement != reference"
    assert np.array_equal(g1, g2)
    return mask, stacked, g1


def gen_priors():
    ref = ref_slices.load()
    out = {}
    for K, ratios in synth.ASPECT_RATIOS.items():
        p = np.array(ref["generate_priors"](ratios), dtype=np.float64)
        q = np.array(np_oracle.generate_priors(ratios), dtype=np.float64)
        assert p.shape == (129 * K + 1, 4) and np.array_equal(p, q)
        out["K%d" % K] = p
        out["ratios%d" % K] = np.array(ratios, dtype=np.float64)
    p = np.array(ref["generate_priors"](synth.ASPECT_RATIOS[5], 0.2, 0.9, False), dtype=np.float64)
    q = np.array(np_oracle.generate_priors(synth.ASPECT_RATIOS[5], 0.2, 0.9, False), dtype=np.float64)
    assert np.array_equal(p, q)
    out["K5_unrestricted_0.2_0.9"] = p
    np.savez_compressed(os.path.join(GOLD, "priors.npz"), **out)
    print("priors.npz", {k: v.shape for k, v in out.items()})


def gen_match():
    d = synth.make_train_inputs(K=5, B=6, M=20, seed=7, alpha=1000.0, edge_cases=True)
    mask, stacked, gidx = ref_match(d)
    np.savez_compressed(os.path.join(GOLD, "match_small.npz"), priors=d["priors"],
                        locations=d["locations"], confidences=d["confidences"], gt=d["gt"],
                        num_gt=d["num_gt"], alpha=np.float32(d["alpha"]), mask=mask,
                        stacked_gt=stacked, matched_gt_idx=gidx)
    print("match_small.npz N =", stacked.shape[0])
    for name in ("cfg1", "cfg2"):
        cfg = synth.TRAIN_CONFIGS[name]
        d = synth.make_train_inputs(**cfg)
        mask, stacked, gidx = ref_match(d)
        np.savez_compressed(
            os.path.join(GOLD, "match_%s.npz" % name),
            inputs_sha256=np.array(sha(d["priors"], d["locations"], d["confidences"], d["gt"], d["num_gt"])),
            matched_flat_idx=np.nonzero(mask)[0].astype(np.int32),
            matched_gt_idx=gidx[mask == 1].astype(np.int32), stacked_gt=stacked,
            num_gt=d["num_gt"])
        print("match_%s.npz N = %d" % (name, stacked.shape[0]))


def run_detect(d):
    rows = ref_slices.detect_loop_body(
        d["locations"], d["confidences"], d["priors"], d["offsets"], d["patch_dims"],
        d["is_flipped"], d["restrictions"], d["max_to_keep"], d["image_dims"], d["image_ids"])
    mine = np_oracle.postprocess(d["locations"], d["confidences"], d["priors"], d["restrictions"],
                                 d["max_to_keep"], d["offsets"], d["patch_dims"], d["image_dims"],
                                 d["is_flipped"], nms_iou=None)
    ids = np.array([r["image_id"] for r in rows], dtype=np.int64)
    boxes = np.array([r["bbox"] for r in rows], dtype=np.float64).reshape(-1, 4)
    scores = np.array([r["score"] for r in rows], dtype=np.float64)
    # the restatement must reproduce the reference rows exactly
    k = 0
    for b, m in enumerate(mine):
        c = m["boxes"].shape[0]
        assert np.all(ids[k:k + c] == d["image_ids"][b])
        assert np.array_equal(boxes[k:k + c], m["boxes"]), b
        assert np.array_equal(scores[k:k + c], m["scores"].astype(np.float64)), b
        k += c
    assert k == len(rows)
    counts = np.array([m["boxes"].shape[0] for m in mine], dtype=np.int32)
    prior_idx = np.concatenate([m["prior_idx"] for m in mine]).astype(np.int32)
    return ids, boxes, scores, counts, prior_idx


def gen_detect():
    d = synth.make_detect_inputs(K=5, B=12, keep=50, seed=11, patches=True)
    # force an image whose proposals are all filtered out (detect.py:419-420) and
    # exact confidence ties (tie rule: stable argsort then reversal)
    d["restrictions"][5] = np.array([0.45, 0.45, 0.55, 0.55], np.float32)
    d["confidences"][0, 10:40, 0] = d["confidences"][0, 10, 0]
    d["confidences"][4, :, 0] = np.float32(0.5)
    ids, boxes, scores, counts, prior_idx = run_detect(d)
    assert counts[5] == 0
    keys = ("priors", "locations", "confidences", "restrictions", "max_to_keep", "offsets",
            "patch_dims", "image_dims", "is_flipped", "image_ids")
    np.savez_compressed(os.path.join(GOLD, "detect_small.npz"), **{k: d[k] for k in keys},
                        out_image_id=ids, out_bbox=boxes, out_score=scores, out_count=counts,
                        out_prior_idx=prior_idx)
    print("detect_small.npz rows =", len(ids), "counts =", counts.tolist())

    cfg = dict(synth.DETECT_CONFIGS["cfg3"])
    cfg.pop("nms_iou")
    d = synth.make_detect_inputs(**cfg)
    head = 16
    for k in ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims",
              "image_dims", "is_flipped", "image_ids"):
        d[k] = d[k][:head]
    ids, boxes, scores, counts, prior_idx = run_detect(d)
    np.savez_compressed(os.path.join(GOLD, "detect_cfg3_head.npz"),
                        inputs_sha256=np.array(sha(d["priors"], d["locations"], d["confidences"])),
                        out_bbox=boxes, out_score=scores, out_count=counts, out_prior_idx=prior_idx)
    print("detect_cfg3_head.npz rows =", len(ids))


def gen_eval():
    """eval_small.npz: the reference's own eval loop (eval.py:142-175, executed verbatim) on B=6 images,
    with exact confidence ties; the numpy restatement must reproduce its rows bit for bit."""
    d = synth.make_detect_inputs(K=5, B=6, keep=100, seed=8)
    d["confidences"][1, 5:60, 0] = d["confidences"][1, 5, 0]
    d["confidences"][3, :, 0] = np.float32(0.25)
    rows = ref_slices.eval_loop_body(d["locations"], d["confidences"], d["priors"], 299, d["image_ids"])
    mine = np_oracle.eval_topk(d["locations"], d["confidences"], d["priors"], 299, d["image_ids"], k=100)
    a = np.array([[float(np.asarray(v).reshape(-1)[0]) for v in r] for r in rows], dtype=np.float64)
    b = np.array(mine, dtype=np.float64)
    assert a.shape == b.shape == (600, 7) and np.array_equal(a, b)
    np.savez_compressed(os.path.join(GOLD, "eval_small.npz"), priors=d["priors"], locations=d["locations"],
                        confidences=d["confidences"], image_ids=d["image_ids"], rows=a)
    print("eval_small.npz rows =", a.shape[0])


PATCH_CASES = [(600, 800, (299, 299), (113, 113)), (299, 299, (299, 299), (113, 113)),
               (480, 640, (185, 185), (69, 69)), (200, 500, (299, 299), (113, 113)),
               (525, 412, (185, 185), (69, 69)), (299, 412, (299, 299), (113, 113))]   # == tests/test_patches.py CASES


def gen_patches():
    """patches.npz: the reference's own extract_patches (detect.py:20-72) on all-zero images."""
    from multibox_b200 import patches
    ref = ref_slices.load()["extract_patches"]
    out = {}
    for i, (h, w, dims, strides) in enumerate(PATCH_CASES):
        _, off, restr, n = ref(np.zeros((h, w, 3), np.float32), dims, strides)
        o2, r2, n2 = patches.extract_patches(h, w, dims, strides)
        assert int(n) == int(n2) and np.array_equal(off, o2) and np.array_equal(restr, r2)
        out["off_%d" % i] = off
        out["restr_%d" % i] = restr
    np.savez_compressed(os.path.join(GOLD, "patches.npz"), **out)
    print("patches.npz cases =", len(PATCH_CASES))


def check_nplog(exhaustive):
    step = 1 if exhaustive else 101
    bad = 0
    for lo in range(1, 0x7f800000, 1 << 26):
        x = np.arange(lo, min(lo + (1 << 26), 0x7f800000), step, dtype=np.uint32).view(np.float32)
        bad += int((np.log(x).view(np.uint32) != c_oracle.nplog(x).view(np.uint32)).sum())
    print("np.log vs orc_nplogf over positive finite float32 (step %d): %d mismatches" % (step, bad))
    assert bad == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exhaustive", action="store_true")
    args = ap.parse_args()
    assert ref_slices.available(), "reference checkout not found"
    os.makedirs(GOLD, exist_ok=True)
    check_nplog(args.exhaustive)
    gen_priors()
    gen_match()
    gen_detect()
    gen_eval()
    gen_patches()


if __name__ == "__main__":
    main()
