"""Multi-patch detection: patch geometry against the reference's own
extract_patches (detect.py:20-72, executed from a line slice when the checkout is present, and
against the committed golden rows otherwise), the per-image plan of detect.py:189-271, and -- on
the GPU -- the pooled cross-patch merge against its oracle specification."""
import os

import numpy as np
import pytest

from multibox_b200 import patches
from oracle import np_oracle, ref_slices

# the DETECTION section of the reference's config.yaml.example:62-84
EXAMPLE_CFG = dict(USE_ORIGINAL_IMAGE=True, ORIGINAL_IMAGE_MAX_TO_KEEP=200,
                   USE_FLIPPED_ORIGINAL_IMAGE=False, FLIPPED_IMAGE_MAX_TO_KEEP=100,
                   CROPS=[dict(HEIGHT=299, WIDTH=299, HEIGHT_STRIDE=113, WIDTH_STRIDE=113, FLIP=False, MAX_TO_KEEP=50),
                          dict(HEIGHT=185, WIDTH=185, HEIGHT_STRIDE=69, WIDTH_STRIDE=69, FLIP=False, MAX_TO_KEEP=50)])
CASES = [(600, 800, (299, 299), (113, 113)), (299, 299, (299, 299), (113, 113)), (480, 640, (185, 185), (69, 69)),
         (200, 500, (299, 299), (113, 113)), (525, 412, (185, 185), (69, 69)), (299, 412, (299, 299), (113, 113))]


@pytest.mark.skipif(not ref_slices.available(), reason="reference checkout not present")
@pytest.mark.parametrize("h,w,dims,strides", CASES)
def test_extract_patches_equals_reference(h, w, dims, strides):
    ref = ref_slices.load()["extract_patches"]
    _, off, restr, n = ref(np.zeros((h, w, 3), np.float32), dims, strides)
    o2, r2, n2 = patches.extract_patches(h, w, dims, strides)
    assert int(n) == int(n2)
    assert off.dtype == o2.dtype and restr.dtype == r2.dtype
    assert np.array_equal(off, o2) and np.array_equal(restr, r2)


def test_extract_patches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "patches.npz"))
    for i, (h, w, dims, strides) in enumerate(CASES):
        o2, r2, n2 = patches.extract_patches(h, w, dims, strides)
        assert np.array_equal(o2, g["off_%d" % i]) and np.array_equal(r2, g["restr_%d" % i])


def test_patch_plan_follows_the_reference_order():
    plan = patches.patch_plan(600, 800, EXAMPLE_CFG)
    n299 = 3 * 5          # (600-299)//113+1 = 3 rows, (800-299)//113+1 = 5 columns
    n185 = 7 * 9
    n = 1 + n299 + n185
    assert plan["offsets"].shape == (n, 2) and plan["restrictions"].shape == (n, 4)
    # original image first: whole-image restriction, its own size as the patch size (detect.py:204-222)
    assert plan["offsets"][0].tolist() == [0, 0] and plan["patch_dims"][0].tolist() == [600, 800]
    assert plan["restrictions"][0].tolist() == [0., 0., 1., 1.] and plan["max_to_keep"][0, 0] == 200
    assert (plan["patch_dims"][1:1 + n299] == 299).all() and (plan["patch_dims"][1 + n299:] == 185).all()
    assert (plan["max_to_keep"][1:] == 50).all() and (plan["is_flipped"] == 0).all()
    assert (plan["image_dims"] == np.array([600, 800])).all()
    # a crop in the interior is restricted on all four sides, the top-left one only right/bottom
    assert plan["restrictions"][1].tolist() == [0., 0., np.float32(0.9), np.float32(0.9)]
    interior = 1 + 1 * 5 + 1
    assert plan["restrictions"][interior].tolist() == [np.float32(0.1), np.float32(0.1), np.float32(0.9), np.float32(0.9)]
    cfg = dict(EXAMPLE_CFG, USE_FLIPPED_ORIGINAL_IMAGE=True)
    plan2 = patches.patch_plan(600, 800, cfg)
    assert plan2["is_flipped"][:2, 0].tolist() == [0, 1] and plan2["max_to_keep"][1, 0] == 100
    # an image smaller than a crop yields no patch of that size (detect.py:62-67)
    small = patches.patch_plan(150, 150, EXAMPLE_CFG)
    assert small["offsets"].shape[0] == 1
    bp = patches.batch_plan([(600, 800), (150, 150)], EXAMPLE_CFG)
    assert bp["image_index"].tolist() == [0] * n + [1]


@pytest.mark.gpu
@pytest.mark.parametrize("nms", [0.5, None])
def test_merge_patches_vs_oracle(cuda_device, nms):
    import torch
    from multibox_b200 import detect, synth
    from gpu_util import dev
    dims = [(600, 800), (299, 299), (350, 420), (150, 150)]
    cfg = dict(EXAMPLE_CFG, USE_FLIPPED_ORIGINAL_IMAGE=True)
    plan = patches.batch_plan(dims, cfg)
    Bp = plan["offsets"].shape[0]
    q = synth.make_detect_inputs(K=5, B=Bp, keep=200, seed=77)
    # interleave the patches of different images: the merge must not rely on contiguity
    perm = np.random.default_rng(1).permutation(Bp)
    meta = {k: np.ascontiguousarray(plan[k][perm]) for k in plan}
    post = detect.postprocess(dev(q["locations"]), dev(q["confidences"]), dev(q["priors"]),
                              restrictions=dev(meta["restrictions"]), max_to_keep=dev(meta["max_to_keep"]),
                              offsets=dev(meta["offsets"]), patch_dims=dev(meta["patch_dims"]),
                              image_dims=dev(meta["image_dims"]), is_flipped=dev(meta["is_flipped"]),
                              nms_iou=0.5, k_max=200)
    merged = patches.merge_patches(post, meta["image_index"], len(dims), nms_iou=nms, max_detections=300)
    torch.cuda.synchronize()
    ref_post = np_oracle.postprocess(q["locations"], q["confidences"], q["priors"], meta["restrictions"],
                                     meta["max_to_keep"], meta["offsets"], meta["patch_dims"], meta["image_dims"],
                                     meta["is_flipped"], nms_iou=0.5)
    ref = np_oracle.merge_patches(ref_post, meta["image_index"], len(dims), nms_iou=nms, max_detections=300)
    cnt = merged["count"].cpu().numpy()
    assert cnt.tolist() == [m["boxes"].shape[0] for m in ref]
    assert cnt[0] > 200 or nms is not None          # pooled from many patches
    for i, m in enumerate(ref):
        c = cnt[i]
        assert np.array_equal(merged["boxes"][i, :c].cpu().numpy(), m["boxes"]), i      # float64, bit-exact
        assert np.array_equal(merged["scores"][i, :c].cpu().numpy(), m["scores"]), i
        assert np.array_equal(merged["source_patch"][i, :c].cpu().numpy(), m["source_patch"]), i
        assert (merged["source_patch"][i, c:] == -1).all()
