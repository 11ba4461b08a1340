"""-m gpu parity tests of the two data formats either side of the hot path (SURVEY.md section 8
f3 / f4): per-head NHWC inputs (reference model.py:295-322 fused away) and ragged ground truth
(reference inputs.py:340-348 padding removed).  Both must be BIT-IDENTICAL to the dense / padded
entry points on the equivalent inputs, which are themselves checked against the oracle."""
import numpy as np
import pytest
import torch

from multibox_b200 import detect, loss, synth
from oracle import np_oracle
from gpu_util import dev

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("K,B,M,dist", [(5, 32, 20, "uniform"), (7, 600, 100, "coco_person"), (11, 9, 200, "full")])
def test_ragged_gt_equals_padded(cuda_device, K, B, M, dist):
    d = synth.make_train_inputs(K=K, B=B, M=M, dist=dist, seed=40 + K, edge_cases=True)
    flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
    padded, num = np_oracle.pad_ragged_gt(flat, off, M)        # the reference's own padding step
    assert np.array_equal(padded, d["gt"] * (np.arange(M)[None, :, None] < d["num_gt"][:, None, None]))
    loc, conf = dev(d["locations"]), dev(d["confidences"]).view(B, d["P"])
    a = loss.match_loss_raw(loc, conf, dev(padded), dev(num), dev(d["priors"]), d["alpha"], want_mask=True,
                            want_gt_idx=True, want_stacked=True)
    r = loss.match_loss_ragged_raw(loc, conf, dev(flat), dev(off), dev(d["priors"]), d["alpha"], M, want_mask=True,
                                   want_gt_idx=True, want_stacked=True)
    torch.cuda.synchronize()
    assert a["results"][2].item() == 0 and r["results"][2].item() == 0
    for k in ("mask", "matched_gt_idx", "d_locations", "d_confidences"):
        assert torch.equal(a[k], r[k]), k
    n = int(a["n_stacked"].item())
    assert n == int(r["n_stacked"].item()) == flat.shape[0]
    assert torch.equal(a["stacked_gt"][:n], r["stacked_gt"][:n])
    assert np.array_equal(_bits(a["results"][:8]), _bits(r["results"][:8]))
    ref = np_oracle.add_loss(d["locations"], d["confidences"], padded, num, d["priors"], d["alpha"])
    assert np.array_equal(r["matched_gt_idx"].cpu().numpy(), ref["matched_gt_idx"])
    np.testing.assert_allclose(r["results"][:2].cpu().numpy(), [ref["location_loss"], ref["confidence_loss"]],
                               rtol=RTOL)


def test_ragged_autograd_and_overflow(cuda_device):
    d = synth.make_train_inputs(K=5, B=6, M=20, seed=3)
    flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
    loc = dev(d["locations"]).requires_grad_(True)
    conf = dev(d["confidences"]).requires_grad_(True)
    ll, cl = loss.add_loss_ragged(loc, conf, dev(flat), dev(off), dev(d["priors"]), d["alpha"], 20)
    (ll + cl).backward()
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], d["alpha"])
    np.testing.assert_allclose(loc.grad.cpu().numpy(), ref["d_locations"], rtol=RTOL)
    np.testing.assert_allclose(conf.grad.cpu().numpy(), ref["d_confidences"], rtol=RTOL)
    np.testing.assert_allclose([ll.item(), cl.item()], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    # an image with more rows than the declared capacity is reported, not silently mis-solved
    cap = int(d["num_gt"].max()) - 1
    with pytest.raises(ValueError):
        loss.add_loss_ragged(dev(d["locations"]), dev(d["confidences"]), dev(flat), dev(off), dev(d["priors"]),
                             d["alpha"], cap)


@pytest.mark.parametrize("K,B,M,dist,logits", [(5, 32, 20, "uniform", True), (5, 7, 20, "full", False),
                                               (7, 500, 100, "coco_person", True), (11, 5, 200, "uniform", True)])
def test_head_layout_equals_concatenated(cuda_device, K, B, M, dist, logits):
    d = synth.make_train_inputs(K=K, B=B, M=M, dist=dist, seed=60 + K, edge_cases=True)
    conf_in = d["logits"] if logits else d["confidences"]
    hl, hc = synth.split_heads(d["locations"], conf_in, K)
    # the oracle's layout step (reference model.py:295-320) puts the heads back in prior order
    loc_cat, conf_cat = np_oracle.concat_heads(hl, hc)
    assert np.array_equal(loc_cat, d["locations"]) and np.array_equal(conf_cat, conf_in)
    assert sum(loss.head_priors(K)) == d["P"]
    flags = 1 if logits else 0
    a = loss.match_loss_raw(dev(d["locations"]), dev(conf_in).view(B, d["P"]), dev(d["gt"]), dev(d["num_gt"]),
                            dev(d["priors"]), d["alpha"], flags=flags, want_mask=True, want_gt_idx=True,
                            want_conf_out=logits)
    h = loss.match_loss_heads_raw([dev(t) for t in hl], [dev(t) for t in hc], dev(d["gt"]), dev(d["num_gt"]),
                                  dev(d["priors"]), d["alpha"], flags=flags, want_mask=True, want_gt_idx=True,
                                  want_conf_out=logits)
    torch.cuda.synchronize()
    assert a["results"][2].item() == 0 and h["results"][2].item() == 0
    assert torch.equal(a["mask"], h["mask"]) and torch.equal(a["matched_gt_idx"], h["matched_gt_idx"])
    assert np.array_equal(_bits(a["results"][:8]), _bits(h["results"][:8]))
    if logits:
        assert torch.equal(a["confidences"], h["confidences"])
    # per-head gradients == the dense gradients cut the same way
    gl, gc = synth.split_heads(a["d_locations"].cpu().numpy(), a["d_confidences"].cpu().numpy(), K)
    for k in range(len(hl)):
        assert np.array_equal(h["d_head_locations"][k].cpu().numpy(), gl[k]), k
        assert np.array_equal(h["d_head_confidences"][k].cpu().numpy(), gc[k]), k


def test_head_layout_autograd_vs_oracle(cuda_device):
    K, B = 5, 6
    d = synth.make_train_inputs(K=K, B=B, M=20, seed=17, edge_cases=True)
    hl, hc = synth.split_heads(d["locations"], d["logits"], K)
    tl = [dev(t).requires_grad_(True) for t in hl]
    tc = [dev(t).requires_grad_(True) for t in hc]
    ll, cl = loss.add_loss_from_heads(tl, tc, dev(d["gt"]), dev(d["num_gt"]), dev(d["priors"]), d["alpha"])
    (ll + cl).backward()
    ref = np_oracle.add_loss_from_logits(d["locations"], d["logits"], d["gt"], d["num_gt"], d["priors"], d["alpha"])
    np.testing.assert_allclose([ll.item(), cl.item()], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    gl, gc = synth.split_heads(ref["d_locations"], ref["d_logits"], K)
    for k in range(6):
        np.testing.assert_allclose(tl[k].grad.cpu().numpy(), gl[k], rtol=RTOL, atol=1e-30)
        np.testing.assert_allclose(tc[k].grad.cpu().numpy(), gc[k], rtol=2e-5, atol=1e-12)
    # the plain-torch un-fused path (concat, then the dense entry) gives the same losses
    loc_cat, logit_cat = loss.concat_heads([t.detach() for t in tl], [t.detach() for t in tc])
    l2, c2 = loss.add_loss_from_logits(loc_cat, logit_cat, dev(d["gt"]), dev(d["num_gt"]), dev(d["priors"]), d["alpha"])
    assert l2.item() == ll.item() and c2.item() == cl.item()


def test_head_layout_with_ragged_gt(cuda_device):
    K, B, M = 7, 40, 100
    d = synth.make_train_inputs(K=K, B=B, M=M, dist="coco_person", seed=5)
    hl, hc = synth.split_heads(d["locations"], d["logits"], K)
    flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
    a = loss.match_loss_heads_raw([dev(t) for t in hl], [dev(t) for t in hc], dev(d["gt"]), dev(d["num_gt"]),
                                  dev(d["priors"]), d["alpha"], flags=1, want_gt_idx=True)
    r = loss.match_loss_heads_raw([dev(t) for t in hl], [dev(t) for t in hc], dev(flat), None, dev(d["priors"]),
                                  d["alpha"], flags=1, gt_row_offsets=dev(off), max_num_bboxes=M, want_gt_idx=True)
    torch.cuda.synchronize()
    assert torch.equal(a["matched_gt_idx"], r["matched_gt_idx"])
    assert np.array_equal(_bits(a["results"][:8]), _bits(r["results"][:8]))


def test_head_table_errors(cuda_device):
    d = synth.make_train_inputs(K=5, B=2, M=20, seed=1)
    hl, hc = synth.split_heads(d["locations"], d["logits"], 5)
    with pytest.raises(ValueError):     # one head missing: priors do not add up
        loss.match_loss_heads_raw([dev(t) for t in hl[:-1]], [dev(t) for t in hc[:-1]], dev(d["gt"]),
                                  dev(d["num_gt"]), dev(d["priors"]), d["alpha"], flags=1)


@pytest.mark.parametrize("nms,logits", [(None, True), (0.5, True), (0.5, False)])
def test_detect_head_layout_equals_concatenated(cuda_device, nms, logits):
    q = synth.make_detect_inputs(K=5, B=24, keep=100, seed=33, patches=True)
    conf_in = q["logits"] if logits else q["confidences"]
    hl, hc = synth.split_heads(q["locations"], conf_in, 5)
    kw = dict(restrictions=dev(q["restrictions"]), max_to_keep=dev(q["max_to_keep"]), offsets=dev(q["offsets"]),
              patch_dims=dev(q["patch_dims"]), image_dims=dev(q["image_dims"]), is_flipped=dev(q["is_flipped"]),
              nms_iou=nms, k_max=100, logits=logits)
    a = detect.postprocess(dev(q["locations"]), dev(conf_in), dev(q["priors"]), **kw)
    h = detect.postprocess_heads([dev(t) for t in hl], [dev(t) for t in hc], dev(q["priors"]), **kw)
    torch.cuda.synchronize()
    for k in ("count", "prior_idx", "scores", "boxes", "patch_boxes"):
        assert torch.equal(a[k], h[k]), k
    if not logits:      # and against the oracle's loop body on the concatenated arrays
        post = np_oracle.postprocess(q["locations"], q["confidences"], q["priors"], q["restrictions"],
                                     q["max_to_keep"], q["offsets"], q["patch_dims"], q["image_dims"],
                                     q["is_flipped"], nms_iou=nms)
        cnt = h["count"].cpu().numpy()
        for b, m in enumerate(post):
            assert cnt[b] == m["boxes"].shape[0]
            assert np.array_equal(h["prior_idx"][b, :cnt[b]].cpu().numpy(), m["prior_idx"])
