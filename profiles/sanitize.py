"""compute-sanitizer driver (GPU box): tiny invocations of every kernel family, meant to be run as
  compute-sanitizer --tool {memcheck,racecheck,synccheck} python profiles/sanitize.py
Covers: register-resident matching kernel (static + dynamic scheduling, logits, heads, ragged,
generic kernel, order / scan kernels, detect kernel (NMS on / off, heads,
pooled merge), filter / convert kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import _lib, detect, loss, patches, synth  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


d = synth.make_train_inputs(K=5, B=6, M=20, seed=1, edge_cases=True)
args = (dev(d["locations"]), dev(d["confidences"]).view(6, -1), dev(d["gt"]), dev(d["num_gt"]), dev(d["priors"]))
for kw in (dict(), dict(flags=_lib.FLAG_GENERIC), dict(warps=4), dict(warps=16),
           dict(warps=4, cols=6), dict(flags=_lib.FLAG_LOGITS, want_conf_out=True)):
    out = loss.match_loss_raw(*args, d["alpha"], want_mask=True, want_gt_idx=True, want_stacked=True, **kw)
    torch.cuda.synchronize()
    assert out["results"][2].item() == 0, kw
# dynamic scheduling: more images than resident CTAs (small P keeps it quick)
big = synth.make_train_inputs(K=5, B=700, M=20, seed=2)
out = loss.match_loss_raw(dev(big["locations"]), dev(big["confidences"]).view(700, -1), dev(big["gt"]),
                          dev(big["num_gt"]), dev(big["priors"]), big["alpha"], want_mask=True)
torch.cuda.synchronize()
hl, hc = synth.split_heads(d["locations"], d["logits"], 5)
flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
loss.match_loss_heads_raw([dev(t) for t in hl], [dev(t) for t in hc], dev(flat), None, dev(d["priors"]), d["alpha"],
                          flags=1, gt_row_offsets=dev(off), max_num_bboxes=20, want_mask=True)
loss.match_loss_ragged_raw(args[0], args[1], dev(flat), dev(off), args[4], d["alpha"], 20, want_stacked=True)
torch.cuda.synchronize()

q = synth.make_detect_inputs(K=5, B=6, keep=100, seed=3, patches=True)
kw = dict(restrictions=dev(q["restrictions"]), max_to_keep=dev(q["max_to_keep"]), offsets=dev(q["offsets"]),
          patch_dims=dev(q["patch_dims"]), image_dims=dev(q["image_dims"]), is_flipped=dev(q["is_flipped"]))
for nms in (None, 0.5):
    post = detect.postprocess(dev(q["locations"]), dev(q["confidences"]), dev(q["priors"]), nms_iou=nms, k_max=100, **kw)
hl, hc = synth.split_heads(q["locations"], q["logits"], 5)
detect.postprocess_heads([dev(t) for t in hl], [dev(t) for t in hc], dev(q["priors"]), nms_iou=0.5, k_max=100, **kw)
patches.merge_patches(post, np.array([0, 0, 1, 1, 2, 2]), 3, nms_iou=0.5, max_detections=150)
boxes = np.clip(q["locations"][1] + q["priors"], 0., 1.)
detect.filter_proposals(dev(boxes), dev(q["confidences"][1]), q["restrictions"][1])
detect.convert_proposals(dev(boxes[:50]), q["offsets"][1], q["patch_dims"][1], q["image_dims"][1], 1)
torch.cuda.synchronize()
print("sanitize driver done")
