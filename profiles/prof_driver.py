"""Short driver for ncu captures: runs each hot kernel a few times on the bench
workloads.  Usage (on the GPU box, under ncu):  python profiles/prof_driver.py [cfg2|big|detect|all] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import detect, loss, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def flush_l2():
    torch.empty(256 << 20, dtype=torch.uint8, device="cuda").fill_(1)


if which in ("cfg2", "all"):
    d = synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"])
    step = loss.MultiboxLossStep(d["B"], d["P"], d["M"], d["priors"], d["alpha"])
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(reps):
        flush_l2()
        step.step(*args)
    torch.cuda.synchronize()

if which in ("big", "all"):
    d = synth.make_train_inputs(K=11, B=1024, M=200, dist="uniform", seed=1005)
    step = loss.MultiboxLossStep(d["B"], d["P"], d["M"], d["priors"], d["alpha"])
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(reps):
        flush_l2()
        step.step(*args)
    torch.cuda.synchronize()

if which in ("lat148",):
    d = synth.make_train_inputs(K=5, B=148, M=20, dist="full", seed=5)
    step = loss.MultiboxLossStep(d["B"], d["P"], d["M"], d["priors"], d["alpha"], warps=8)
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(reps):
        flush_l2()
        step.step(*args)
    torch.cuda.synchronize()

if which in ("b4096",):
    d = synth.make_train_inputs(K=5, B=4096, M=20, seed=3)
    step = loss.MultiboxLossStep(d["B"], d["P"], d["M"], d["priors"], d["alpha"])
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(reps):
        flush_l2()
        step.step(*args)
    torch.cuda.synchronize()

if which in ("cfg4", "all"):
    d = synth.make_train_inputs(**dict(synth.TRAIN_CONFIGS["cfg4"]))
    step = loss.MultiboxLossStep(d["B"], d["P"], d["M"], d["priors"], d["alpha"])
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(reps):
        flush_l2()
        step.step(*args)
    torch.cuda.synchronize()

if which in ("detect", "all"):
    q = synth.make_detect_inputs(**{k: v for k, v in synth.DETECT_CONFIGS["cfg3"].items()})
    t = {k: dev(q[k]) for k in ("locations", "confidences", "priors", "restrictions", "max_to_keep", "offsets",
                                "patch_dims", "image_dims", "is_flipped")}
    out = {}
    for _ in range(reps):
        flush_l2()
        detect.postprocess(t["locations"], t["confidences"], t["priors"], restrictions=t["restrictions"],
                           max_to_keep=t["max_to_keep"], offsets=t["offsets"], patch_dims=t["patch_dims"],
                           image_dims=t["image_dims"], is_flipped=t["is_flipped"], nms_iou=0.5, k_max=200, out=out)
    torch.cuda.synchronize()
print("prof_driver done", which, reps)
