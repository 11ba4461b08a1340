"""Host-buffer step (configs[1]) under the different transfer modes (GPU box):
python profiles/e2e_modes.py  -- wall-clock microseconds per step_pinned()."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import loss, synth  # noqa: E402

d = synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"])
B, P, M = d["B"], d["P"], d["M"]
modes = [("graph + D2H copy + stream sync", dict(use_graph=True)),
         ("graph + host-mapped results (poll)", dict(use_graph=True, host_results=True)),
         ("graph + host-mapped results + zero-copy inputs", dict(use_graph=True, host_results=True, zero_copy=True)),
         ("no graph + host-mapped results + zero-copy inputs", dict(use_graph=False, host_results=True, zero_copy=True)),
         ("no graph + host-mapped results", dict(use_graph=False, host_results=True))]
for name, kw in modes:
    steps = []
    for r in range(4):
        hs = loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], **kw)
        np.copyto(hs.h_loc.numpy(), np.roll(d["locations"], r, axis=0))
        np.copyto(hs.h_conf.numpy(), np.roll(d["confidences"].reshape(B, P), r, axis=0))
        np.copyto(hs.h_gt.numpy(), np.roll(d["gt"], r, axis=0))
        np.copyto(hs.h_ng.numpy(), np.roll(d["num_gt"], r, axis=0))
        steps.append(hs)
    for i in range(20):
        v = steps[i % 4].step_pinned()
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(300):
            v = steps[i % 4].step_pinned()
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) / 300 * 1e6)
    print("%-52s %6.1f us per step (min of 3: %.1f)  losses %s" % (name, np.median(res), min(res), v))


# detect (configs[2]): packed host inputs -> packed host outputs
from multibox_b200 import detect  # noqa: E402
q = synth.make_detect_inputs(**synth.DETECT_CONFIGS["cfg3"])
names = ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims", "is_flipped")
for name, kw in (("detect: H2D copy + kernel + D2H copy (graph)", dict(zero_copy=False)),
                 ("detect: zero-copy (kernel reads / writes the pinned buffers)", dict(zero_copy=True))):
    steps = []
    for r in range(2):
        ds = detect.DetectStep(q["B"], q["P"], q["keep"], q["priors"], nms_iou=0.5, use_graph=True, **kw)
        ds.fill_host(**{k: np.roll(q[k], r, axis=0) for k in names})
        steps.append(ds)
    for i in range(10):
        steps[i % 2].run_pinned()
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(100):
            out = steps[i % 2].run_pinned()
        res.append((time.perf_counter() - t0) / 100 * 1e6)
    print("%-62s %6.1f us per step (min %.1f)  count[0]=%d" % (name, np.median(res), min(res), int(out["count"][0])))
