"""Sweep CTA shapes of the matching kernel on the bench workloads (GPU box):
python profiles/tune_match.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import loss, synth  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def timeit(fn, reps):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def sweep(name, d, variants, reps):
    args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
    pri = dev(d["priors"])
    print("== %s: B=%d P=%d M=%d mean n=%.1f" % (name, d["B"], d["P"], d["M"], d["num_gt"].mean()))
    for v in variants:
        warps, cols, generic = v[:3]
        out = {}

        def fn():
            loss.match_loss_raw(args[0], args[1], args[2], args[3], pri, d["alpha"], flags=4 if generic else 0,
                                warps=warps, cols=cols, out=out)
        try:
            med, mn = timeit(fn, reps)
            print("  warps=%2d cols=%d %-8s median %9.1f us  min %9.1f us  -> %.3g img/s" %
                  (warps, cols, "generic" if generic else "reg", med, mn, d["B"] / (med * 1e-6)))
        except Exception as e:
            print("  warps=%2d cols=%d failed: %s" % (warps, cols, str(e)[:80]))


V_SMALL = [(0, 0, False), (4, 0, False), (8, 0, False), (16, 0, False), (8, 4, False), (4, 8, False)]
sweep("cfg2", synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"]), V_SMALL, 20)
sweep("cfg2 B=256", synth.make_train_inputs(K=5, B=256, M=20, seed=3), V_SMALL, 20)
sweep("cfg2 B=4096", synth.make_train_inputs(K=5, B=4096, M=20, seed=3), V_SMALL, 10)
sweep("cfg4", synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg4"]),
      [(4, 0, True), (0, 0, False), (4, 0, False), (8, 0, False), (16, 0, False), (16, 3, False), (8, 5, False)], 10)
sweep("big K=11 B=1024", synth.make_train_inputs(K=11, B=1024, M=200, dist="uniform", seed=1005),
      [(8, 0, True), (0, 0, False), (8, 0, False), (16, 0, False), (16, 4, False)], 5)
