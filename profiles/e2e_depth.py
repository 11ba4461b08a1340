"""Host-buffer (e2e) throughput of configs[1] against the number of steps kept in flight (GPU box):
python profiles/e2e_depth.py
Each step: packed pinned host inputs read by the kernel over PCIe (zero_copy), 64-byte result block
stored into mapped host memory and polled by the host (host_results), programmatic dependent launch.
depth = steps submitted before the host waits for the oldest one; depth + 1 step objects rotate."""
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
K = 2000
MODES = {"zero_copy+pdl": dict(host_results=True, zero_copy=True, pdl=True),
         "own_stream": dict(host_results=True, own_stream=True)}     # copy engine + kernel, one stream per object
for mode, depth in [(m, k) for m in sys.argv[1:] or list(MODES) for k in (1, 2, 3, 4, 5, 7)]:
    n = depth + 1
    hs = [loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], **MODES[mode]) for _ in range(n)]
    for r, s_ in enumerate(hs):
        np.copyto(s_.h_loc.numpy(), np.roll(d["locations"], r, 0))
        np.copyto(s_.h_conf.numpy(), np.roll(d["confidences"].reshape(B, P), r, 0))
        np.copyto(s_.h_gt.numpy(), np.roll(d["gt"], r, 0))
        np.copyto(s_.h_ng.numpy(), np.roll(d["num_gt"], r, 0))
        s_.step_pinned()
    best = None
    for rep in range(3):
        pend = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            s_ = hs[i % n]
            s_.submit_pinned()
            pend.append(s_)
            if len(pend) > depth:
                pend.pop(0).wait()
        while pend:
            pend.pop(0).wait()
        dt = (time.perf_counter() - t0) / K
        best = dt if best is None else min(best, dt)
    print("%-14s depth %2d: %.2f us per step  (%.2f M images/s)" % (mode, depth, 1e6 * best, B / best / 1e6))
    del hs
