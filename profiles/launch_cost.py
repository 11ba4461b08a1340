"""CPU-side cost of one step launch vs the GPU-side step time (GPU box): python profiles/launch_cost.py"""
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
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
for pdl in (False, True):
    step = loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], pdl=pdl)
    launch = step.prepare(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]))
    for _ in range(20):
        launch()
    torch.cuda.synchronize()
    for K in (50, 200, 800):
        t0 = time.perf_counter()
        for _ in range(K):
            launch()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("pdl=%d K=%4d: CPU enqueue %.2f us/launch, until drained %.2f us/step" %
              (pdl, K, 1e6 * (t1 - t0) / K, 1e6 * (t2 - t0) / K))
# what the pieces of the closure cost on the CPU
K = 20000
t0 = time.perf_counter()
for _ in range(K):
    torch.cuda.current_stream(step.device).cuda_stream
t1 = time.perf_counter()
for _ in range(K):
    torch._C._cuda_getCurrentRawStream(0)
t2 = time.perf_counter()
for _ in range(K):
    torch.cuda.current_device()
t3 = time.perf_counter()
print("current_stream().cuda_stream %.2f us, _cuda_getCurrentRawStream %.2f us, current_device %.2f us" %
      (1e6 * (t1 - t0) / K, 1e6 * (t2 - t1) / K, 1e6 * (t3 - t2) / K))
