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

# host-buffer (e2e) path: what the host spends per step
hs = [loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], host_results=True, zero_copy=True, pdl=True) for _ in range(4)]
for r, s_ in enumerate(hs):
    np.copyto(s_.h_loc.numpy(), np.roll(d["locations"], r, 0))
    np.copyto(s_.h_conf.numpy(), np.roll(d["confidences"].reshape(B, P), r, 0))
    np.copyto(s_.h_gt.numpy(), np.roll(d["gt"], r, 0))
    np.copyto(s_.h_ng.numpy(), np.roll(d["num_gt"], r, 0))
    s_.step_pinned()
K = 400
t0 = time.perf_counter()
for i in range(K):
    hs[i % 4].step_pinned()
t1 = time.perf_counter()
pend = []
for i in range(K):
    hs[i % 4].submit_pinned()
    pend.append(hs[i % 4])
    if len(pend) > 1:
        pend.pop(0).wait()
pend.pop(0).wait()
t2 = time.perf_counter()
ts, tw = 0.0, 0.0
for i in range(K):
    a = time.perf_counter()
    hs[i % 4].submit_pinned()
    b = time.perf_counter()
    hs[i % 4].wait()
    c = time.perf_counter()
    ts += b - a
    tw += c - b
t3 = time.perf_counter()
for i in range(K):
    np.copyto(hs[i % 4].h_loc.numpy(), d["locations"])
    np.copyto(hs[i % 4].h_conf.numpy(), d["confidences"].reshape(B, P))
    np.copyto(hs[i % 4].h_gt.numpy(), d["gt"])
    np.copyto(hs[i % 4].h_ng.numpy(), d["num_gt"])
t4 = time.perf_counter()
print("e2e: one in flight %.2f us/step; two in flight %.2f us/step; submit_pinned() alone %.2f us, wait() after it %.2f us; "
      "staging np.copyto of the four arrays %.2f us" %
      (1e6 * (t1 - t0) / K, 1e6 * (t2 - t1) / K, 1e6 * ts / K, 1e6 * tw / K, 1e6 * (t4 - t3) / K))
