"""Times the detect kernel on the bench shapes (GPU box): python profiles/time_detect.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multibox_b200 import detect, synth  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def clustered(q, n_clusters=12, seed=1):
    """Detector-like outputs: every prior snaps to one of a few objects (+ jitter), so that NMS
    suppresses most of the top-k (random-init heads suppress almost nothing)."""
    rng = np.random.default_rng(seed)
    B, P = q["B"], q["P"]
    for b in range(B):
        c = rng.random((n_clusters, 4)).astype(np.float32)[rng.integers(0, n_clusters, P)]
        x1, x2 = np.minimum(c[:, 0], c[:, 2]) * 0.8, np.maximum(c[:, 0], c[:, 2]) * 0.8 + 0.1
        y1, y2 = np.minimum(c[:, 1], c[:, 3]) * 0.8, np.maximum(c[:, 1], c[:, 3]) * 0.8 + 0.1
        box = np.stack([x1, y1, x2, y2], 1) + rng.normal(0, 0.004, (P, 4)).astype(np.float32)
        q["locations"][b] = box.astype(np.float32) - q["priors"]
    return q


for name, kw in (("cfg3 B=256 K=5", dict(K=5, B=256, keep=200, seed=1003)),
                 ("cfg3 B=256 K=5 CLUSTERED boxes", dict(K=5, B=256, keep=200, seed=1003)),
                 ("K=11 B=1024", dict(K=11, B=1024, keep=200, seed=5)),
                 ("K=5 B=4096", dict(K=5, B=4096, keep=200, seed=6)),
                 ("K=5 B=4096 CLUSTERED boxes", dict(K=5, B=4096, keep=200, seed=6))):
    q = synth.make_detect_inputs(**kw)
    if "CLUSTERED" in name:
        q = clustered(q)
    t = {k: dev(q[k]) for k in ("locations", "confidences", "priors", "restrictions", "max_to_keep", "offsets",
                                "patch_dims", "image_dims", "is_flipped")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for nms in (None, 0.5):
        for warps in (8, 16):
            out = {}

            def fn():
                detect.postprocess(t["locations"], t["confidences"], t["priors"], restrictions=t["restrictions"],
                                   max_to_keep=t["max_to_keep"], offsets=t["offsets"], patch_dims=t["patch_dims"],
                                   image_dims=t["image_dims"], is_flipped=t["is_flipped"], nms_iou=nms, k_max=200,
                                   warps=warps, out=out)
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print("%s nms=%s warps=%d: %.1f us -> %.3g img/s" % (name, nms, warps, np.median(ts), q["B"] / np.median(ts) * 1e6))
