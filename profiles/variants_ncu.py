"""Runs the cfg2 step with several kernel variants (for an ncu launch list)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multibox_b200 import loss, synth  # noqa: E402

d = synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"])
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
args = (dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]))
pri = dev(d["priors"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for warps, cluster in ((8, 1), (8, 2), (8, 4), (16, 1), (16, 2)):
    out = {}
    for _ in range(3):
        flush.fill_(1)
        loss.match_loss_raw(args[0], args[1], args[2], args[3], pri, d["alpha"], warps=warps, cluster=cluster, out=out)
    torch.cuda.synchronize()
