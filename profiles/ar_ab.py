"""A/B of the fused all-reduce variants under torchrun (GPU box, N >= 2):
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/ar_ab.py
Back-to-back device-resident configs[1] steps (PDL + deferred all-reduce), max over ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import dist as mdist  # noqa: E402
from multibox_b200 import loss, synth  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
B = 32
sets = []
for sd in range(8):
    dp = synth.make_train_inputs(K=5, B=B * world, M=20, seed=1002 + sd)
    lo, hi = mdist.shard_range(B * world)
    sets.append({k: dp[k][lo:hi] for k in ("locations", "confidences", "gt", "num_gt")})
from multibox_b200 import _lib  # noqa: E402
variants = [("no all-reduce (N=1 kernel)", None, 0), ("deferred lag 4, relay batch 1", (4, 1), 0),
            ("deferred lag 8, relay batch 4", (8, 4), 0), ("deferred lag 12, relay batch 8 (default)", (12, 8), 0),
            ("blocking", (12, 8), 1)]
for name, xp, blocking in variants:
    peer = mdist.PeerAllreduce() if xp is not None else None
    if xp is not None:
        _lib.check(_lib.load().mbx_allreduce_config(xp[0], xp[1]), "mbx_allreduce_config")
    step = loss.MultiboxLossStep(B, dp["P"], 20, dp["priors"], dp["alpha"], peer=peer, deferred_allreduce=not blocking,
                                 pdl=True)
    launches = [step.prepare(dev_(x["locations"]), dev_(x["confidences"]).view(B, -1), dev_(x["gt"]), dev_(x["num_gt"]))
                for x in sets]
    for i in range(50):
        launches[i % 8]()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(1000):
            launches[i % 8]()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, t.item())
    st = int(step.out["results"].cpu()[2].item())
    fb = -1
    if peer is not None:      # steps whose words the relay had not delivered in time (pulled over NVLink instead)
        fb = int(peer._keep[0][28:32].view(torch.int32).item())
    if rank == 0:
        print("%-42s %.2f us per step (status %d, %d of ~%d steps took the pull route)" % (name, best, st, fb, 50 + 8 + 3000))
dist.barrier()
dist.destroy_process_group()
