"""Multi-GPU check of the fused (peer-memory) loss all-reduce, run under torchrun on the GPU box:
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/dist_check.py
Every rank solves its contiguous shard of one global batch; the fused global losses must equal
(a) an NCCL all-reduce of the local fp64 losses and (b) the oracle on the whole batch."""
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
peer = mdist.PeerAllreduce()
B = 8 * world + 3
d = synth.make_train_inputs(K=5, B=B, M=20, seed=4321, edge_cases=True)
lo, hi = mdist.shard_range(B)
sh = mdist.shard_batch({k: d[k] for k in ("locations", "confidences", "gt", "num_gt")}, B)
from oracle import np_oracle
ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], d["alpha"])
ok = True
for deferred in (False, True):
    step = loss.MultiboxLossStep(hi - lo, d["P"], 20, d["priors"], d["alpha"], peer=peer, use_graph=True,
                                 deferred_allreduce=deferred)
    for it in range(7):       # several steps: exercises the slot ring and the sequence counter
        scale = 1.0 + it      # a different global sum every step
        ll, cl = step.step_host(sh["locations"] * np.float32(1.0), sh["confidences"], sh["gt"], sh["num_gt"])
        t64 = step.out["results"][4:8].view(torch.float64).clone()
        dist.all_reduce(t64)
        if not deferred:
            gl, gc = step.global_losses()
            ok &= abs(gl - t64[0].item()) <= 1e-12 * abs(gl) and abs(gc - t64[1].item()) <= 1e-12 * abs(gc)
        elif it >= 1:
            gl, gc = step.global_losses()      # belongs to the previous step (same inputs here)
            ok &= abs(gl - t64[0].item()) <= 1e-12 * abs(gl) and abs(gc - t64[1].item()) <= 1e-12 * abs(gc)
    gl, gc = step.flush() if deferred else step.global_losses()
    ok &= abs(gl - ref["location_loss_f64"]) <= 1e-9 * abs(gl) and abs(gc - ref["confidence_loss_f64"]) <= 1e-9 * abs(gc)
    if rank == 0:
        print("dist_check world=%d deferred=%s: fused global losses %.6f %.6f  oracle %.6f %.6f  -> %s"
              % (world, deferred, gl, gc, ref["location_loss_f64"], ref["confidence_loss_f64"], "OK" if ok else "MISMATCH"))
# a shard larger than the resident CTAs: heavy-first dynamic scheduling + the deferred poster CTA
Bbig = 700 * world
dbig = synth.make_train_inputs(K=7, B=Bbig, M=100, dist="coco_person", seed=99)
lo, hi = mdist.shard_range(Bbig)
shb = mdist.shard_batch({k: dbig[k] for k in ("locations", "confidences", "gt", "num_gt")}, Bbig)
for deferred in (False, True):
    step = loss.MultiboxLossStep(hi - lo, dbig["P"], 100, dbig["priors"], dbig["alpha"], peer=peer, use_graph=True,
                                 deferred_allreduce=deferred, host_results=True)
    for it in range(4):
        step.step_host(shb["locations"], shb["confidences"], shb["gt"], shb["num_gt"])
    t64 = step.out["results"][4:8].view(torch.float64).clone().cuda()    # (the result block lives in pinned host memory)
    dist.all_reduce(t64)
    gl, gc = step.flush() if deferred else step.global_losses()
    okb = abs(gl - t64[0].item()) <= 1e-12 * abs(gl) and abs(gc - t64[1].item()) <= 1e-12 * abs(gc)
    solo = loss.MultiboxLossStep(hi - lo, dbig["P"], 100, dbig["priors"], dbig["alpha"], use_graph=False)
    l1, c1 = solo.step_host(shb["locations"], shb["confidences"], shb["gt"], shb["num_gt"])
    l2 = step.out["results"][0].item()
    okb &= (l1 == l2)
    ok &= okb
    if rank == 0:
        print("dist_check world=%d big shard (%d images/rank) deferred=%s: %.4f %.4f -> %s"
              % (world, hi - lo, deferred, gl, gc, "OK" if okb else "MISMATCH"))
# programmatic dependent launch (MBX_FLAG_PDL) + deferred fused all-reduce: 40 back-to-back device-resident
# steps over three different shards, no host synchronisation in between; the newest step's global sums
# (flush) must equal the NCCL all-reduce of that step's local sums
Bp = 32 * world
sets = []
for sd in (11, 12, 13):
    dp = synth.make_train_inputs(K=5, B=Bp, M=20, seed=sd)
    lo, hi = mdist.shard_range(Bp)
    sets.append(mdist.shard_batch({k: dp[k] for k in ("locations", "confidences", "gt", "num_gt")}, Bp))
step = loss.MultiboxLossStep(hi - lo, dp["P"], 20, dp["priors"], dp["alpha"], peer=peer, deferred_allreduce=True,
                             pdl=True)
dev_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
launches = [step.prepare(dev_(x["locations"]), dev_(x["confidences"]).view(hi - lo, -1), dev_(x["gt"]), dev_(x["num_gt"]))
            for x in sets]
torch.cuda.synchronize()
dist.barrier()
for it in range(40):
    launches[it % 3]()
gl, gc = step.flush()
t64 = step.out["results"][4:8].view(torch.float64).clone()
dist.all_reduce(t64)
okp = abs(gl - t64[0].item()) <= 1e-12 * abs(gl) and abs(gc - t64[1].item()) <= 1e-12 * abs(gc)
ok &= okp
if rank == 0:
    print("dist_check world=%d PDL + deferred, 40 back-to-back steps: fused %.6f %.6f nccl %.6f %.6f -> %s"
          % (world, gl, gc, t64[0].item(), t64[1].item(), "OK" if okp else "MISMATCH"))
# pipelined host-buffer steps (bench.py's e2e mode): rotating step objects on streams of their own, each with its
# OWN exchange state; every object's newest step (flush) must equal the NCCL all-reduce of its local sums
npipe = 4
objs = [loss.MultiboxLossStep(hi - lo, dp["P"], 20, dp["priors"], dp["alpha"], peer=mdist.PeerAllreduce(),
                              deferred_allreduce=True, host_results=True, own_stream=True) for _ in range(npipe)]
pend = []
for it in range(3 * npipe + 2):
    o, x = objs[it % npipe], sets[it % 3]
    if len(pend) == npipe:
        pend.pop(0).wait()
    np.copyto(o.h_loc.numpy(), x["locations"])
    np.copyto(o.h_conf.numpy(), x["confidences"].reshape(hi - lo, -1))
    np.copyto(o.h_gt.numpy(), x["gt"])
    np.copyto(o.h_ng.numpy(), x["num_gt"])
    o.submit_pinned()
    pend.append(o)
while pend:
    pend.pop(0).wait()
oko = True
for o in objs:
    gl, gc = o.flush()
    t64 = o.h_res[4:8].view(torch.float64).clone().cuda()
    dist.all_reduce(t64)
    oko &= abs(gl - t64[0].item()) <= 1e-12 * abs(gl) and abs(gc - t64[1].item()) <= 1e-12 * abs(gc)
ok &= oko
if rank == 0:
    print("dist_check world=%d pipelined own-stream steps (%d objects, one exchange state each): %s -> %s"
          % (world, npipe, "%.6f %.6f" % (gl, gc), "OK" if oko else "MISMATCH"))
# the final detection gather: every rank post-processes its shard of one global batch of patches; the packed
# all-gather must reproduce, bit for bit and in batch order, what one GPU computes on the whole batch
from multibox_b200 import detect  # noqa: E402
Bq = 6 * world + 0
qd = synth.make_detect_inputs(K=5, B=Bq, keep=50, seed=5, patches=True)
names = ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims", "is_flipped")
lo, hi = mdist.shard_range(Bq)


def run_detect(sl):
    t = {k: dev_(qd[k][sl]) for k in names}
    return detect.postprocess(t["locations"], t["confidences"], dev_(qd["priors"]), restrictions=t["restrictions"],
                              max_to_keep=t["max_to_keep"], offsets=t["offsets"], patch_dims=t["patch_dims"],
                              image_dims=t["image_dims"], is_flipped=t["is_flipped"], nms_iou=0.5, k_max=50)


mine = run_detect(slice(lo, hi))
glob = mdist.gather_detections({k: mine[k] for k in ("boxes", "scores", "prior_idx", "count")})
full = run_detect(slice(0, Bq))
torch.cuda.synchronize()
okg = all(torch.equal(glob[k], full[k]) for k in ("boxes", "scores", "prior_idx", "count"))
ok &= okg
if rank == 0:
    print("dist_check world=%d detection all-gather (%d patches, packed single NCCL call): %s"
          % (world, Bq, "OK" if okg else "MISMATCH"))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
