"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing in multibox_b200/dist.py:
image sharding, SUM all-reduce of the loss scalars, detection / stacked-GT gathers.
The per-rank compute is stood in for by the oracle here (no GPU in this
container); the thing under test is the host-side sharding + collective logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multibox_b200 import dist as mdist
from multibox_b200 import synth
from oracle import np_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 7                                       # odd on purpose: shards of 4 and 3
        d = synth.make_train_inputs(K=5, B=B, M=20, seed=77, edge_cases=True)
        lo, hi = mdist.shard_range(B)
        sh = mdist.shard_batch({k: d[k] for k in ("locations", "confidences", "gt", "num_gt", "priors")}, B)
        assert sh["locations"].shape[0] == hi - lo and sh["priors"].shape == d["priors"].shape
        out = np_oracle.add_loss(sh["locations"], sh["confidences"], sh["gt"], sh["num_gt"], d["priors"], d["alpha"])
        losses = torch.tensor([out["location_loss_f64"], out["confidence_loss_f64"]], dtype=torch.float64)
        mdist.allreduce_losses(losses)
        stacked = mdist.gather_stacked_gt(torch.from_numpy(out["stacked_gt"]))
        mask = mdist.gather_variable_batch(torch.from_numpy(out["mask"].reshape(hi - lo, -1)), B)
        # detections: equal shards (B=6)
        qd = synth.make_detect_inputs(K=5, B=6, keep=20, seed=5, patches=True)
        l2, h2 = mdist.shard_range(6)
        post = np_oracle.postprocess(qd["locations"][l2:h2], qd["confidences"][l2:h2], qd["priors"],
                                     qd["restrictions"][l2:h2], qd["max_to_keep"][l2:h2], qd["offsets"][l2:h2],
                                     qd["patch_dims"][l2:h2], qd["image_dims"][l2:h2], qd["is_flipped"][l2:h2])
        pad_idx = torch.full((h2 - l2, 20), -1, dtype=torch.int32)
        cnt = torch.zeros((h2 - l2,), dtype=torch.int32)
        for i, m in enumerate(post):
            c = m["prior_idx"].shape[0]
            pad_idx[i, :c] = torch.from_numpy(m["prior_idx"].astype(np.int32))
            cnt[i] = c
        g = mdist.gather_detections({"prior_idx": pad_idx, "count": cnt})
        if rank == 0:
            q.put(dict(losses=losses.numpy(), stacked=stacked.numpy(), mask=mask.numpy(),
                       idx=g["prior_idx"].numpy(), cnt=g["count"].numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.make_train_inputs(K=5, B=7, M=20, seed=77, edge_cases=True)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], d["alpha"])
    np.testing.assert_allclose(got["losses"], [ref["location_loss_f64"], ref["confidence_loss_f64"]], rtol=1e-12)
    assert np.array_equal(got["stacked"], ref["stacked_gt"])          # (image, prior) order preserved
    assert np.array_equal(got["mask"].reshape(-1), ref["mask"])
    qd = synth.make_detect_inputs(K=5, B=6, keep=20, seed=5, patches=True)
    post = np_oracle.postprocess(qd["locations"], qd["confidences"], qd["priors"], qd["restrictions"],
                                 qd["max_to_keep"], qd["offsets"], qd["patch_dims"], qd["image_dims"],
                                 qd["is_flipped"])
    for b, m in enumerate(post):
        c = m["prior_idx"].shape[0]
        assert got["cnt"][b] == c and np.array_equal(got["idx"][b, :c], m["prior_idx"])


def test_shard_range_covers_batch():
    for B in (1, 7, 32, 1024, 8191):
        for ws in (1, 2, 4, 8):
            r = [mdist.shard_range(B, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
