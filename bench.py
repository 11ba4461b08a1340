#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see the task description).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the training hot path (GT->prior matching + multibox
loss forward/backward) over one batch of synthetic head outputs.  The N=1
workload is BASELINE.json configs[1]: Inception-ResNet-v2-shaped head outputs,
5 aspect ratios (P=646 priors), batch 32, MAX_NUM_BBOXES=20.  With N>1 every rank
runs the same per-GPU batch (weak scaling; images are independent) and the two
loss scalars are SUM-all-reduced every step (fused into the kernel over NVLink
peer memory, checked against an NCCL all-reduce after the timed region).  Rank 0
prints ONE JSON line.  Further objects in the same line:
  "serialized"         the same steps without programmatic dependent launch
  "detect"             BASELINE configs[2] (decode + top-k + NMS, batch 256) [+ the detection all-gather at N>1]
  "strong_cfg3"        BASELINE configs[3] as written: batch 1024 TOTAL, sharded by image over the N GPUs
  "coco_person_shape"  the same shape, 1024 images per GPU (weak)
  "throughput_shape"   a BASELINE configs[4] per-GPU shard of the training step (K=11, M=200, 1024 images)
  "cfg5_detect"        the configs[4] detect leg (K=11, 1024 images per GPU, NMS 0.5)
  "layouts"            per-head / ragged entry points against the un-fused route
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------- helpers
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"       # /opt/skills/guides/B200_PROFILING.md


def traffic_from_profiles(kernel):
    """dram bytes per launch from the committed ncu summary, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(kernel)
            return e["dram_bytes_per_launch"] if isinstance(e, dict) else e
    except Exception:
        return None


def counters_from_profiles(key):
    """Per-launch ncu counters (warp instructions executed, DRAM bytes) of the committed capture of this
    workload (profiles/ncu_counters.json), for the issue-slot roofline."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def issue_roofline(key, kernel_s, sm_mhz, sms=148):
    """Issue-slot roofline of the dominant kernel: warp instructions per launch (ncu smsp__inst_executed.sum of
    the committed capture) / (SMs x 4 schedulers x clock x kernel time).  This, not HBM, is the bound the
    assignment solver runs against."""
    c = counters_from_profiles(key)
    if not c or not kernel_s:
        return None
    mhz = sm_mhz or 1965.0
    peak = sms * 4 * mhz * 1e6              # warp instructions / s
    ach = c["warp_inst_per_launch"] / kernel_s
    return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
            "warp_inst_per_launch": c["warp_inst_per_launch"], "sm_mhz": mhz,
            "source": "profiles/ncu_counters.json (%s)" % c.get("capture", "ncu --set full")}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def train_bytes_per_image(P, M, nbar):
    """Algorithmic HBM bytes of match + loss fwd/bwd per image as this bench runs
    it (minus the outputs the step does not request):
    read offsets 16P + conf 4P + GT 16*n + count 4; write d_loc 16P + d_conf 4P +
    per-image partials 20.  (mask / matched index / stacked GT are optional
    outputs, not produced in the training step.)"""
    return 40 * P + 16 * nbar + 24


def detect_bytes_per_image(P, k):
    """read offsets 16P + conf 4P + restriction 16 + keep 4 + conversion 28;
    write k * (boxes f64 32 + patch boxes 16 + score 4 + idx 4) + count 4."""
    return 20 * P + 56 * k + 52


# ----------------------------------------------------------------------------- CPU legs (oracle)
def _cpu_train_once(d):
    from oracle import np_oracle
    return np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], d["alpha"])


_REF_D = None      # the batch, inherited by the forked workers (no per-step pickling)


def _cpu_train_shard(args):
    lo, hi = args
    d = _REF_D
    from oracle import np_oracle
    out = np_oracle.add_loss(d["locations"][lo:hi], d["confidences"][lo:hi], d["gt"][lo:hi], d["num_gt"][lo:hi],
                             d["priors"], d["alpha"])
    return out["location_loss_f64"], out["confidence_loss_f64"]


def cpu_baseline_train(d, budget_s=12.0):
    """The oracle port (numpy/scipy restatement with the reference's loop
    structure), one process / one core -- the way the reference executes this
    path (under the GIL inside tf.py_func)."""
    _cpu_train_once(d)
    n, t0 = 0, time.perf_counter()
    while True:
        _cpu_train_once(d)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 2000:
            break
    return {"value": d["B"] * n / el, "unit": "images/s", "cores": 1, "kind": "port",
            "sample": "%d passes of the full %d-image batch (numpy/scipy oracle port, 1 process) in %.1f s"
                      % (n, d["B"], el),
            "host_cpus": os.cpu_count()}


def cpu_baseline_train_c(d, budget_s=3.0):
    """Extra, stronger CPU figure: the plain-C port of the matching (cost + LSAP),
    single thread.  Matching only (it is >95% of the CPU path)."""
    from oracle import c_oracle
    B = d["B"]
    loc = (d["locations"].reshape(-1, 4) + np.tile(d["priors"], (B, 1))).astype(np.float32)
    conf = d["confidences"].reshape(-1) + np.float32(1e-10)
    c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
        n += 1
    el = time.perf_counter() - t0
    return {"value": B * n / el, "unit": "images/s", "cores": 1, "kind": "port",
            "sample": "%d passes, plain-C matching port (oracle/c), 1 thread" % n}


def cpu_baseline_detect(q, budget_s=6.0, nms=0.5):
    from oracle import np_oracle
    sub = 32
    args = [q[k][:sub] for k in ("locations", "confidences")] + [q["priors"]] + \
        [q[k][:sub] for k in ("restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims", "is_flipped")]
    n, t0 = 0, time.perf_counter()
    while True:
        np_oracle.postprocess(*args, nms_iou=nms)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": sub * n / el, "unit": "images/s", "cores": 1, "kind": "port",
            "sample": "%d passes over the first %d images of the batch (numpy oracle port)" % (n, sub)}


def _cpu_detect_shard(args):
    lo, hi = args
    q = _REF_Q
    from oracle import np_oracle
    sub = [q[k][lo:hi] for k in ("locations", "confidences")] + [q["priors"]] + \
        [q[k][lo:hi] for k in ("restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims", "is_flipped")]
    out = np_oracle.postprocess(*sub, nms_iou=q["nms_iou"])
    return int(sum(m["boxes"].shape[0] for m in out))


_REF_Q = None


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the
    oracle port: numpy + scipy restatement with the reference's loops) on all
    host cores, sharded by image across processes.  Each step processes the SAME GLOBAL batch the GPU
    arm processes at this N (32 images per GPU x N)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from multibox_b200 import synth
    N = max(1, args.gpus)
    cfg = dict(synth.TRAIN_CONFIGS["cfg2"])
    parts = []
    for r in range(N):                      # the N ranks' batches (same seeds as the GPU arm)
        c = dict(cfg)
        c["seed"] = cfg["seed"] + 7919 * r
        parts.append(synth.make_train_inputs(**c))
    d = dict(parts[0])
    for k in ("locations", "confidences", "gt", "num_gt"):
        d[k] = np.concatenate([q[k] for q in parts], 0)
    d["B"] = parts[0]["B"] * N
    B = d["B"]
    cores = max(1, min(os.cpu_count() or 1, B))
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    global _REF_D, _REF_Q
    _REF_D = d
    qcfg = dict(synth.DETECT_CONFIGS["cfg3"])
    _REF_Q = synth.make_detect_inputs(**qcfg)
    bounds = [((B * i) // cores, (B * (i + 1)) // cores) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_train_shard, bounds)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_train_shard, bounds)
        el = time.perf_counter() - t0
        # detect leg (BASELINE configs[2]): a bounded sample of the 256-image batch, all cores
        QB = _REF_Q["B"]
        sub = min(QB, 4 * cores)
        qb = [((sub * i) // cores, (sub * (i + 1)) // cores) for i in range(cores)]
        pool.map(_cpu_detect_shard, qb)
        dn, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 8.0:
            pool.map(_cpu_detect_shard, qb)
            dn += 1
        del_ = time.perf_counter() - t1
    value = B * args.steps / el
    dval = sub * dn / del_
    line = {
        "impl": "reference", "metric": "match+loss images/sec", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config_dict(parts[0], args.gpus),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "each step = the full global batch of %d images (%d per GPU x %d), sharded by "
                                   "image over %d processes; numpy/scipy oracle port of reference loss.py:8-117"
                                   % (B, parts[0]["B"], N, cores)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detect": {"metric": "decode+NMS images/sec", "value": dval, "unit": "images/s",
                   "cpu_baseline": {"value": dval, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "%d passes over the first %d images of the configs[2] batch, sharded "
                                              "over %d processes; numpy port of reference detect.py:408-436 + the "
                                              "project's greedy NMS (IoU 0.5)" % (dn, sub, cores)}},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def train_nsets(d):
    """Device-resident input sets the GPU arm rotates over (> 2x L2 worth of inputs)."""
    B, P, M = d["B"], d["P"], d["M"]
    per_set = 4 * (B * P * 4 + B * P + B * M * 4 + B)
    return max(2, min(1024, (2 * L2_BYTES) // per_set + 1))


def train_config_dict(d, n_gpus):
    """The `config` object of the bench line -- the SAME dict from both arms (the reference arm runs the same
    workload on the host; the cache / launch entries describe how the GPU arm keeps its timing honest)."""
    return {"workload": "BASELINE configs[1]: Inception-ResNet-v2 299x299 multibox head outputs (random init), "
                        "5 aspect ratios, P=%d priors, batch %d per GPU, MAX_NUM_BBOXES=%d: GT->prior matching + "
                        "location/confidence loss fwd/bwd" % (d["P"], d["B"], d["M"]),
            "K": d["K"], "P": d["P"], "batch_per_gpu": d["B"], "global_batch": d["B"] * n_gpus, "M": d["M"],
            "alpha": d["alpha"], "mean_gt_per_image": float(d["num_gt"].mean()),
            "parallelism": "image-sharded x%d" % n_gpus,
            "cache": "GPU arm: inputs rotate over %d device-resident sets (> 2x L2) so no step finds its inputs in L2"
                     % train_nsets(d),
            "launch": "GPU arm: one kernel per step, launched back to back with programmatic dependent launch "
                      "(MBX_FLAG_PDL): step k+1's load / logs / assignment solve start while step k's "
                      "stores, last-CTA reduction and completion are still in flight; every write of "
                      "step k+1 (gradients, losses, workspace) waits for step k (griddepcontrol.wait). "
                      "All K steps do all their work inside the timed region; the 'serialized' object is "
                      "the same loop without the overlap"}


# ----------------------------------------------------------------------------- GPU legs
def rotated_sets(t, nsets):
    """nsets device copies of a batch tensor, images rolled by r: same distribution,
    distinct addresses, so that consecutive steps never find their inputs in L2."""
    import torch
    return [torch.roll(t, shifts=r, dims=0).contiguous() if r else t.clone() for r in range(nsets)]


MIN_TIMED_STEPS = 200      # a region shorter than this is repeated (see time_region)
LAST_REGION = {}           # host-side time the last timed region spent ENQUEUEING its steps (per step)


def time_region(fn, steps, warmup, barrier):
    """W untimed steps, then `steps` steps between barrier+synchronize on both sides, timed with CUDA
    events on the launching (current) stream.  A step of this path lasts 5-40 us, so a region of a few
    dozen steps would be a sub-millisecond measurement: when `steps` < MIN_TIMED_STEPS the region is
    repeated back to back (R x `steps` steps inside ONE event pair) and the time of `steps` steps is
    reported as total / R -- every line states steps and timed_steps."""
    import torch
    for i in range(warmup):
        fn(i)
    reps = max(1, -(-MIN_TIMED_STEPS // max(1, steps)))
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_cpu = time.perf_counter()
    for i in range(steps * reps):
        fn(warmup + i)
    t_cpu = time.perf_counter() - t_cpu
    e1.record()
    torch.cuda.synchronize()
    barrier()
    LAST_REGION["host_enqueue_us_per_step"] = 1e6 * t_cpu / (steps * reps)
    return e0.elapsed_time(e1) / 1e3 / reps


_PIPE_PEERS = []


def pipe_peer(r, peer):
    """The r-th PeerAllreduce of the pipelined host-buffer steps (one exchange state per stream; created
    once, in the same order on every rank, and reused by every bench object)."""
    if peer is None or peer.world == 1:
        return None
    from multibox_b200 import dist as mdist
    while len(_PIPE_PEERS) <= r:
        _PIPE_PEERS.append(mdist.PeerAllreduce())
    return _PIPE_PEERS[r]


def bench_train(d, steps, warmup, world, barrier, peer, want_e2e=True, pdl=True, check_allreduce=False,
                static_schedule=False):
    """Device-resident steps (value) and host-buffer steps (e2e) of match + loss fwd/bwd on batch `d`."""
    import torch
    from multibox_b200 import loss
    B, P, M = d["B"], d["P"], d["M"]
    dev = torch.device("cuda", torch.cuda.current_device())
    nsets = train_nsets(d)
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in ("locations", "confidences", "gt", "num_gt")}
    locs = rotated_sets(t["locations"], nsets)
    confs = rotated_sets(t["confidences"].view(B, P), nsets)
    gts = rotated_sets(t["gt"], nsets)
    ngs = rotated_sets(t["num_gt"], nsets)
    step = loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], device=dev, peer=peer, deferred_allreduce=True,
                                 pdl=pdl, static_schedule=static_schedule)
    # one pre-marshalled launch closure per input set: a step is ONE foreign call + one kernel
    launches = [step.prepare(locs[s], confs[s], gts[s], ngs[s]) for s in range(nsets)]
    torch.cuda.synchronize()
    barrier()

    def one(i):
        # N > 1: the SUM all-reduce of the two loss scalars is fused into this kernel (peer
        # memory over NVLink, mbx_match_loss_allreduce); no separate collective is launched
        launches[i % nsets]()

    sec = time_region(one, steps, warmup, barrier)
    res = {"sec": sec, "nsets": nsets, "launches_per_step": 1,
           "host_enqueue_us_per_step": LAST_REGION.get("host_enqueue_us_per_step")}
    if check_allreduce and world > 1:
        # the fused all-reduce against NCCL on the same step: complete the newest step's reduction
        # (deferred mode), then SUM-all-reduce the local fp64 sums of that step over NCCL
        import torch.distributed as dist
        g = step.flush()
        local = step.out["results"].to(dev).view(torch.float64)[2:4].clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        want = [float(x) for x in local.cpu()]
        ok = all(abs(a - b) <= 1e-12 * max(1.0, abs(b)) for a, b in zip(g, want))
        res["allreduce_check"] = "ok" if ok else "MISMATCH fused=%r nccl=%r" % (list(g), want)
        try:     # (diagnostics only: the symmetric buffer's step counter and pull-route counter, rank-local)
            raw = peer._keep[0][16:32].view(torch.int32).cpu().tolist()
            res["allreduce_stats"] = {"steps_on_this_buffer": raw[0], "forwarded_by_relay": raw[1],
                                      "completed_by_pull_route": raw[3],
                                      "note": "deferred steps whose words the relay kernel had not delivered were "
                                              "pulled from the peers' outboxes over NVLink instead (rank 0's "
                                              "counters since the buffer was created: prepare + warm-up + timed)"}
        except Exception:       # noqa: BLE001
            pass
    if want_e2e:
        # host path.  Four step objects (own pinned staging buffer, own mapped result block, own gradients)
        # rotate, so the H2D source is not one hot buffer; the kernel streams the packed pinned inputs
        # over PCIe itself (zero_copy) and stores the 64-byte result block into mapped host memory.
        hsets = 4
        roll = [dict(locations=np.roll(d["locations"], r, axis=0),
                     confidences=np.roll(d["confidences"].reshape(B, P), r, axis=0),
                     gt=np.roll(d["gt"], r, axis=0), num_gt=np.roll(d["num_gt"], r, axis=0)) for r in range(hsets)]

        def make(use_graph, use_pdl):
            out = []
            for r in range(hsets):
                hs = loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], device=dev, use_graph=use_graph, peer=peer,
                                           deferred_allreduce=True, host_results=True, zero_copy=True, pdl=use_pdl)
                stage(hs, r)
                out.append(hs)
            return out

        def stage(hs, r):
            np.copyto(hs.h_loc.numpy(), roll[r]["locations"])
            np.copyto(hs.h_conf.numpy(), roll[r]["confidences"])
            np.copyto(hs.h_gt.numpy(), roll[r]["gt"])
            np.copyto(hs.h_ng.numpy(), roll[r]["num_gt"])

        def wall(fn, n, drain=None, nwarm=hsets):
            nwarm = max(warmup, nwarm)       # (every rotating object has run once before the timed region)
            for i in range(nwarm):
                fn(i)
            if drain:
                drain()
            reps = max(1, -(-MIN_TIMED_STEPS // max(1, n)))       # (same rule as time_region)
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n * reps):
                fn(nwarm + i)
            if drain:
                drain()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        last = {}
        # (a) one step in flight: submit, poll, next (the round-1 mode; CUDA graph of the one kernel)
        ser = make(True, False)

        def e2e_serial(i):
            last["v"] = ser[i % hsets].step_pinned(validate=True)
            last["i"] = i

        res["e2e_serial_sec"] = wall(e2e_serial, steps)
        res["last"] = last["v"]
        res["last_global"] = ser[last["i"] % hsets].flush()
        # (b) SEVERAL steps in flight: rotating step objects, each with a CUDA stream of its own
        # (own_stream=True).  submit = ONE foreign call that enqueues {H2D copy of the packed pinned inputs on
        # the copy engine, the kernel} on the object's stream; the copy of step k+1 overlaps the kernels of the
        # steps before it, and the host only polls the mapped result block of the OLDEST step.  A single
        # host-buffer step has ~35 us of latency (launch, PCIe, solve, result write-back); profiles/e2e_depth.py
        # measured the depth sweep (round 2: kernel-streamed zero-copy inputs under PDL level off at 15.2 us per
        # step, copy engine + own streams at 13.0 us with seven steps in flight).  With a fused all-reduce every
        # object has its own PeerAllreduce (steps on different streams cannot share one exchange state).
        npipe = 8 if B * P * 20 < (4 << 20) else 4
        pipe = []
        for r in range(npipe):
            hs = loss.MultiboxLossStep(B, P, M, d["priors"], d["alpha"], device=dev, peer=pipe_peer(r, peer),
                                       deferred_allreduce=True, host_results=True, own_stream=True)
            stage(hs, r % hsets)
            pipe.append(hs)
        pend = []

        def e2e_pipe(i, numpy_in=False):
            hs = pipe[i % npipe]
            if numpy_in:                      # a caller that holds plain numpy arrays pays this staging copy
                stage(hs, i % hsets)
            hs.submit_pinned()
            pend.append(hs)
            if len(pend) > npipe - 1:         # npipe - 1 steps stay in flight
                last["p"] = pend.pop(0).wait(validate=True)

        def drain():
            while pend:
                last["p"] = pend.pop(0).wait(validate=True)

        res["e2e_sec"] = wall(e2e_pipe, steps, drain, nwarm=npipe)
        res["e2e_numpy_in_sec"] = wall(lambda i: e2e_pipe(i, True), steps, drain, nwarm=npipe)
        res.update(h2d=pipe[0].h2d_bytes, d2h=pipe[0].d2h_bytes, last_pipe=last["p"], steps_in_flight=npipe - 1)
    return res


def bench_layouts(d, barrier, steps=20, warmup=3):
    """Per-head / ragged entry points against the reference's own layout steps done in torch."""
    import torch
    from multibox_b200 import loss, synth
    B, P, M, K = d["B"], d["P"], d["M"], d["K"]
    dev = torch.device("cuda", torch.cuda.current_device())
    hl, hc = synth.split_heads(d["locations"], d["logits"], K)
    per_set = 4 * B * P * 5
    nsets = max(2, min(64, (2 * L2_BYTES) // per_set + 1))
    hls = [[torch.from_numpy(np.roll(t, r, axis=0).copy()).to(dev) for t in hl] for r in range(nsets)]
    hcs = [[torch.from_numpy(np.roll(t, r, axis=0).copy()).to(dev) for t in hc] for r in range(nsets)]
    gt = torch.from_numpy(d["gt"]).to(dev)
    ng = torch.from_numpy(d["num_gt"]).to(dev)
    flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
    flat_t, off_t = torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev)
    pri = torch.from_numpy(d["priors"]).to(dev)
    out_dense = {}

    def unfused(i):       # model.py:295-320 as torch ops, then the dense kernel (sigmoid fused either way)
        loc, logit = loss.concat_heads(hls[i % nsets], hcs[i % nsets])
        loss.match_loss_raw(loc.contiguous(), logit.view(B, P), gt, ng, pri, d["alpha"], flags=1, out=out_dense)

    out_h, out_r = {}, {}

    def fused(i):
        loss.match_loss_heads_raw(hls[i % nsets], hcs[i % nsets], gt, ng, pri, d["alpha"], flags=1, out=out_h)

    def fused_ragged(i):
        loss.match_loss_heads_raw(hls[i % nsets], hcs[i % nsets], flat_t, None, pri, d["alpha"], flags=1,
                                  gt_row_offsets=off_t, max_num_bboxes=M, out=out_r)

    res = {"workload": "configs[3] shape (K=%d, P=%d, B=%d per GPU, M=%d): match + loss fwd/bwd from the six heads' "
                       "NHWC outputs" % (K, P, B, M)}
    for name, fn in (("concat_in_torch_then_dense_ms", unfused), ("fused_heads_ms", fused),
                     ("fused_heads_ragged_gt_ms", fused_ragged)):
        res[name] = 1e3 * time_region(fn, steps, warmup, barrier) / steps
    res["note"] = ("the fused route never writes the concatenated [B,P,5] tensor nor splits its gradient "
                   "(4 x 20P bytes per image of extra HBM traffic and 2+ extra launches in the un-fused route; "
                   "the un-fused figure excludes the gradient split)")
    return res


def bench_detect(q, steps, warmup, barrier, nms_iou, want_e2e=True, zero_copy=False, world=1, pdl=True):
    import torch
    from multibox_b200 import detect
    B, P, keep = q["B"], q["P"], q["keep"]
    dev = torch.device("cuda", torch.cuda.current_device())
    per_set = 4 * (B * P * 5)
    nsets = max(2, min(256, (2 * L2_BYTES) // per_set + 1))
    names = ("locations", "confidences", "restrictions", "max_to_keep", "offsets", "patch_dims", "image_dims",
             "is_flipped")
    t = {k: torch.from_numpy(np.ascontiguousarray(q[k])).to(dev) for k in names}
    pri = torch.from_numpy(q["priors"]).to(dev)
    sets = {k: rotated_sets(t[k], nsets) for k in names}
    out = {}

    def one(i):
        s = i % nsets
        detect.postprocess(sets["locations"][s], sets["confidences"][s], pri, restrictions=sets["restrictions"][s],
                           max_to_keep=sets["max_to_keep"][s], offsets=sets["offsets"][s],
                           patch_dims=sets["patch_dims"][s], image_dims=sets["image_dims"][s],
                           is_flipped=sets["is_flipped"][s], nms_iou=nms_iou, k_max=keep, want_patch_boxes=False,
                           out=out, pdl=pdl_now[0])

    pdl_now = [False]
    ser_sec = time_region(one, steps, warmup, barrier)      # every kernel waits for the previous one to drain
    pdl_now[0] = bool(pdl)
    sec = time_region(one, steps, warmup, barrier)          # programmatic dependent launch: steps overlap
    res = {"sec": sec, "ser_sec": ser_sec, "kernel_ms": 1e3 * ser_sec / steps, "nsets": nsets}
    if world > 1:
        # the final detection gather (north star): every rank's padded detections all-gathered over NCCL
        # (NVLink) after its kernel, every step, inside the timed region
        from multibox_b200 import dist as mdist
        gathered = {}

        def one_gather(i):
            one(i)
            gathered["g"] = mdist.gather_detections({k: out[k] for k in ("boxes", "scores", "prior_idx", "count")})

        res["gather_sec"] = time_region(one_gather, steps, warmup, barrier)
        g = gathered["g"]
        res["gather_bytes"] = int(sum(v.numel() * v.element_size() for v in g.values()))
        res["gather_rows"] = int(g["count"].shape[0])
    if want_e2e:
        hsets = 2
        dsteps_ = []
        for r in range(hsets):
            ds = detect.DetectStep(B, P, keep, q["priors"], nms_iou=nms_iou, device=dev, use_graph=True,
                                   zero_copy=zero_copy)
            ds.fill_host(**{k: np.roll(q[k], r, axis=0) for k in names})
            dsteps_.append(ds)

        def e2e_step(i):
            return dsteps_[i % hsets].run_pinned()      # graph: packed H2D, kernel, packed D2H; sync

        for i in range(max(warmup, hsets)):
            e2e_step(i)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            e2e_step(i)
        torch.cuda.synchronize()
        res.update(e2e_sec=time.perf_counter() - t0, h2d=dsteps_[0].h2d_bytes, d2h=dsteps_[0].d2h_bytes)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the detect / throughput-shape / CPU legs")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from multibox_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

        def barrier():
            dist.barrier()

        from multibox_b200 import dist as mdist
        peer = mdist.PeerAllreduce()
    else:
        def barrier():
            pass

        peer = None

    peak, peak_kind = measured_peak()
    cfg = dict(synth.TRAIN_CONFIGS["cfg2"])
    cfg["seed"] = cfg["seed"] + 7919 * rank          # every rank owns different images
    d = synth.make_train_inputs(**cfg)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(x):
        """min / median / max over ranks of a per-rank time (separates slowest-image skew from communication)."""
        if world == 1:
            return None
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        v = sorted(float(p.item()) for p in parts)
        return {"min": v[0], "median": v[len(v) // 2], "max": v[-1]}

    def train_object(dd, steps, label, counters_key, note, want_e2e):
        """value / roofline (/ e2e) of the training step on batch `dd` (per GPU)."""
        tt = bench_train(dd, steps, 3, world, barrier, peer, want_e2e=want_e2e, pdl=True)
        ts = bench_train(dd, steps, 3, world, barrier, peer, want_e2e=False, pdl=False)
        sec, ssec = max_over_ranks(tt["sec"]), max_over_ranks(ts["sec"])
        Bq, Pq = dd["B"], dd["P"]
        k_s = ssec / steps                                      # one serialized back-to-back launch
        nb = train_bytes_per_image(Pq, dd["M"], float(dd["num_gt"].mean())) * Bq
        ach = nb / k_s / 1e9
        o = {"workload": label, "value": world * Bq * steps / sec, "unit": "images/s", "ms_per_step": 1e3 * sec / steps,
             "steps": steps, "batch_per_gpu": Bq, "global_batch": Bq * world,
             "serialized_ms_per_step": 1e3 * k_s,
             "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                          "kernel_ms": 1e3 * k_s, "bytes_per_launch": nb,
                          "traffic": (counters_from_profiles(counters_key) or {}).get("dram_bytes_per_launch"),
                          "note": note},
             "issue_roofline": issue_roofline(counters_key, k_s, (clocks or {}).get("sm_mhz") if clocks else None)}
        o["roofline"]["frac_at_step_rate"] = nb / (sec / steps) / 1e9 / peak      # with consecutive steps overlapped (PDL)
        if want_e2e:
            e2 = max_over_ranks(tt["e2e_sec"])
            o["e2e"] = {"value": world * Bq * steps / e2, "unit": "images/s", "h2d_bytes_per_step": tt["h2d"],
                        "d2h_bytes_per_step": tt["d2h"],
                        "numpy_in_value": world * Bq * steps / max_over_ranks(tt["e2e_numpy_in_sec"]),
                        "one_in_flight_value": world * Bq * steps / max_over_ranks(tt["e2e_serial_sec"])}
        return o

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tr = bench_train(d, args.steps, args.warmup, world, barrier, peer, pdl=True, check_allreduce=True)
    clocks = sampler.stop() if rank == 0 else None
    # the same steps WITHOUT programmatic dependent launch: every kernel waits for the previous one to drain
    # (this per-step time is the kernel's duration for the roofline: one launch, back to back, no overlap)
    ser = bench_train(d, args.steps, args.warmup, world, barrier, peer, want_e2e=False, pdl=False)

    sec, ssec = max_over_ranks(tr["sec"]), max_over_ranks(ser["sec"])
    e2e_sec = max_over_ranks(tr["e2e_sec"])
    kernel_ms = 1e3 * ssec / args.steps
    B, P, M = d["B"], d["P"], d["M"]
    nbar = float(d["num_gt"].mean())
    bytes_per_launch = train_bytes_per_image(P, M, nbar) * B
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": "match+loss images/sec", "value": world * B * args.steps / sec, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "timed_steps": args.steps * max(1, -(-MIN_TIMED_STEPS // max(1, args.steps))),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config_dict(d, world),
        "serialized": {"value": world * B * args.steps / ssec, "unit": "images/s", "ms_per_step": kernel_ms,
                       "per_rank_sec": per_rank(ser["sec"])},
        "per_rank_sec": per_rank(tr["sec"]),
        "host_enqueue_us_per_step": tr.get("host_enqueue_us_per_step"),
        "e2e": {"value": world * B * args.steps / e2e_sec, "unit": "images/s",
                "h2d_bytes_per_step": tr["h2d"], "d2h_bytes_per_step": tr["d2h"],
                "numpy_in_value": world * B * args.steps / max_over_ranks(tr["e2e_numpy_in_sec"]),
                "one_in_flight_value": world * B * args.steps / max_over_ranks(tr["e2e_serial_sec"]),
                "steps_in_flight": tr["steps_in_flight"],
                "how": "MultiboxLossStep(host_results=True, own_stream=True): rotating step objects, each with its own "
                       "CUDA stream, SEVEN steps in flight (submit_pinned / wait).  Each step's inputs sit in one packed "
                       "PINNED host buffer (already staged there: 'value'; np.copyto of the caller's numpy arrays into it "
                       "inside the timed loop: 'numpy_in_value'); submit is ONE foreign call (mbx_match_plan_launch_staged) "
                       "that enqueues the H2D copy of h2d_bytes_per_step bytes and the ONE kernel of the step on the "
                       "object's stream, every step; the kernel solves and stores the 64-byte loss/status block straight "
                       "into mapped pinned host memory (loss all-reduce fused in when N > 1, one exchange state per "
                       "object); the host polls the OLDEST step's launch sequence word and checks its status, every step; "
                       "gradients stay on the device for the backward pass; wall clock.  'one_in_flight_value' = submit, "
                       "poll, next (CUDA graph of one kernel that streams the inputs over PCIe itself; the round-1 mode)"},
        "gpu_launches": tr["launches_per_step"] * args.steps,
        "collective": ("loss SUM all-reduce over NVLink peer memory without an NCCL call per step (tagged 8-byte words, "
                       "no fence / remote atomic): step k leaves its sums in its own outbox (local stores: the "
                       "matching kernel does no NVLink access), a one-warp relay kernel on a side stream forwards "
                       "eight steps' words per burst into every rank's table, and step k completes step k-12's "
                       "reduction (k-1's without PDL) from its own table -- or pulls the words from the peers' "
                       "outboxes when the relay has not delivered them; the last step is flushed after the timed "
                       "region; allreduce_check = the fused global sums of the last step against an NCCL "
                       "all-reduce of the ranks' local fp64 sums") if world > 1 else None,
        "allreduce_check": tr.get("allreduce_check"),
        "allreduce_stats": tr.get("allreduce_stats"),
        "last_losses": {"local": tr.get("last"), "global": tr.get("last_global"), "pipelined_local": tr.get("last_pipe")},
        "roofline": {"bound": "hbm", "kernel": "mbx_match_loss_reg_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "peak_kind": "of " + peak_kind,
                     "bytes_per_launch": bytes_per_launch, "kernel_ms": kernel_ms,
                     "kernel_ms_how": "per-step time of the serialized back-to-back region (one launch per step, "
                                      "no overlap): an upper bound of the kernel's duration",
                     "frac_at_step_rate": bytes_per_launch / (sec / args.steps) / 1e9 / peak,
                     "traffic": (counters_from_profiles("cfg2") or {}).get("dram_bytes_per_launch"),
                     "note": "launch/latency-bound at this batch: %d CTAs on 148 SMs, %.2f MB per launch; the "
                             "solver is bound by dependent shared-memory scans and issue slots, not HBM "
                             "(SURVEY.md 8d)" % (B, bytes_per_launch / 1e6)},
        "issue_roofline": issue_roofline("cfg2", kernel_ms * 1e-3, (clocks or {}).get("sm_mhz") if clocks else None),
        "clocks": clocks,
    }
    if world > 1:
        dist.barrier()

    if not args.no_extras:
        # ---- secondary: detect path, BASELINE configs[2]
        def detect_object(q, dsteps, label, counters_key, want_e2e=True):
            dr = bench_detect(q, dsteps, max(3, min(args.warmup, 5)), barrier, q["nms_iou"], want_e2e=want_e2e,
                              world=world)
            dsec = max_over_ranks(dr["sec"])
            dk = 1e3 * max_over_ranks(dr["ser_sec"]) / dsteps      # one serialized launch: the kernel's duration
            dbytes = detect_bytes_per_image(q["P"], q["keep"]) * q["B"]
            dach = dbytes / (dk * 1e-3) / 1e9
            o = {"metric": "decode+NMS images/sec", "value": world * q["B"] * dsteps / dsec, "unit": "images/s",
                 "steps": dsteps, "ms_per_step": 1e3 * dsec / dsteps, "serialized_ms_per_step": dk,
                 "config": {"workload": label, "launch": "back-to-back launches with programmatic dependent launch "
                            "(the next step's load / sort / NMS overlap this step's store phase); "
                            "serialized_ms_per_step = the same loop without the overlap"},
                 "roofline": {"bound": "hbm", "kernel": "mbx_detect_kernel", "achieved": dach, "peak": peak,
                              "unit": "GB/s", "frac": dach / peak, "bytes_per_launch": dbytes, "kernel_ms": dk,
                              "traffic": (counters_from_profiles(counters_key) or {}).get("dram_bytes_per_launch")},
                 "issue_roofline": issue_roofline(counters_key, dk * 1e-3, (clocks or {}).get("sm_mhz") if clocks else None)}
            if want_e2e:
                de2e = max_over_ranks(dr["e2e_sec"])
                o["e2e"] = {"value": world * q["B"] * dsteps / de2e, "unit": "images/s",
                            "h2d_bytes_per_step": dr["h2d"], "d2h_bytes_per_step": dr["d2h"]}
            if world > 1:
                gsec = max_over_ranks(dr["gather_sec"])
                o["with_gather"] = {"value": world * q["B"] * dsteps / gsec, "unit": "images/s",
                                    "ms_per_step": 1e3 * gsec / dsteps, "gathered_bytes_per_step": dr["gather_bytes"],
                                    "gathered_images": dr["gather_rows"],
                                    "how": "kernel + NCCL all-gather of the padded detections (boxes f64, scores, "
                                           "prior index, count) of all ranks, every step, inside the timed region"}
            return o

        qcfg = dict(synth.DETECT_CONFIGS["cfg3"])
        qcfg["seed"] += 7919 * rank
        q = synth.make_detect_inputs(**qcfg)
        dsteps = max(20, min(args.steps, 100))
        line["detect"] = detect_object(q, dsteps, "BASELINE configs[2]: sigmoid outputs -> decode + clip + filter + "
                                       "top-200 + greedy NMS (IoU 0.5) + convert, batch %d per GPU, P=%d"
                                       % (q["B"], q["P"]), "detect")
        # ---- BASELINE configs[4], detect leg: K=11 (P=1420), 8192 images over 8 GPUs = 1024 per GPU, NMS 0.5
        q5cfg = dict(synth.DETECT_CONFIGS["cfg5d"])
        q5cfg["B"] = 1024
        q5cfg["seed"] += 7919 * rank
        q5 = synth.make_detect_inputs(**q5cfg)
        line["cfg5_detect"] = detect_object(q5, 10, "BASELINE configs[4] detect leg: 11 aspect ratios (P=%d), 1024 images "
                                            "per GPU (= 8192 over 8 GPUs), top-200 + NMS 0.5" % q5["P"], "detect_cfg5")
        # ---- a configs[4]-shaped per-GPU shard of the training step: where the kernel is throughput-bound
        tcfg = dict(K=11, B=1024, M=200, dist="uniform", seed=1005 + 7919 * rank, alpha=1000.0)
        td = synth.make_train_inputs(**tcfg)
        line["throughput_shape"] = train_object(
            td, 10, "BASELINE configs[4] shape: 11 aspect ratios (P=1420), MAX_NUM_BBOXES=200, 1024 images per GPU "
                    "(= 8192 over 8 GPUs)", "big",
            "assignment solver: n*P cheap cost bounds + a few exact costs per row (fp32 cost + fp64 duals), "
            "issue/latency-bound, reported against the HBM figure as the contract asks", want_e2e=True)
        evals = float((td["num_gt"].astype(np.float64) * td["P"]).sum())
        line["throughput_shape"]["cost_entries_per_s"] = line["throughput_shape"]["value"] / 1024.0 * evals
        # ---- BASELINE configs[3] as written: batch 1024 TOTAL, sharded by image over the N GPUs (strong scaling)
        c4 = dict(synth.TRAIN_CONFIGS["cfg4"])
        full = synth.make_train_inputs(**c4)
        lo, hi = (1024 * rank) // world, (1024 * (rank + 1)) // world
        shard = dict(full)
        for k in ("locations", "confidences", "logits", "gt", "num_gt"):
            if k in shard:
                shard[k] = np.ascontiguousarray(full[k][lo:hi])
        shard["B"] = hi - lo
        so = train_object(shard, 20, "BASELINE configs[3]: COCO-person-shaped training step (K=7, P=904, "
                          "MAX_NUM_BBOXES=100), batch 1024 TOTAL sharded by image over %d GPU(s): %d images per GPU"
                          % (world, hi - lo), "cfg4" if world == 1 else None,
                          "strong scaling: the per-GPU shard shrinks with N (128 images per GPU at N=8: the kernel's "
                          "latency regime)", want_e2e=True)
        so["scaling"] = "strong"
        so["value"] = 1024 * 20 / (so["ms_per_step"] * 1e-3 * 20)       # the GLOBAL batch per step time
        so["global_batch"] = 1024
        if "e2e" in so:
            for kk in ("value", "numpy_in_value", "one_in_flight_value"):
                so["e2e"][kk] = so["e2e"][kk] * 1024.0 / (shard["B"] * world)
        line["strong_cfg3"] = so
        # ---- the same shape, 1024 images per GPU (weak)
        c4w = dict(c4)
        c4w["seed"] += 7919 * rank
        d4 = synth.make_train_inputs(**c4w)
        line["coco_person_shape"] = train_object(
            d4, 20, "BASELINE configs[3] shape: 7 aspect ratios (P=904), MAX_NUM_BBOXES=100, COCO-person-like GT "
                    "counts (mean %.1f), %d images per GPU" % (float(d4["num_gt"].mean()), d4["B"]), "cfg4",
            "sparse GT: the step is dominated by the exact fp32 log arithmetic and per-image fixed costs, not by HBM",
            want_e2e=False)
        # ---- the data formats either side of the path, configs[3] shape:
        # per-head NHWC inputs (concat + sigmoid fused away) and ragged ground truth, against the
        # un-fused route (torch.cat of the six heads, then the dense entry point)
        line["layouts"] = bench_layouts(d4, barrier)
        if rank == 0 and world == 1:
            line["cpu_baseline"] = cpu_baseline_train(d)
            line["cpu_baseline_c_port"] = cpu_baseline_train_c(d)
            line["detect"]["cpu_baseline"] = cpu_baseline_detect(q)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _guard_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout; libraries (NCCL prints its version
    banner on stdout) must not.  Route fd 1 to stderr for the run and keep the real stdout
    for the final line."""
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(real, "w")


if __name__ == "__main__":
    _real_stdout = _guard_stdout()
    _print = print

    def print(*a, **k):      # noqa: A001  (the JSON line goes to the real stdout)
        k.setdefault("file", _real_stdout)
        _print(*a, **k)
        _real_stdout.flush()

    sys.exit(main())
