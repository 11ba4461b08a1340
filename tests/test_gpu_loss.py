"""-m gpu parity tests of the fused loss forward/backward (through the C ABI and
the autograd mirror of reference loss.py:55) against the oracle.
Tolerance (BASELINE.json north_star): losses, gradients within 1e-5 relative in
fp32; matched indices bit-exact."""
import numpy as np
import pytest
import torch

from multibox_b200 import loss, synth
from oracle import np_oracle
from gpu_util import dev

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _run_gpu(d, alpha, logits=False, warps=0, generic=False, cols=0):
    conf_in = d["logits"] if logits else d["confidences"]
    out = loss.match_loss_raw(dev(d["locations"]), dev(conf_in).view(d["B"], d["P"]), dev(d["gt"]), dev(d["num_gt"]),
                              dev(d["priors"]), alpha, flags=(1 if logits else 0) | (4 if generic else 0),
                              want_mask=True, want_gt_idx=True,
                              want_stacked=True, want_grads=True, want_conf_out=logits, warps=warps, cols=cols)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check(d, alpha, out, ref):
    B, P = d["B"], d["P"]
    assert out["results"][2] == 0
    assert np.array_equal(out["mask"], ref["mask"])
    assert np.array_equal(out["matched_gt_idx"], ref["matched_gt_idx"])
    n = int(out["n_stacked"][0])
    assert np.array_equal(out["stacked_gt"][:n], ref["stacked_gt"])
    assert int(out["results"][3]) == int(ref["mask"].sum())
    f64 = out["results"].view(np.float64)[2:4]
    assert abs(f64[0] - ref["location_loss_f64"]) <= RTOL * abs(ref["location_loss_f64"]) + 1e-30
    assert abs(f64[1] - ref["confidence_loss_f64"]) <= RTOL * abs(ref["confidence_loss_f64"])
    np.testing.assert_allclose(out["results"][0], ref["location_loss"], rtol=RTOL)
    np.testing.assert_allclose(out["results"][1], ref["confidence_loss"], rtol=RTOL)
    np.testing.assert_allclose(out["d_locations"], ref["d_locations"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(out["d_confidences"].reshape(B, P, 1), ref["d_confidences"], rtol=RTOL, atol=0)


@pytest.mark.parametrize("name,alpha", [("cfg2", 1000.0), ("cfg2", 1.0)])
def test_loss_cfg2(cuda_device, name, alpha):
    cfg = dict(synth.TRAIN_CONFIGS[name])
    cfg["alpha"] = alpha
    d = synth.make_train_inputs(edge_cases=True, **cfg)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], alpha)
    _check(d, alpha, _run_gpu(d, alpha), ref)


@pytest.mark.parametrize("K,B,M,dist", [(7, 48, 100, "coco_person"), (11, 6, 200, "uniform")])
def test_loss_other_shapes(cuda_device, K, B, M, dist):
    d = synth.make_train_inputs(K=K, B=B, M=M, dist=dist, seed=31 + K, edge_cases=True)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
    for warps, generic, cols in ((0, False, 0), (16, False, 0), (0, True, 0), (1, True, 0), (8, False, 8),
                                 (4, False, 0), (16, False, 4)):
        if cols and cols * warps * 32 < d["P"]:
            continue
        _check(d, 1000.0, _run_gpu(d, 1000.0, warps=warps, generic=generic, cols=cols), ref)


def test_loss_from_logits(cuda_device):
    d = synth.make_train_inputs(K=5, B=8, M=20, seed=55, edge_cases=True)
    out = _run_gpu(d, 1000.0, logits=True)
    out_g = _run_gpu(d, 1000.0, logits=True, generic=True)
    for k in ("results", "d_confidences", "d_locations", "mask", "confidences"):
        a, b = out[k].view(np.uint32), out_g[k].view(np.uint32)
        if k == "results":
            a, b = a[:15], b[:15]          # word 15 is the launch sequence number
        assert np.array_equal(a, b), k      # both kernels, same bits
    s_gpu = out["confidences"].reshape(d["B"], d["P"], 1)
    s_ref = torch.sigmoid(torch.from_numpy(d["logits"])).numpy()
    np.testing.assert_allclose(s_gpu, s_ref, rtol=2e-6, atol=1e-37)
    # feed the kernel's own sigmoid output to the oracle: everything downstream must agree
    ref = np_oracle.add_loss(d["locations"], s_gpu, d["gt"], d["num_gt"], d["priors"], 1000.0)
    assert np.array_equal(out["mask"], ref["mask"]) and np.array_equal(out["matched_gt_idx"], ref["matched_gt_idx"])
    np.testing.assert_allclose(out["results"][:2], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    d_logits = ref["d_confidences"] * s_gpu * (np.float32(1.) - s_gpu)
    np.testing.assert_allclose(out["d_confidences"].reshape(d["B"], d["P"], 1), d_logits, rtol=RTOL, atol=0)


def test_autograd_mirror_of_add_loss(cuda_device):
    d = synth.make_train_inputs(K=5, B=4, M=20, seed=12)
    locs = dev(d["locations"]).requires_grad_(True)
    confs = dev(d["confidences"]).requires_grad_(True)
    ll, cl = loss.add_loss(locs, confs, dev(d["gt"]), dev(d["num_gt"]), dev(d["priors"]), 1000.0)
    assert ll.shape == () and cl.shape == () and ll.dtype == torch.float32
    (ll + 2.0 * cl).backward()
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
    np.testing.assert_allclose(ll.item(), ref["location_loss"], rtol=RTOL)
    np.testing.assert_allclose(cl.item(), ref["confidence_loss"], rtol=RTOL)
    np.testing.assert_allclose(locs.grad.cpu().numpy(), ref["d_locations"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(confs.grad.cpu().numpy(), 2.0 * ref["d_confidences"], rtol=RTOL, atol=0)


def test_reference_test_properties(cuda_device):
    """The five properties reference model_tests.py pins, on the CUDA path."""
    rng = np.random.default_rng(0)
    priors = rng.uniform(size=(646, 4)).astype(np.float32)          # model_tests.py:111
    gt = np.zeros((2, 5, 4), np.float32)
    gt[0, 0] = [0.1, 0.1, 0.9, 0.9]
    locs = rng.normal(0, 0.1, size=(2, 646, 4)).astype(np.float32)
    confs = rng.uniform(0.01, 0.99, size=(2, 646, 1)).astype(np.float32)

    def run(sl, counts):
        ll, cl = loss.add_loss(dev(locs[sl]), dev(confs[sl]), dev(gt[sl]), dev(np.array(counts, np.int32)),
                               dev(priors), 1.0)
        return ll.item(), cl.item()

    ll, cl = run(slice(0, 1), [1])          # testSingleBoundingBox
    assert ll > 0 and cl > 0
    ll, cl = run(slice(0, 1), [0])          # testNoGTBoundingBox
    assert ll == 0.0 and cl > 0
    ll, cl = run(slice(0, 2), [1, 0])       # testBoxWithNoBox
    assert ll > 0 and cl > 0


@pytest.mark.parametrize("use_graph", [False, True])
def test_step_object_host_path(cuda_device, use_graph):
    d = synth.make_train_inputs(K=5, B=32, M=20, seed=1002)
    step = loss.MultiboxLossStep(32, d["P"], 20, d["priors"], 1000.0, use_graph=use_graph)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
    for _ in range(3):      # workspace / output reuse across steps
        ll, cl = step.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"])
        np.testing.assert_allclose([ll, cl], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    np.testing.assert_allclose(step.out["d_locations"].cpu().numpy(), ref["d_locations"], rtol=RTOL, atol=0)
    # a second, different batch through the same object (graph replays must pick up new host data)
    d2 = synth.make_train_inputs(K=5, B=32, M=20, seed=77)
    ref2 = np_oracle.add_loss(d2["locations"], d2["confidences"], d2["gt"], d2["num_gt"], d2["priors"], 1000.0)
    ll, cl = step.step_host(d2["locations"], d2["confidences"], d2["gt"], d2["num_gt"])
    np.testing.assert_allclose([ll, cl], [ref2["location_loss"], ref2["confidence_loss"]], rtol=RTOL)
    bad = d2["locations"].copy()
    bad[3, 5, 1] = np.nan
    with pytest.raises(ValueError):
        step.step_host(bad, d2["confidences"], d2["gt"], d2["num_gt"])
    ll, cl = step.step_host(d2["locations"], d2["confidences"], d2["gt"], d2["num_gt"])      # recovers
    np.testing.assert_allclose([ll, cl], [ref2["location_loss"], ref2["confidence_loss"]], rtol=RTOL)


def test_step_object_odd_batch(cuda_device):
    """Packed staging must keep every section 16-byte aligned for any B (found on a 9-image shard)."""
    for B in (1, 3, 9):
        d = synth.make_train_inputs(K=5, B=B, M=20, seed=B)
        step = loss.MultiboxLossStep(B, d["P"], 20, d["priors"], 1000.0, use_graph=True)
        ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
        ll, cl = step.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"])
        np.testing.assert_allclose([ll, cl], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
        assert step.global_losses() == pytest.approx((ref["location_loss_f64"], ref["confidence_loss_f64"]), rel=1e-9)


def test_determinism(cuda_device):
    d = synth.make_train_inputs(K=7, B=300, M=100, dist="coco_person", seed=2)
    a = _run_gpu(d, 1000.0)
    b = _run_gpu(d, 1000.0)
    assert np.array_equal(a["results"].view(np.uint32)[:15], b["results"].view(np.uint32)[:15])
    assert b["results"].view(np.uint32)[15] == a["results"].view(np.uint32)[15] + 1      # launch sequence number
    assert np.array_equal(a["d_confidences"].view(np.uint32), b["d_confidences"].view(np.uint32))


def test_dynamic_schedule_matches_static(cuda_device):
    """More images than resident CTAs: the heavy-first dynamically scheduled launch must give
    bit-identical losses, gradients and matches to the static image -> CTA assignment
    (per-image partials are reduced in image order whatever the processing order), run after
    run (the scheduler's queue counter is re-armed by each launch)."""
    from multibox_b200 import _lib
    d = synth.make_train_inputs(K=7, B=1500, M=100, dist="coco_person", seed=4)

    def run(flags):
        out = loss.match_loss_raw(dev(d["locations"]), dev(d["confidences"]).view(d["B"], d["P"]), dev(d["gt"]),
                                  dev(d["num_gt"]), dev(d["priors"]), 1000.0, flags=flags, want_mask=True,
                                  want_gt_idx=True, want_stacked=True, want_grads=True)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    st = run(_lib.FLAG_STATIC)
    assert st["results"][2] == 0
    ns = int(st["n_stacked"][0])
    assert ns == int(d["num_gt"].sum())
    for _ in range(3):
        dy = run(0)
        assert int(dy["n_stacked"][0]) == ns
        for k in ("mask", "matched_gt_idx", "d_locations", "d_confidences"):
            assert np.array_equal(st[k], dy[k]), k
        assert np.array_equal(st["stacked_gt"][:ns], dy["stacked_gt"][:ns])      # (rows beyond n_stacked are never written)
        assert np.array_equal(st["results"][:8].view(np.uint32), dy["results"][:8].view(np.uint32))
    # and both agree with the oracle on a sample of images
    sub = slice(0, 64)
    ref = np_oracle.add_loss(d["locations"][sub], d["confidences"][sub], d["gt"][sub], d["num_gt"][sub], d["priors"],
                             1000.0)
    P = d["P"]
    assert np.array_equal(dy["matched_gt_idx"].reshape(-1)[:64 * P], ref["matched_gt_idx"])


@pytest.mark.parametrize("use_graph,zero_copy", [(False, False), (True, False), (True, True), (False, True)])
def test_step_object_host_mapped_results(cuda_device, use_graph, zero_copy):
    """host_results=True: the kernel stores the result block into mapped pinned host memory and the
    host polls the launch sequence word instead of synchronising -- same numbers, step after step."""
    B = 32
    step = loss.MultiboxLossStep(B, 646, 20, synth.make_train_inputs(K=5, B=1, M=20, seed=0)["priors"], 1000.0,
                                 use_graph=use_graph, host_results=True, zero_copy=zero_copy)
    plain = loss.MultiboxLossStep(B, 646, 20, step.priors, 1000.0, use_graph=False)
    for seed in (1002, 7, 8, 9, 10):
        d = synth.make_train_inputs(K=5, B=B, M=20, seed=seed)
        ll, cl = step.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"])
        l2, c2 = plain.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"])
        assert (ll, cl) == (l2, c2)
        ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
        np.testing.assert_allclose([ll, cl], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
        assert step.global_losses() == pytest.approx((ref["location_loss_f64"], ref["confidence_loss_f64"]), rel=1e-9)
        torch.cuda.synchronize()
        np.testing.assert_allclose(step.out["d_locations"].cpu().numpy(), ref["d_locations"], rtol=RTOL)
    # a data-dependent failure still surfaces as the ValueError scipy would raise
    d = synth.make_train_inputs(K=5, B=B, M=20, seed=3)
    bad = d["confidences"].copy()
    bad[0, 5, 0] = np.nan
    with pytest.raises(ValueError):
        step.step_host(d["locations"], bad, d["gt"], d["num_gt"])
    ll, cl = step.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"])      # and the object recovers
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
    np.testing.assert_allclose([ll, cl], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)


def test_step_objects_on_own_streams(cuda_device):
    """own_stream=True: every object enqueues {H2D copy of its packed pinned inputs, kernel} on a stream of
    its own with one foreign call (mbx_match_plan_launch_staged); four rotating objects keep several steps
    in flight (copies and kernels of different steps overlap) and every step returns exactly what the plain
    path returns for its data, gradients included."""
    B, n = 32, 4
    pri = synth.make_train_inputs(K=5, B=1, M=20, seed=0)["priors"]
    objs = [loss.MultiboxLossStep(B, 646, 20, pri, 1000.0, host_results=True, own_stream=True) for _ in range(n)]
    plain = loss.MultiboxLossStep(B, 646, 20, pri, 1000.0)
    data = [synth.make_train_inputs(K=5, B=B, M=20, seed=500 + i) for i in range(7)]
    want, grads = [], []
    for d in data:
        want.append(plain.step_host(d["locations"], d["confidences"], d["gt"], d["num_gt"]))
        torch.cuda.synchronize()
        grads.append(plain.out["d_locations"].cpu().clone())
    pend, got = [], []
    for rnd in range(3):
        for i, d in enumerate(data):
            o = objs[(rnd * len(data) + i) % n]
            if len(pend) == n:                      # the object is still in flight: complete the oldest step first
                po, pi = pend.pop(0)
                got.append((pi, po.wait()))
                po.stream.synchronize()             # (gradients are read on another stream)
                assert torch.equal(po.out["d_locations"].cpu(), grads[pi])
            np.copyto(o.h_loc.numpy(), d["locations"])
            np.copyto(o.h_conf.numpy(), d["confidences"].reshape(B, -1))
            np.copyto(o.h_gt.numpy(), d["gt"])
            np.copyto(o.h_ng.numpy(), d["num_gt"])
            o.submit_pinned()
            pend.append((o, i))
    while pend:
        po, pi = pend.pop(0)
        got.append((pi, po.wait()))
    assert len(got) == 3 * len(data)
    for pi, val in got:
        assert val == want[pi]
    with pytest.raises(ValueError):
        loss.MultiboxLossStep(B, 646, 20, pri, 1000.0, own_stream=True)      # needs host_results


def test_step_pinned_foreign_buffer_all_modes(cuda_device):
    """step_pinned(pinned=...) with a caller-owned packed pinned buffer: same result as step_host in
    every (use_graph, zero_copy, host_results) mode (the launch closure / graph is bound to the
    object's own staging buffer, so the data must be brought there)."""
    B = 8
    d0 = synth.make_train_inputs(K=5, B=B, M=20, seed=71)
    d1 = synth.make_train_inputs(K=5, B=B, M=20, seed=72)
    ref = np_oracle.add_loss(d1["locations"], d1["confidences"], d1["gt"], d1["num_gt"], d1["priors"], 1000.0)
    for use_graph in (False, True):
        for zero_copy in (False, True):
            for host_results in (False, True):
                step = loss.MultiboxLossStep(B, 646, 20, d0["priors"], 1000.0, use_graph=use_graph,
                                             zero_copy=zero_copy, host_results=host_results)
                step.step_host(d0["locations"], d0["confidences"], d0["gt"], d0["num_gt"])     # stale data in h_in
                other = torch.empty_like(step.h_in).pin_memory()
                views = step._views(other)
                for v, a in zip(views, (d1["locations"], d1["confidences"].reshape(B, -1), d1["gt"], d1["num_gt"])):
                    v.copy_(torch.from_numpy(np.ascontiguousarray(a)))
                ll, cl = step.step_pinned(pinned=other)
                np.testing.assert_allclose([ll, cl], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL,
                                           err_msg=str((use_graph, zero_copy, host_results)))


def test_programmatic_dependent_launch_is_invisible(cuda_device):
    """MBX_FLAG_PDL lets step k+1 start while step k is still running and only orders the WRITES: a
    long run of back-to-back steps over alternating heavy (every image 20 GT boxes) and empty batches --
    the later step is often the faster one -- must leave exactly the last step's outputs behind, and
    every intermediate state observed through a stream-ordered copy must be that step's."""
    B, P, M = 32, 646, 20
    heavy = synth.make_train_inputs(K=5, B=B, M=M, dist="full", seed=5)
    light = synth.make_train_inputs(K=5, B=B, M=M, dist="uniform", seed=6)
    light["num_gt"][:] = 0
    mid = synth.make_train_inputs(K=5, B=B, M=M, dist="uniform", seed=7)
    batches = [heavy, light, mid]
    refs = [np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
            for d in batches]
    step = loss.MultiboxLossStep(B, P, M, heavy["priors"], 1000.0, pdl=True)
    launches = [step.prepare(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]))
                for d in batches]
    torch.cuda.synchronize()
    order = [0, 1, 0, 1, 2, 1, 0, 0, 1, 2] * 6
    snaps = []
    for k, w in enumerate(order):
        launches[w]()
        if k % 7 == 3 or k == len(order) - 1:          # stream-ordered observation of this step's outputs
            snaps.append((w, step.out["results"].clone(), step.out["d_locations"].clone(),
                          step.out["d_confidences"].clone()))
    torch.cuda.synchronize()
    for w, res, dl, dc in snaps:
        r = res.cpu().numpy()
        assert r[2] == 0
        f64 = r.view(np.float64)[2:4]
        assert abs(f64[0] - refs[w]["location_loss_f64"]) <= RTOL * abs(refs[w]["location_loss_f64"]) + 1e-30
        assert abs(f64[1] - refs[w]["confidence_loss_f64"]) <= RTOL * abs(refs[w]["confidence_loss_f64"])
        np.testing.assert_allclose(dl.cpu().numpy(), refs[w]["d_locations"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(dc.cpu().numpy().reshape(B, P, 1), refs[w]["d_confidences"], rtol=RTOL, atol=0)


def test_wrong_current_device_is_handled(cuda_device):
    """Tensors on cuda:1 while cuda:0 is current: the wrappers switch the current device
    for the launch (the C side launches on the current device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = synth.make_train_inputs(K=5, B=4, M=20, seed=9)
    ref = np_oracle.add_loss(d["locations"], d["confidences"], d["gt"], d["num_gt"], d["priors"], 1000.0)
    one = torch.device("cuda", 1)
    t = [torch.from_numpy(np.ascontiguousarray(d[k])).to(one) for k in ("locations", "confidences", "gt", "num_gt")]
    with torch.cuda.device(0):
        out = loss.match_loss_raw(t[0], t[1].view(4, -1), t[2], t[3], torch.from_numpy(d["priors"]).to(one), 1000.0)
        torch.cuda.synchronize(one)
    np.testing.assert_allclose(out["results"][:2].cpu().numpy(), [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    with pytest.raises(ValueError):
        loss.match_loss_raw(t[0], t[1].view(4, -1), t[2].to("cuda:0"), t[3], torch.from_numpy(d["priors"]).to(one), 1000.0)


def test_programmatic_dependent_launch_with_dynamic_scheduling(cuda_device):
    """PDL on batches larger than the resident CTAs: the heavy-first order is computed inside the kernel
    and consecutive launches alternate between two scheduler slots (order array, queue counter, ready
    flag), so an early-started step never touches what the previous one still uses.  Back-to-back steps
    over batches of very different weight must equal the plain (serialised) launches bit for bit."""
    B, P, M = 700, 646, 20
    heavy = synth.make_train_inputs(K=5, B=B, M=M, dist="full", seed=15)
    mixed = synth.make_train_inputs(K=5, B=B, M=M, dist="uniform", seed=16)
    light = synth.make_train_inputs(K=5, B=B, M=M, dist="uniform", seed=17)
    light["num_gt"][::2] = 0
    batches = [heavy, mixed, light]
    plain = loss.MultiboxLossStep(B, P, M, heavy["priors"], 1000.0)
    want = []
    for d in batches:
        o = plain.step(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]))
        torch.cuda.synchronize()
        want.append((o["results"].cpu().numpy().copy(), o["d_locations"].cpu().numpy().copy(),
                     o["d_confidences"].cpu().numpy().copy()))
    ref = np_oracle.add_loss(mixed["locations"], mixed["confidences"], mixed["gt"], mixed["num_gt"], mixed["priors"], 1000.0)
    np.testing.assert_allclose(want[1][0][:2], [ref["location_loss"], ref["confidence_loss"]], rtol=RTOL)
    step = loss.MultiboxLossStep(B, P, M, heavy["priors"], 1000.0, pdl=True)
    launches = [step.prepare(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]))
                for d in batches]
    torch.cuda.synchronize()
    order = [0, 2, 1, 2, 0, 0, 2, 1, 1, 2] * 5
    snaps = []
    for k, w in enumerate(order):
        launches[w]()
        if k % 6 == 4 or k == len(order) - 1:
            snaps.append((w, step.out["results"].clone(), step.out["d_locations"].clone(),
                          step.out["d_confidences"].clone()))
    torch.cuda.synchronize()
    for w, res, dl, dc in snaps:
        assert np.array_equal(res.cpu().numpy()[:15].view(np.uint32), want[w][0][:15].view(np.uint32)), w
        assert np.array_equal(dl.cpu().numpy().view(np.uint32), want[w][1].view(np.uint32)), w
        assert np.array_equal(dc.cpu().numpy().view(np.uint32), want[w][2].view(np.uint32)), w
