"""Matching + multibox loss -- host-side mirror of the reference's ``loss.py``.

Same function names, argument order, tensor layouts and error behaviour as the
reference (``compute_assignments`` loss.py:8, ``add_loss`` loss.py:55), with
CUDA ``torch.Tensor`` arguments in place of numpy arrays / TF tensors.  All
arithmetic runs in the hand-written sm_100a kernels behind the C ABI
(``mbx_match_loss``); this module only allocates outputs, passes pointers and
turns the device status word into the ``ValueError`` scipy would raise at
reference loss.py:40.  There is no CPU path.
"""
import ctypes

import torch

from . import _lib

SMALL_EPSILON = 1e-10   # reference loss.py:6

_workspaces = {}


def _workspace(device, nbytes):
    """Zero-initialised scratch, one per (device, stream); grown on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _f32c(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch.Tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32c(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch.Tensor (there is no CPU path)" % name)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _same_device(*tensors):
    """The device all (non-None) argument tensors live on; mixed devices are an argument error.
    Every C-ABI launch below runs under `torch.cuda.device(dev)`: the library launches on the CUDA
    *current* device (its occupancy / shared-memory attribute caches are keyed by it), which need not
    be the tensors' device in a multi-GPU process."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise TypeError("CUDA tensors expected (there is no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("all tensors of one call must live on one device (got %s and %s)" % (dev, t.device))
    return dev


def raise_for_status(status):
    """Same failures, same exception type as scipy.optimize.linear_sum_assignment."""
    status = int(status)
    if status & _lib.STATUS_INVALID_COST:
        raise ValueError("matrix contains invalid numeric entries")
    if status & _lib.STATUS_INFEASIBLE:
        raise ValueError("cost matrix is infeasible")
    if status & _lib.STATUS_BAD_NUM_GT:
        raise ValueError("num_gt_bboxes entry outside [0, MAX_NUM_BBOXES]")
    if status & _lib.STATUS_AR_TIMEOUT:
        raise RuntimeError("fused loss all-reduce timed out: a peer rank did not launch its step")


def match_loss_raw(locations, confidences, gt_bboxes, num_gt, priors, alpha, flags=0,
                   want_mask=False, want_gt_idx=False, want_stacked=False, want_grads=True,
                   want_conf_out=False, warps=0, cols=0, out=None, peer=None):
    """Thin wrapper of ``mbx_match_loss`` (see include/multibox_b200.h).  Inputs
    must already be contiguous fp32/int32 CUDA tensors; locations [B,P,4],
    confidences [B,P].  Returns a dict of device tensors; nothing synchronises.
    `out` may carry preallocated output tensors (same keys) to avoid allocation.
    `peer` (multibox_b200.dist.PeerAllreduce) fuses the cross-GPU SUM of the two
    losses into the kernel; results[8:12].view(float64) then holds the global sums."""
    lib = _lib.load()
    B, P = locations.shape[0], locations.shape[1]
    M = gt_bboxes.shape[1]
    dev = _same_device(locations, confidences, gt_bboxes, num_gt, priors)
    out = {} if out is None else out

    def buf(name, want, shape, dtype):
        if not want:
            return None
        t = out.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t

    mask = buf("mask", want_mask, (B * P,), torch.int32)
    gt_idx = buf("matched_gt_idx", want_gt_idx, (B * P,), torch.int32)
    stacked = buf("stacked_gt", want_stacked, (max(B * M, 1), 4), torch.float32)
    n_stacked = buf("n_stacked", want_stacked, (1,), torch.int32)
    d_loc = buf("d_locations", want_grads, (B, P, 4), torch.float32)
    d_conf = buf("d_confidences", want_grads, (B, P, 1), torch.float32)
    conf_out = buf("confidences", want_conf_out, (B, P, 1), torch.float32)
    results = buf("results", True, (_lib.RESULT_WORDS,), torch.float32)
    nbytes = lib.mbx_match_workspace_bytes(B, P, M)
    ws = _workspace(dev, nbytes)
    flags = int(flags) | (int(warps) << _lib.FLAG_WARPS_SHIFT) | (int(cols) << _lib.FLAG_COLS_SHIFT)
    args = (_lib.ptr(locations), _lib.ptr(confidences), _lib.ptr(gt_bboxes), _lib.ptr(num_gt), _lib.ptr(priors),
            B, P, M, float(alpha), flags,
            _lib.ptr(mask), _lib.ptr(gt_idx), _lib.ptr(stacked), _lib.ptr(n_stacked),
            _lib.ptr(d_loc), _lib.ptr(d_conf), _lib.ptr(conf_out), _lib.ptr(results),
            _lib.ptr(ws), ws.numel())
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        if peer is not None and peer.world > 1:
            rc = lib.mbx_match_loss_allreduce(*args, peer.ptr_array, peer.world, peer.rank, stream)
        else:
            rc = lib.mbx_match_loss(*args, stream)
    _lib.check(rc, "mbx_match_loss")
    return out


def compute_assignments(locations, confidences, gt_bboxes, num_gt_bboxes, batch_size, alpha,
                        return_indices=False):
    """Drop-in for reference loss.py:8-53 (the tf.py_func body).

    locations [B*P,4] (prior already added), confidences [B*P] (epsilon already
    added), gt_bboxes [B,M,4], num_gt_bboxes [B] int32, all CUDA tensors.
    Returns ``[assignment_partitions int32 [B*P], stacked_gt_bboxes f32 [N,4]]``
    (``+ [matched_gt_idx int32 [B*P]]`` with return_indices).  Raises ValueError
    where scipy's linear_sum_assignment would.  The result shape is data
    dependent, so this call synchronises (one 32-byte read-back)."""
    B = int(batch_size)
    loc = _f32c(locations, "locations")
    conf = _f32c(confidences, "confidences")
    P = loc.shape[0] // B                 # reference loss.py:16
    loc = loc.view(B, P, 4)
    conf = conf.view(B, P)
    gt = _f32c(gt_bboxes, "gt_bboxes")
    ng = _i32c(num_gt_bboxes, "num_gt_bboxes")
    out = match_loss_raw(loc, conf, gt, ng, None, alpha, flags=_lib.FLAG_BOUNDARY,
                         want_mask=True, want_gt_idx=return_indices, want_stacked=True,
                         want_grads=False)
    res = out["results"].cpu()
    raise_for_status(res[2].item())
    n = int(out["n_stacked"].item())
    ret = [out["mask"], out["stacked_gt"][:n]]
    if return_indices:
        ret.append(out["matched_gt_idx"])
    return ret


class _AddLoss(torch.autograd.Function):
    """Fused forward + backward: one kernel produces both losses and both
    gradients (the matching itself is non-differentiable, as the tf.py_func at
    reference loss.py:82 is)."""

    @staticmethod
    def forward(ctx, locations, confidences, batched_bboxes, batched_num_bboxes, bbox_priors,
                alpha, flags, validate):
        loc = _f32c(locations, "locations")
        B, P = loc.shape[0], loc.shape[1]
        conf = _f32c(confidences, "confidences").view(B, P)
        out = match_loss_raw(loc, conf, _f32c(batched_bboxes, "batched_bboxes"),
                             _i32c(batched_num_bboxes, "batched_num_bboxes"),
                             _f32c(bbox_priors, "bbox_priors"), alpha, flags=flags)
        res = out["results"]
        if validate:
            raise_for_status(res[2].item())
        ctx.save_for_backward(out["d_locations"], out["d_confidences"])
        ctx.conf_shape = confidences.shape
        ctx.mark_non_differentiable(res)
        return res[0], res[1], res

    @staticmethod
    def backward(ctx, g_loc, g_conf, _g_res):
        d_loc, d_conf = ctx.saved_tensors
        gl = d_loc * g_loc if g_loc is not None else None
        gc = (d_conf * g_conf).view(ctx.conf_shape) if g_conf is not None else None
        return gl, gc, None, None, None, None, None, None


def add_loss(locations, confidences, batched_bboxes, batched_num_bboxes, bbox_priors,
             location_loss_alpha, validate=True):
    """Drop-in for reference loss.py:55-117.

    locations [B,P,4] predicted offsets, confidences [B,P,1] post-sigmoid,
    batched_bboxes [B,M,4], batched_num_bboxes [B] int32, bbox_priors [P,4].
    Returns ``(location_loss, confidence_loss)`` as 0-d fp32 CUDA tensors (batch
    sums, reference loss.py:100-101) that back-propagate to `locations` and
    `confidences`.  With validate=True (default) the device status word is read
    back (one sync) and the ValueError of scipy's solver is reproduced."""
    loc_loss, conf_loss, _ = _AddLoss.apply(locations, confidences, batched_bboxes, batched_num_bboxes,
                                            bbox_priors, float(location_loss_alpha), 0, bool(validate))
    return loc_loss, conf_loss


def add_loss_from_logits(locations, logits, batched_bboxes, batched_num_bboxes, bbox_priors,
                         location_loss_alpha, validate=True):
    """Extension: same as add_loss but takes the pre-sigmoid head output and
    fuses reference model.py:322 into the kernel; gradients flow to the logits."""
    loc_loss, conf_loss, _ = _AddLoss.apply(locations, logits, batched_bboxes, batched_num_bboxes,
                                            bbox_priors, float(location_loss_alpha), _lib.FLAG_LOGITS,
                                            bool(validate))
    return loc_loss, conf_loss


def pad_ragged_gt(gt_flat, gt_row_offsets, max_num_bboxes):
    """The padded block the reference's input pipeline builds (inputs.py:340-348): [B,M,4] zero
    padded + counts [B].  Host/torch helper for tests and callers that need the padded form."""
    off = gt_row_offsets.to(torch.int64)
    B = off.numel() - 1
    counts = (off[1:] - off[:-1]).to(torch.int32)
    out = torch.zeros((B, int(max_num_bboxes), 4), dtype=torch.float32, device=gt_flat.device)
    for b in range(B):
        n = int(counts[b])
        out[b, :n] = gt_flat[int(off[b]):int(off[b]) + n]
    return out, counts


def match_loss_ragged_raw(locations, confidences, gt_flat, gt_row_offsets, priors, alpha, max_num_bboxes,
                          flags=0, want_mask=False, want_gt_idx=False, want_stacked=False, want_grads=True,
                          out=None, validate_offsets=False):
    """``mbx_match_loss_ragged``: ground truth as CSR (gt_flat [N,4] f32, gt_row_offsets [B+1] int32)
    instead of the zero-padded [B,M,4] block of reference inputs.py:340-348.  `max_num_bboxes` is
    the per-image capacity M.  Same outputs as match_loss_raw.  The kernel cannot check the offsets
    against N (it never sees N): `validate_offsets=True` checks them on the host first (one read-back);
    the autograd entry point add_loss_ragged does so when validate=True."""
    lib = _lib.load()
    B, P = locations.shape[0], locations.shape[1]
    M = int(max_num_bboxes)
    dev = _same_device(locations, confidences, gt_flat, gt_row_offsets, priors)
    out = {} if out is None else out
    if validate_offsets:
        off = gt_row_offsets.to(torch.int64).cpu()
        if off.numel() != B + 1 or int(off[0]) != 0 or int(off[-1]) != gt_flat.shape[0] or bool((off[1:] < off[:-1]).any()):
            raise ValueError("gt_row_offsets must be %d non-decreasing int32 values from 0 to N=%d"
                             % (B + 1, gt_flat.shape[0]))

    def buf(name, want, shape, dtype):
        if not want:
            return None
        t = out.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t

    N = gt_flat.shape[0]
    mask = buf("mask", want_mask, (B * P,), torch.int32)
    gt_idx = buf("matched_gt_idx", want_gt_idx, (B * P,), torch.int32)
    stacked = buf("stacked_gt", want_stacked, (max(N, 1), 4), torch.float32)
    n_stacked = buf("n_stacked", want_stacked, (1,), torch.int32)
    d_loc = buf("d_locations", want_grads, (B, P, 4), torch.float32)
    d_conf = buf("d_confidences", want_grads, (B, P, 1), torch.float32)
    results = buf("results", True, (_lib.RESULT_WORDS,), torch.float32)
    ws = _workspace(dev, lib.mbx_match_workspace_bytes(B, P, M))
    gt_arg = gt_flat if N > 0 else torch.zeros((1, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mbx_match_loss_ragged(_lib.ptr(locations), _lib.ptr(confidences), _lib.ptr(gt_arg),
                                       _lib.ptr(gt_row_offsets), _lib.ptr(priors), B, P, M, float(alpha), int(flags),
                                       _lib.ptr(mask), _lib.ptr(gt_idx), _lib.ptr(stacked), _lib.ptr(n_stacked),
                                       _lib.ptr(d_loc), _lib.ptr(d_conf), None, _lib.ptr(results),
                                       _lib.ptr(ws), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "mbx_match_loss_ragged")
    return out


class _AddLossRagged(torch.autograd.Function):
    @staticmethod
    def forward(ctx, locations, confidences, gt_flat, gt_row_offsets, bbox_priors, alpha, M, flags, validate):
        loc = _f32c(locations, "locations")
        B, P = loc.shape[0], loc.shape[1]
        conf = _f32c(confidences, "confidences").view(B, P)
        out = match_loss_ragged_raw(loc, conf, _f32c(gt_flat, "gt_flat"), _i32c(gt_row_offsets, "gt_row_offsets"),
                                    _f32c(bbox_priors, "bbox_priors"), alpha, M, flags=flags,
                                    validate_offsets=bool(validate))
        res = out["results"]
        if validate:
            raise_for_status(res[2].item())
        ctx.save_for_backward(out["d_locations"], out["d_confidences"])
        ctx.conf_shape = confidences.shape
        ctx.mark_non_differentiable(res)
        return res[0], res[1], res

    @staticmethod
    def backward(ctx, g_loc, g_conf, _g_res):
        d_loc, d_conf = ctx.saved_tensors
        gl = d_loc * g_loc if g_loc is not None else None
        gc = (d_conf * g_conf).view(ctx.conf_shape) if g_conf is not None else None
        return gl, gc, None, None, None, None, None, None, None


def add_loss_ragged(locations, confidences, gt_flat, gt_row_offsets, bbox_priors, location_loss_alpha,
                    max_num_bboxes, validate=True, logits=False):
    """add_loss (reference loss.py:55-117) with RAGGED ground truth: gt_flat [N,4] holds every
    image's boxes back to back, gt_row_offsets [B+1] int32 the row range of each image -- the
    format that replaces the MAX_NUM_BBOXES zero padding of reference inputs.py:340-348
    (SURVEY.md section 8 f4).  Bit-identical to add_loss on the padded equivalent."""
    loc_loss, conf_loss, _ = _AddLossRagged.apply(locations, confidences, gt_flat, gt_row_offsets, bbox_priors,
                                                  float(location_loss_alpha), int(max_num_bboxes),
                                                  _lib.FLAG_LOGITS if logits else 0, bool(validate))
    return loc_loss, conf_loss


# ----------------------------------------------------------------------------- head layout (model.py:295-322)
def head_priors(num_bboxes_per_cell, grids=(8, 6, 4, 3, 2, 1)):
    """Priors contributed by each detection head, in the reference's concatenation order
    (model.py:314-320): g*g*K for the five grids, 1 for the 1x1 head (model.py:281-287)."""
    return [g * g * (num_bboxes_per_cell if g > 1 else 1) for g in grids]


def concat_heads(head_locations, head_confidences):
    """The reference's own layout step (model.py:295-320, without the sigmoid): NHWC head outputs
    [B,g,g,K*4] / [B,g,g,K] -> locations [B,P,4], confidences [B,P,1].  Plain torch; used by the
    tests as the un-fused path the head-layout kernels must reproduce."""
    B = head_locations[0].shape[0]
    loc = torch.cat([t.reshape(B, -1) for t in head_locations], dim=1).reshape(B, -1, 4)
    conf = torch.cat([t.reshape(B, -1) for t in head_confidences], dim=1).reshape(B, -1, 1)
    return loc, conf


def make_heads_struct(head_locations, head_confidences, d_locations=None, d_confidences=None):
    """ctypes `mbx_heads` for a list of per-head CUDA tensors (kept alive by the caller)."""
    hs = _lib.Heads()
    n = len(head_locations)
    if not 1 <= n <= _lib.MAX_HEADS or len(head_confidences) != n:
        raise ValueError("need 1..%d heads, locations and confidences alike" % _lib.MAX_HEADS)
    hs.num_heads = n
    B = head_locations[0].shape[0]
    P = 0
    for k in range(n):
        hl, hc = head_locations[k], head_confidences[k]
        pri = hc.numel() // B
        if hl.numel() != B * pri * 4:
            raise ValueError("head %d: locations %s do not match confidences %s" % (k, tuple(hl.shape), tuple(hc.shape)))
        hs.head_priors[k] = pri
        hs.locations[k] = hl.data_ptr()
        hs.confidences[k] = hc.data_ptr()
        hs.d_locations[k] = d_locations[k].data_ptr() if d_locations is not None else None
        hs.d_confidences[k] = d_confidences[k].data_ptr() if d_confidences is not None else None
        P += pri
    return hs, B, P


def match_loss_heads_raw(head_locations, head_confidences, gt_bboxes, num_gt, priors, alpha, flags=0,
                         gt_row_offsets=None, max_num_bboxes=None, want_mask=False, want_gt_idx=False,
                         want_grads=True, want_conf_out=False, out=None):
    """``mbx_match_loss_heads``: the training step fed from the per-head conv outputs.  Returns a dict
    with per-head gradient lists ``d_head_locations`` / ``d_head_confidences`` (same shapes as the
    inputs).  `out` (a dict returned by an earlier call with the same shapes) re-uses the outputs."""
    lib = _lib.load()
    hl = [_f32c(t, "head_locations") for t in head_locations]
    hc = [_f32c(t, "head_confidences") for t in head_confidences]
    dev = _same_device(*(hl + hc + [gt_bboxes, num_gt, gt_row_offsets, priors]))
    out = {} if out is None else out
    dl = dc = None
    if want_grads:
        dl = out.get("d_head_locations") or [torch.empty_like(t) for t in hl]
        dc = out.get("d_head_confidences") or [torch.empty_like(t) for t in hc]
    hs, B, P = make_heads_struct(hl, hc, dl, dc)
    if P != priors.shape[0]:
        raise ValueError("heads hold %d priors, bbox_priors %d" % (P, priors.shape[0]))
    M = int(max_num_bboxes) if gt_row_offsets is not None else gt_bboxes.shape[1]

    def buf(name, want, shape, dtype):
        if not want:
            return None
        t = out.get(name)
        return t if t is not None else torch.empty(shape, dtype=dtype, device=dev)

    mask = buf("mask", want_mask, (B * P,), torch.int32)
    gt_idx = buf("matched_gt_idx", want_gt_idx, (B * P,), torch.int32)
    conf_out = buf("confidences", want_conf_out, (B, P, 1), torch.float32)
    results = buf("results", True, (_lib.RESULT_WORDS,), torch.float32)
    out.update(d_head_locations=dl, d_head_confidences=dc, mask=mask, matched_gt_idx=gt_idx, confidences=conf_out,
               results=results)
    ws = _workspace(dev, lib.mbx_match_workspace_bytes(B, P, M))
    with torch.cuda.device(dev):
        rc = lib.mbx_match_loss_heads(ctypes.byref(hs), _lib.ptr(gt_bboxes), _lib.ptr(num_gt),
                                      _lib.ptr(gt_row_offsets), _lib.ptr(priors), B, P, M, float(alpha), int(flags),
                                      _lib.ptr(mask), _lib.ptr(gt_idx), None, None, _lib.ptr(conf_out),
                                      _lib.ptr(results), _lib.ptr(ws), ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "mbx_match_loss_heads")
    return out


class _AddLossHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt, num_gt, priors, alpha, flags, validate, nheads, *heads):
        hl, hc = heads[:nheads], heads[nheads:]
        out = match_loss_heads_raw(hl, hc, _f32c(gt, "batched_bboxes"), _i32c(num_gt, "batched_num_bboxes"),
                                   _f32c(priors, "bbox_priors"), alpha, flags=flags)
        res = out["results"]
        if validate:
            raise_for_status(res[2].item())
        ctx.nheads = nheads
        ctx.shapes = [t.shape for t in heads]
        ctx.save_for_backward(*(out["d_head_locations"] + out["d_head_confidences"]))
        ctx.mark_non_differentiable(res)
        return res[0], res[1], res

    @staticmethod
    def backward(ctx, g_loc, g_conf, _g_res):
        n = ctx.nheads
        saved = ctx.saved_tensors
        grads = []
        for k, t in enumerate(saved):
            g = g_loc if k < n else g_conf
            grads.append((t * g).view(ctx.shapes[k]) if g is not None else None)
        return (None, None, None, None, None, None, None) + tuple(grads)


def add_loss_from_heads(head_locations, head_logits, batched_bboxes, batched_num_bboxes, bbox_priors,
                        location_loss_alpha, validate=True, logits=True):
    """add_loss fed straight from the detection heads: `head_locations[h]`
    / `head_logits[h]` are the NHWC conv outputs of head h ([B,g,g,K*4] / [B,g,g,K], reference
    model.py:213-293); the reshape + concat + sigmoid of model.py:295-322 happen inside the kernel
    and the gradients come back per head, so the concatenated [B,P,5] tensor and its gradient are
    never written to HBM.  With logits=False the confidences are taken as already-sigmoided."""
    n = len(head_locations)
    loc_loss, conf_loss, _ = _AddLossHeads.apply(batched_bboxes, batched_num_bboxes, bbox_priors,
                                                 float(location_loss_alpha), _lib.FLAG_LOGITS if logits else 0,
                                                 bool(validate), n, *(list(head_locations) + list(head_logits)))
    return loc_loss, conf_loss


class _PlanHolder:
    """Owns one mbx_match_plan (destroyed with the launch closure that uses it)."""

    def __init__(self, lib, plan):
        self._lib, self._plan = lib, plan

    def __del__(self):
        try:
            if self._plan:
                self._lib.mbx_match_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass


class MultiboxLossStep:
    """Allocation-free training-step object: preallocated outputs, one kernel
    launch per step, and a host-buffer path for callers that hold HOST arrays
    (the reference's tf.py_func boundary hands numpy arrays over).

    step(...)        device tensors in, device results out (no sync)
    step_host(...)   host numpy arrays in -> losses out: ONE packed pinned H2D copy,
                     one kernel, one 32-byte D2H read-back, optionally replayed as a
                     CUDA graph (use_graph=True) so a step costs one graph launch.
    Inputs are staged in a single contiguous buffer laid out as
    [locations B*P*4 | confidences B*P | gt B*M*4 | num_gt B] (fp32 / int32).
    """

    def __init__(self, B, P, M, priors, alpha, device="cuda", logits=False, want_mask=False,
                 want_stacked=False, warps=0, use_graph=False, peer=None, deferred_allreduce=False,
                 host_results=False, zero_copy=False, pdl=False, static_schedule=False, own_stream=False):
        self.B, self.P, self.M, self.alpha = B, P, M, float(alpha)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.flags = _lib.FLAG_LOGITS if logits else 0
        # pdl: programmatic dependent launch (MBX_FLAG_PDL).  The caller promises that the step's INPUTS are
        # not produced by the kernel that precedes the step on the stream (true for the pinned host staging
        # of this object, and for device tensors produced earlier): consecutive steps then overlap -- the
        # load, logs and solve of step k+1 run while step k finishes -- with identical results.
        if pdl:
            self.flags |= _lib.FLAG_PDL
        # static_schedule (MBX_FLAG_STATIC): image -> CTA assignment by index instead of the heavy-first
        # dynamic order when the batch exceeds the resident CTAs: no order kernel in front of the step, which
        # also lets pdl overlap consecutive steps at such sizes; right for batches of similar images.
        if static_schedule:
            self.flags |= _lib.FLAG_STATIC
        self.warps = warps
        self.priors = _f32c(torch.as_tensor(priors).to(self.device), "priors")
        self.want_mask, self.want_stacked = want_mask, want_stacked
        self.peer = peer if (peer is not None and peer.world > 1) else None
        if self.peer is not None and deferred_allreduce:
            self.flags |= _lib.FLAG_AR_DEFERRED
        self.out = {}
        # packed staging: one pinned host buffer, one device buffer, typed views into both
        def up4(x):          # every section starts 16-byte aligned
            return (x + 3) // 4 * 4
        o_conf = up4(B * P * 4)
        o_gt = up4(o_conf + B * P)
        o_ng = up4(o_gt + B * M * 4)
        self._sections = ((0, B * P * 4), (o_conf, B * P), (o_gt, B * M * 4), (o_ng, B))
        words = up4(o_ng + B)
        self.h_in = torch.empty((words,), dtype=torch.float32).pin_memory()
        self.d_in = torch.empty((words,), dtype=torch.float32, device=self.device)
        self.h_loc, self.h_conf, self.h_gt, self.h_ng = self._views(self.h_in)
        self.d_loc_in, self.d_conf_in, self.d_gt_in, self.d_ng_in = self._views(self.d_in)
        self.h_res = torch.empty((_lib.RESULT_WORDS,), dtype=torch.float32).pin_memory()
        self.h2d_bytes = 4 * words
        self.d2h_bytes = 4 * _lib.RESULT_WORDS
        self.use_graph = bool(use_graph)
        self._graph = None
        self._launch = None
        # host_results: the kernel stores the 64-byte result block straight into mapped pinned host
        # memory and the host polls the launch sequence word (results[15]) -- no D2H copy node, no
        # stream synchronisation on the step's critical path.  Host-buffer path only.
        self.host_results = bool(host_results)
        # zero_copy: the kernel reads the packed inputs straight from the MAPPED PINNED staging buffer
        # over PCIe (its streaming, read-once 16-byte loads are the host->device transfer): no H2D
        # copy node in front of the kernel.  Host-buffer path only.
        self.zero_copy = bool(zero_copy)
        if self.host_results:
            self.flags |= _lib.FLAG_HOST_RESULTS
            self.out["results"] = self.h_res
            self._res_u32 = self.h_res.numpy().view("uint32")
            self._res_f32 = self.h_res.numpy()
            self._res_u32[15] = 0
        # own_stream: the object's host-buffer steps run on a CUDA stream of its own -- ONE foreign call
        # enqueues the H2D copy of the packed pinned inputs (copy engine, full PCIe rate) and the kernel
        # (mbx_match_plan_launch_staged).  A caller that rotates a few such objects overlaps the copy of step
        # k+1 with the kernel of step k without any event plumbing.  Host-buffer path with host_results only.
        # With a fused all-reduce every object needs its OWN PeerAllreduce (steps on different streams must
        # not share one exchange state).
        self.stream = None
        if own_stream:
            if not self.host_results or self.zero_copy or use_graph:
                raise ValueError("own_stream needs host_results=True, zero_copy=False, use_graph=False")
            self.flags &= ~_lib.FLAG_PDL          # (the kernel follows a copy on its stream, not a kernel)
            self.stream = torch.cuda.Stream(device=self.device)
            self._staged = None

    def _views(self, buf):
        (a0, n0), (a1, n1), (a2, n2), (a3, n3) = self._sections
        B, P, M = self.B, self.P, self.M
        return (buf[a0:a0 + n0].view(B, P, 4), buf[a1:a1 + n1].view(B, P), buf[a2:a2 + n2].view(B, M, 4),
                buf[a3:a3 + n3].view(torch.int32))

    def step(self, locations, confidences, gt, num_gt):
        return match_loss_raw(locations, confidences.view(self.B, self.P), gt, num_gt, self.priors,
                              self.alpha, flags=self.flags, want_mask=self.want_mask,
                              want_gt_idx=self.want_mask, want_stacked=self.want_stacked,
                              want_grads=True, warps=self.warps, out=self.out, peer=self.peer)

    def prepare(self, locations, confidences, gt, num_gt, inputs=None):
        """Returns a zero-argument callable that launches the step on these (fixed)
        device tensors with all ctypes arguments pre-marshalled: the launch costs a
        single foreign call (for latency-critical loops and CUDA-graph capture).
        `inputs` (four tensors) replaces the input POINTERS of the closure after the outputs were
        set up with the device tensors -- used for mapped pinned host inputs (zero_copy)."""
        import ctypes as c
        lib = _lib.load()
        out = self.step(locations, confidences, gt, num_gt)     # allocates outputs / workspace once
        if inputs is not None:
            torch.cuda.current_stream(self.device).synchronize()
            locations, confidences, gt, num_gt = inputs
        B, P, M = self.B, self.P, self.M
        ws = _workspace(self.device, lib.mbx_match_workspace_bytes(B, P, M))
        flags = int(self.flags) | (int(self.warps) << _lib.FLAG_WARPS_SHIFT)
        keep = (locations, confidences, gt, num_gt, ws, out)    # keep the tensors alive
        # a prepared launch (mbx_match_plan): the argument list is marshalled ONCE; a step then costs a
        # two-argument foreign call + one kernel launch (the step at batch 32 is shorter than marshalling
        # 24 ctypes arguments and building a torch Stream object)
        plan = c.c_void_p()
        rc = lib.mbx_match_plan_create(
            c.byref(plan), locations.data_ptr(), confidences.data_ptr(), gt.data_ptr(), num_gt.data_ptr(),
            self.priors.data_ptr(), B, P, M, c.c_float(self.alpha), c.c_uint(flags),
            _lib.ptr(out.get("mask")), _lib.ptr(out.get("matched_gt_idx")), _lib.ptr(out.get("stacked_gt")),
            _lib.ptr(out.get("n_stacked")), out["d_locations"].data_ptr(), out["d_confidences"].data_ptr(), None,
            out["results"].data_ptr(), ws.data_ptr(), ws.numel(),
            self.peer.ptr_array if self.peer is not None else None,
            self.peer.world if self.peer is not None else 1, self.peer.rank if self.peer is not None else 0)
        _lib.check(rc, "mbx_match_plan_create")
        holder = _PlanHolder(lib, plan)
        dev = self.device
        dev_index = dev.index
        cur_dev, set_dev = torch.cuda.current_device, torch.cuda.set_device
        raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        plan_launch, plan_ptr = lib.mbx_match_plan_launch, plan.value

        def launch(_keep=(keep, holder)):
            prev = cur_dev()
            if prev != dev_index:        # the library launches on the CUDA current device
                set_dev(dev_index)
            try:
                st = raw_stream(dev_index) if raw_stream is not None else torch.cuda.current_stream(dev).cuda_stream
                rc = plan_launch(plan_ptr, st)
            finally:
                if prev != dev_index:
                    set_dev(prev)
            if rc:
                _lib.check(rc, "mbx_match_loss")
        launch.plan_ptr = plan.value
        return launch

    def _enqueue_host_step(self):
        if self.stream is not None:
            rc = self._staged[0](*self._staged[1])               # H2D copy + kernel on the own stream
            if rc:
                _lib.check(rc, "mbx_match_plan_launch_staged")
            return
        if not self.zero_copy:
            self.d_in.copy_(self.h_in, non_blocking=True)       # one H2D copy of the packed inputs
        self._launch()                                           # one kernel (zero_copy: reads h_in itself)
        if not self.host_results:
            self.h_res.copy_(self.out["results"], non_blocking=True)  # losses + status (64 bytes)

    def _wait_host_results(self):
        """Spins on the launch sequence word the kernel stores last into the mapped result block."""
        import time
        u = self._res_u32
        spins = 0
        while u[15] == 0:
            spins += 1
            if spins & 0xffff == 0 and spins > (1 << 22):
                t0 = time.perf_counter()
                (self.stream or torch.cuda.current_stream(self.device)).synchronize()   # something is wrong: fall back
                if u[15] == 0:
                    raise RuntimeError("MultiboxLossStep: the kernel finished without publishing its results "
                                       "(%.3f s)" % (time.perf_counter() - t0))

    def _ensure_ready(self):
        if self._launch is None and self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):       # (the workspace is keyed by the stream)
                self.d_in.copy_(self.h_in, non_blocking=True)
                self._launch = self.prepare(self.d_loc_in, self.d_conf_in, self.d_gt_in, self.d_ng_in)
            self.stream.synchronize()
            lib = _lib.load()
            self._staged = (lib.mbx_match_plan_launch_staged,
                            (self._launch.plan_ptr, self.h_in.data_ptr(), self.d_in.data_ptr(),
                             self.h2d_bytes, self.stream.cuda_stream))
        if self._launch is None:
            host = (self.h_loc, self.h_conf, self.h_gt, self.h_ng) if self.zero_copy else None
            self._launch = self.prepare(self.d_loc_in, self.d_conf_in, self.d_gt_in, self.d_ng_in, inputs=host)
            torch.cuda.current_stream(self.device).synchronize()
        if self.use_graph and self._graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._enqueue_host_step()                        # warm-up on the capture stream
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._enqueue_host_step()
            self._graph = g

    def global_losses(self):
        """(location_loss, confidence_loss) summed over all ranks, as read back by the last
        step_host / step_pinned / flush (fused all-reduce; equals the local losses on one GPU).
        With deferred_allreduce they belong to step `global_step()` (the previous one)."""
        g = self.h_res[8:12].view(torch.float64)
        return float(g[0]), float(g[1])

    def global_step(self):
        return int(self.h_res[14].item())

    def flush(self):
        """Deferred all-reduce: completes the newest step's reduction and reads it back."""
        if self.peer is None:
            return self.global_losses()
        lib = _lib.load()
        import contextlib
        with (torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()):
            ws = _workspace(self.device, lib.mbx_match_workspace_bytes(self.B, self.P, self.M))
            rc = lib.mbx_allreduce_flush(_lib.ptr(self.out["results"]), _lib.ptr(ws), ws.numel(), self.peer.ptr_array,
                                         self.peer.world, self.peer.rank,
                                         torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(rc, "mbx_allreduce_flush")
            if self.out["results"] is not self.h_res:
                self.h_res.copy_(self.out["results"], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        raise_for_status(self.h_res[2].item())
        return self.global_losses()

    def step_host(self, locations, confidences, gt, num_gt, validate=True):
        """numpy in -> (location_loss, confidence_loss) python floats out; the
        gradients stay on the device in self.out (they feed the network's backward)."""
        import numpy as np
        np.copyto(self.h_loc.numpy(), locations.reshape(self.B, self.P, 4))
        np.copyto(self.h_conf.numpy(), confidences.reshape(self.B, self.P))
        np.copyto(self.h_gt.numpy(), gt)
        np.copyto(self.h_ng.numpy(), num_gt)
        return self.step_pinned(validate)

    def submit_pinned(self):
        """Enqueues the step on the data currently in the pinned staging buffer and returns at once;
        `wait()` later returns its losses.  With two step objects a caller keeps TWO steps in flight:
        stage + submit step k+1 while the GPU runs step k, then wait for k (bench.py's e2e loop)."""
        self._ensure_ready()
        if self.host_results:
            self._res_u32[15] = 0
        if self._graph is not None:
            self._graph.replay()
        else:
            self._enqueue_host_step()

    def wait(self, validate=True):
        """Completes the step enqueued by submit_pinned: (location_loss, confidence_loss)."""
        if self.host_results:
            self._wait_host_results()
            r = self._res_f32
            if validate and r[2] != 0.0:
                raise_for_status(float(r[2]))
            return float(r[0]), float(r[1])
        torch.cuda.current_stream(self.device).synchronize()
        if validate:
            raise_for_status(self.h_res[2].item())
        return float(self.h_res[0]), float(self.h_res[1])

    def step_pinned(self, validate=True, pinned=None):
        """Same as step_host when the caller already wrote into the pinned staging
        buffer (`self.h_in`, or another packed pinned buffer passed as `pinned`)."""
        self._ensure_ready()
        if pinned is not None and pinned is not self.h_in:
            # the launch closure and the graph are bound to this object's own staging buffer (the kernel
            # reads h_in itself with zero_copy, a copy node reads it otherwise): bring the caller's data
            # there, in every (use_graph, zero_copy) mode
            if pinned.numel() != self.h_in.numel() or pinned.dtype != self.h_in.dtype:
                raise ValueError("pinned must be a packed float32 buffer of %d words" % self.h_in.numel())
            self.h_in.copy_(pinned)
        self.submit_pinned()
        return self.wait(validate)
