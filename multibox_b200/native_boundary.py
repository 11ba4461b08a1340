"""The literal ``tf.py_func`` drop-in: host numpy in, host numpy out.

Reference boundary (``loss.py:81-82``)::

    params = [locations, confidences, batched_bboxes, batched_num_bboxes, batch_size, location_loss_alpha]
    assignment_partitions, stacked_gt_bboxes = tf.py_func(compute_assignments, params, [tf.int32, tf.float32])

``compute_assignments`` below has exactly that contract (reference ``loss.py:8-53``): it receives HOST
numpy arrays -- locations [B*P,4] float32 (prior already added), confidences [B*P] float32 (epsilon
already added), gt_bboxes [B,M,4] float32 zero padded, num_gt_bboxes [B] int32, batch_size and alpha as
0-d arrays or python scalars -- and returns ``[int32 [B*P] mask, float32 [N,4] stacked_gt]`` as numpy
arrays, raising the ``ValueError`` scipy's ``linear_sum_assignment`` raises at ``loss.py:40``.  A
reference maintainer swaps the callable handed to ``tf.py_func`` and nothing else.

Everything is computed by ``mbx_match_loss`` (MBX_FLAG_BOUNDARY) on the GPU; this module only stages the
arrays through pinned host memory (one packed H2D copy, one packed D2H copy) and keeps the staging
buffers of the last shape alive between calls (the py_func is called once per training step with a fixed
shape).  There is no CPU path: without a CUDA device the call raises.
"""
import threading

import numpy as np
import torch

from . import _lib
from .loss import _workspace, raise_for_status

_local = threading.local()


class _Staging:
    """Pinned + device buffers for one (B, P, M): inputs packed as [loc | conf | gt | num_gt], outputs as
    [mask | n_stacked, results(16) | stacked]."""

    def __init__(self, B, P, M, device):
        self.key = (B, P, M, device.index)
        up4 = lambda x: (x + 3) // 4 * 4      # noqa: E731  (16-byte aligned sections)
        self.o_conf = up4(B * P * 4)
        self.o_gt = up4(self.o_conf + B * P)
        self.o_ng = up4(self.o_gt + B * M * 4)
        n_in = up4(self.o_ng + B)
        self.h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
        self.d_in = torch.empty(n_in, dtype=torch.float32, device=device)
        self.o_meta = up4(B * P)
        self.o_stk = self.o_meta + 32
        n_out = self.o_stk + max(B * M, 1) * 4
        self.h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
        self.d_out = torch.empty(n_out, dtype=torch.float32, device=device)
        self.np_in = self.h_in.numpy()
        self.np_out = self.h_out.numpy()


def _staging(B, P, M, device):
    st = getattr(_local, "staging", None)
    if st is None or st.key != (B, P, M, device.index):
        st = _Staging(B, P, M, device)
        _local.staging = st
    return st


def compute_assignments(locations, confidences, gt_bboxes, num_gt_bboxes, batch_size, alpha, device=None):
    """Drop-in for the body of the reference's ``tf.py_func`` (``loss.py:8-53``); see the module docstring."""
    if not torch.cuda.is_available():
        raise RuntimeError("multibox_b200 has no CPU path: compute_assignments needs a CUDA device")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = int(batch_size)
    loc = np.asarray(locations, dtype=np.float32)
    conf = np.asarray(confidences, dtype=np.float32).reshape(-1)
    gt = np.asarray(gt_bboxes, dtype=np.float32)
    ng = np.asarray(num_gt_bboxes, dtype=np.int32).reshape(-1)
    if B <= 0 or loc.ndim != 2 or loc.shape[1] != 4 or loc.shape[0] % B or gt.ndim != 3 or gt.shape[0] != B \
            or gt.shape[2] != 4 or ng.shape[0] != B or conf.shape[0] != loc.shape[0]:
        raise ValueError("compute_assignments: expected locations [B*P,4], confidences [B*P], gt_bboxes [B,M,4], "
                         "num_gt_bboxes [B] (got %s, %s, %s, %s, batch_size=%d)"
                         % (loc.shape, conf.shape, gt.shape, ng.shape, B))
    P = loc.shape[0] // B                # reference loss.py:16
    M = gt.shape[1]
    st = _staging(B, P, M, dev)
    a = st.np_in
    a[:B * P * 4] = loc.reshape(-1)
    a[st.o_conf:st.o_conf + B * P] = conf
    a[st.o_gt:st.o_gt + B * M * 4] = gt.reshape(-1)
    a[st.o_ng:st.o_ng + B].view(np.int32)[:] = ng
    d = st.d_in
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        d.copy_(st.h_in, non_blocking=True)
        o = st.d_out
        ws = _workspace(dev, lib.mbx_match_workspace_bytes(B, P, M))
        base_in, base_out = d.data_ptr(), o.data_ptr()
        rc = lib.mbx_match_loss(base_in, base_in + 4 * st.o_conf, base_in + 4 * st.o_gt, base_in + 4 * st.o_ng, None,
                                B, P, M, float(alpha), _lib.FLAG_BOUNDARY,
                                base_out, None, base_out + 4 * st.o_stk, base_out + 4 * st.o_meta,
                                None, None, None, base_out + 4 * (st.o_meta + 16),
                                ws.data_ptr(), ws.numel(), stream.cuda_stream)
        _lib.check(rc, "mbx_match_loss")
        st.h_out.copy_(o, non_blocking=True)
        stream.synchronize()
    out = st.np_out
    raise_for_status(out[st.o_meta + 16 + 2])
    n = int(out[st.o_meta:st.o_meta + 1].view(np.int32)[0])
    mask = out[:B * P].view(np.int32).copy()
    stacked = out[st.o_stk:st.o_stk + 4 * n].reshape(n, 4).copy()
    return [mask, stacked]
