"""TEST INFRASTRUCTURE -- ctypes view of oracle/c/mbx_oracle.c (the plain-C
restatement of reference loss.py:8-53 + scipy's LSAP + numpy's fp32 log).
Checker only; never imported by multibox_b200."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmbx_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "c", "mbx_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_nplogf.restype = ctypes.c_float
        _lib.orc_nplogf.argtypes = [ctypes.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nplog(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_nplogf_array(_p(x), _p(out), ctypes.c_int64(x.size))
    return out


def lsap(cost):
    """scipy.optimize.linear_sum_assignment(cost) restated; raises ValueError alike."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    k = min(nr, nc)
    ri = np.zeros(k, dtype=np.int64)
    ci = np.zeros(k, dtype=np.int64)
    rc = lib().orc_lsap(ctypes.c_int64(nr), ctypes.c_int64(nc), _p(cost), _p(ri), _p(ci))
    if rc == 1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == 2:
        raise ValueError("cost matrix is infeasible")
    return ri, ci


def cost_matrix(loc, conf, gt, alpha):
    loc = np.ascontiguousarray(loc, np.float32)
    conf = np.ascontiguousarray(conf, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    P, n = loc.shape[0], gt.shape[0]
    C = np.zeros((P, n), dtype=np.float64)
    lib().orc_cost_matrix(_p(loc), _p(conf), _p(gt), ctypes.c_int64(P), ctypes.c_int64(n),
                          ctypes.c_float(alpha), _p(C))
    return C


def compute_assignments(locations, confidences, gt_bboxes, num_gt_bboxes, batch_size, alpha):
    """Same contract as np_oracle.compute_assignments(return_indices=True)."""
    loc = np.ascontiguousarray(locations, np.float32)
    conf = np.ascontiguousarray(confidences, np.float32)
    gt = np.ascontiguousarray(gt_bboxes, np.float32)
    ng = np.ascontiguousarray(num_gt_bboxes, np.int32)
    B = int(batch_size)
    P = loc.shape[0] // B
    M = gt.shape[1]
    mask = np.zeros(B * P, np.int32)
    gidx = np.zeros(B * P, np.int32)
    stacked = np.zeros((int(ng.sum()) + 1, 4), np.float32)
    ns = ctypes.c_int64(0)
    rc = lib().orc_compute_assignments(_p(loc), _p(conf), _p(gt), _p(ng), ctypes.c_int64(B),
                                       ctypes.c_int64(P), ctypes.c_int64(M), ctypes.c_float(alpha),
                                       _p(mask), _p(gidx), _p(stacked), ctypes.byref(ns))
    if rc == 1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == 2:
        raise ValueError("cost matrix is infeasible")
    return [mask, stacked[:ns.value].copy(), gidx]
