"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle)."""
import ctypes

import numpy as np
import torch

from multibox_b200 import _lib


def dev(t, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(t))
    if dtype is not None:
        x = x.to(dtype)
    return x.cuda()


def boundary_inputs(d):
    """What crosses the reference's py_func boundary (loss.py:67-74)."""
    B = d["B"]
    loc = d["locations"].reshape(-1, 4) + np.tile(d["priors"], (B, 1))
    conf = d["confidences"].reshape(-1) + np.float32(1e-10)
    return loc.astype(np.float32), conf.astype(np.float32)


def gpu_nplog(x):
    lib = _lib.load()
    t = dev(x)
    out = torch.empty_like(t)
    _lib.check(lib.mbx_debug_nplog(t.data_ptr(), out.data_ptr(), ctypes.c_longlong(t.numel()),
                                   torch.cuda.current_stream().cuda_stream), "mbx_debug_nplog")
    return out.cpu().numpy()


def gpu_cost_matrix(loc, conf, gt, alpha):
    lib = _lib.load()
    P, n = loc.shape[0], gt.shape[0]
    C = torch.empty((P, n), dtype=torch.float64, device="cuda")
    l, c, g = dev(loc), dev(conf), dev(gt)
    _lib.check(lib.mbx_debug_cost_matrix(l.data_ptr(), c.data_ptr(), g.data_ptr(), P, n, float(alpha),
                                         C.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "mbx_debug_cost_matrix")
    return C.cpu().numpy()
