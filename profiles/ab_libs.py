"""A/B timing of two builds of the library on the same box (GPU box):
python profiles/ab_libs.py libA.so libB.so   -- times mbx_match_loss on configs[1] (B=32) and a
B=256 batch with each library, interleaved, inputs rotating over > L2 worth of copies."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import synth  # noqa: E402

vp, ci, cu, cf, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_size_t


def load(path):
    lib = ctypes.CDLL(path)
    lib.mbx_match_loss.restype = ci
    lib.mbx_match_loss.argtypes = [vp] * 5 + [ci, ci, ci, cf, cu] + [vp] * 8 + [vp, sz, vp]
    lib.mbx_match_workspace_bytes.restype = sz
    lib.mbx_match_workspace_bytes.argtypes = [ci, ci, ci]
    return lib


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


libs = [(os.path.basename(p), load(p)) for p in sys.argv[1:]]
for label, cfg in (("configs[1] B=32", dict(K=5, B=32, M=20, seed=1002)), ("B=256", dict(K=5, B=256, M=20, seed=3)),
                   ("cfg4 B=1024", dict(K=7, B=1024, M=100, dist="coco_person", seed=1004)),
                   ("big K=11 B=1024", dict(K=11, B=1024, M=200, dist="uniform", seed=1005))):
    d = synth.make_train_inputs(**cfg)
    B, P, M = d["B"], d["P"], d["M"]
    nsets = max(2, min(256, (300 << 20) // (B * P * 20)))
    loc = [dev(np.roll(d["locations"], r, 0)) for r in range(nsets)]
    conf = [dev(np.roll(d["confidences"], r, 0)) for r in range(nsets)]
    gt = [dev(np.roll(d["gt"], r, 0)) for r in range(nsets)]
    ng = [dev(np.roll(d["num_gt"], r, 0)) for r in range(nsets)]
    pri = dev(d["priors"])
    dl = torch.empty((B, P, 4), device="cuda")
    dc = torch.empty((B, P), device="cuda")
    res = torch.empty(16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, lib in libs:
        ws = torch.zeros(lib.mbx_match_workspace_bytes(B, P, M), dtype=torch.uint8, device="cuda")

        def run(i):
            s = i % nsets
            rc = lib.mbx_match_loss(loc[s].data_ptr(), conf[s].data_ptr(), gt[s].data_ptr(), ng[s].data_ptr(),
                                    pri.data_ptr(), B, P, M, 1000.0, 0, None, None, None, None, dl.data_ptr(),
                                    dc.data_ptr(), None, res.data_ptr(), ws.data_ptr(), ws.numel(), st)
            assert rc == 0
        out[name] = run
    for rep in range(3):
        for name, _ in libs:
            run = out[name]
            for i in range(20):
                run(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nrep = 200 if B * P < 1000000 else 20
            for i in range(nrep):
                run(20 + i)
            b.record()
            torch.cuda.synchronize()
            print("%-18s %-26s rep %d: %.2f us per back-to-back launch" % (label, name, rep, a.elapsed_time(b) * 1e3 / nrep))
