"""A/B of programmatic dependent launch (MBX_FLAG_PDL) on back-to-back training steps (GPU box):
python profiles/pdl_ab.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multibox_b200 import synth  # noqa: E402

for name, cfg, steps in (("configs[1] B=32", dict(synth.TRAIN_CONFIGS["cfg2"]), 400),
                         ("K=5 B=148 M=20", dict(K=5, B=148, M=20, seed=3), 200),
                         ("K=7 B=128 M=100 coco (configs[3] / 8 GPUs)", dict(K=7, B=128, M=100, dist="coco_person", seed=1004), 200)):
    d = synth.make_train_inputs(**cfg)
    for pdl in (False, True, False, True):
        tr = bench.bench_train(d, steps, 10, 1, lambda: None, None, want_e2e=False, pdl=pdl)
        print("%-44s pdl=%d  %.2f us/step  (%.3g img/s)" % (name, pdl, 1e6 * tr["sec"] / steps, d["B"] * steps / tr["sec"]))

# batches beyond the resident CTAs: heavy-first dynamic scheduling (order kernel in front, no overlap) against
# the static image -> CTA assignment with programmatic dependent launch (consecutive steps overlap)
for name, cfg, steps in (("configs[3] shape K=7 B=1024 M=100", dict(synth.TRAIN_CONFIGS["cfg4"]), 50),
                         ("K=5 B=4096 M=20", dict(K=5, B=4096, M=20, seed=3), 20),
                         ("configs[4] shape K=11 B=1024 M=200", dict(K=11, B=1024, M=200, dist="uniform", seed=1005), 10)):
    d = synth.make_train_inputs(**cfg)
    for static, pdl in ((False, False), (False, True), (True, True), (False, False), (False, True)):
        tr = bench.bench_train(d, steps, 5, 1, lambda: None, None, want_e2e=False, pdl=pdl, static_schedule=static)
        print("%-40s static=%d pdl=%d  %.2f us/step  (%.3g img/s)" %
              (name, static, pdl, 1e6 * tr["sec"] / steps, d["B"] * steps / tr["sec"]))
