"""Builds a -DMBX_PHASE_TIMING copy of the library into gpurun_out/ and prints the per-phase cycle
totals of the register-resident matching kernel (GPU box).  python profiles/phase_timing.py [warps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multibox_b200 import _build, _lib, synth  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
# built next to the product library (a git-ignored .so travels to the GPU box; gpurun_out/ does not),
# so `python profiles/phase_timing.py --build-only` here saves the compile time on the box
so = os.path.join(ROOT, "multibox_b200", "libmbx_timing.so")
_src = [os.path.join(_build.CSRC, f) for f in os.listdir(_build.CSRC) if f.endswith((".cu", ".cuh"))]
if not os.path.isfile(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in _src):
    _build.build(force=True, extra_flags=("-DMBX_PHASE_TIMING",), lib=so)
if "--build-only" in sys.argv:
    sys.exit(0)
_build.LIB = so
_build.needs_build = lambda: False
lib = _lib.load()
from multibox_b200 import loss  # noqa: E402

detect_mode = "--detect" in sys.argv
if detect_mode:
    sys.argv.remove("--detect")
warps = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


names = ["prologue", "1st step pass 1 (cheap)", "sequential rows", "general scan", "general argmin", "select/log/walk",
         "dual update / exit", "epilogue", "1st step pass 2 (exact)"]
NS = 12      # slots per warp: 9 phase accumulators, exact-cost evaluations of the warp, 2 global timestamps


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


if detect_mode:
    from multibox_b200 import detect
    dn = ["load/decode/keys", "select + sort", "NMS: chunk triangles", "NMS: barrier waits (rest)", "store",
          "NMS: serial resolve", "NMS: wait for resolve", "NMS: cross-chunk"]
    for label, kw in (("cfg3", dict(K=5, B=148, keep=200, seed=1003)), ("K=11", dict(K=11, B=148, keep=200, seed=4))):
        q = synth.make_detect_inputs(**kw)
        B = q["B"]
        out = {"scores": torch.zeros((B, max(200, 8 * 8 * 2 * 2)), dtype=torch.float32, device="cuda")}
        for _ in range(2):
            detect.postprocess(dev(q["locations"]), dev(q["confidences"]), dev(q["priors"]),
                               restrictions=dev(q["restrictions"]), max_to_keep=dev(q["max_to_keep"]),
                               offsets=dev(q["offsets"]), patch_dims=dev(q["patch_dims"]),
                               image_dims=dev(q["image_dims"]), is_flipped=dev(q["is_flipped"]), nms_iou=0.5,
                               k_max=200, out=out)
        torch.cuda.synchronize()
        t = out["scores"].cpu().numpy().reshape(-1).view(np.int64)[:B * 8 * 8].reshape(B, 8, 8)
        print("detect %s: per-warp mean cycles by phase (block 0; warp 0 / other warps)" % label)
        for k, nm in enumerate(dn):
            print("   %-20s %9.0f %9.0f" % (nm, t[0, 0, k], t[0, 1:, k].mean()))
        print("   total %.0f" % t[0, 0].sum())
    sys.exit(0)

for label, d in (("cfg2", synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"])),
                 ("cfg2 full n=20", synth.make_train_inputs(K=5, B=32, M=20, dist="full", seed=5))):
    B, P = d["B"], d["P"]
    cl = 1
    for w in ([warps] if warps else [4, 8, 16]):
        out = {"mask": torch.zeros(max(B * P, B * 16 * NS * 2 * 4 + 64), dtype=torch.int32, device="cuda")}
        evs = []
        for _ in range(3):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss.match_loss_raw(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]),
                                dev(d["priors"]), d["alpha"], want_mask=True, warps=w, out=out)
            b_.record()
            evs.append((a, b_))
        torch.cuda.synchronize()
        t = out["mask"].cpu().numpy().view(np.int64)[:B * cl * w * NS].reshape(B, cl * w, NS)
        b = int(np.argmax(d["num_gt"]))      # grid == B here: CTA b solves image b
        print("%s warps=%d: image %d (n=%d) per-warp mean cycles by phase" % (label, w, b, d["num_gt"][b]))
        tot = t[b, :, :9].mean(0)
        for k, nm in enumerate(names):
            print("   %-24s %9.0f  (%.0f per augmentation)" % (nm, tot[k], tot[k] / max(1, d["num_gt"][b])))
        print("   total %.0f cycles; slowest warp %.0f; exact cost evaluations %d (n*P = %d)" %
              (tot.sum(), t[b, :, :9].sum(1).max(), t[b, :, 9].sum(), d["num_gt"][b] * P))
        # kernel-level timeline from %globaltimer (ns): when each CTA started / finished
        g0, g1 = t[:, :, 10], t[:, :, 11]
        k0 = g0.min()
        cyc = t[:, :, :9].sum(2).max(1)
        slow = int(np.argmax(g1.max(1)))
        print("   timeline (ns after the first CTA started): CTA starts %d..%d, CTA ends %d..%d (last: image %d, n=%d, "
              "%d cycles); event time of the launch %.1f us" %
              (g0.min() - k0, g0.max() - k0, g1.max(1).min() - k0, g1.max() - k0, slow, d["num_gt"][slow], cyc[slow],
               evs[-1][0].elapsed_time(evs[-1][1]) * 1e3))
        order = np.argsort(-cyc)[:5]
        print("   slowest images: " + ", ".join("img %d n=%d %d cyc" % (i, d["num_gt"][i], cyc[i]) for i in order))

# COCO-person-shaped images (K=7, P=904, M=100): the per-image fixed costs at small GT counts
d = synth.make_train_inputs(K=7, B=148, M=100, dist="coco_person", seed=1004)
B, P = d["B"], d["P"]
for w in ([warps] if warps else [8]):
    out = {"mask": torch.zeros(max(B * P, B * 16 * NS * 2 * 4 + 64), dtype=torch.int32, device="cuda")}
    for _ in range(2):
        loss.match_loss_raw(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]),
                            dev(d["priors"]), d["alpha"], want_mask=True, warps=w, out=out)
    torch.cuda.synchronize()
    t = out["mask"].cpu().numpy().view(np.int64)[:B * w * NS].reshape(B, w, NS)
    for want in (0, 3, 8, int(d["num_gt"].max())):
        sel = np.where(d["num_gt"] == want)[0]
        if len(sel) == 0:
            continue
        tot = t[sel][:, :, :9].mean(0).mean(0)
        print("cfg4-shape warps=%d: %d images with n=%d, mean cycles by phase: %s  total %.0f; exact evals/image %.0f" %
              (w, len(sel), want, " ".join("%s=%.0f" % (nm.replace(" ", "_"), tot[k]) for k, nm in enumerate(names)),
               tot.sum(), t[sel][:, :, 9].sum(1).mean()))

# configs[4]-shaped images (K=11, P=1420, M=200, n ~ U{0..200}): where the solver's time goes
d = synth.make_train_inputs(K=11, B=148, M=200, dist="uniform", seed=1005)
B, P = d["B"], d["P"]
for w in ([warps] if warps in (8, 16) else [16]):
    out = {"mask": torch.zeros(max(B * P, B * 16 * NS * 2 * 4 + 64), dtype=torch.int32, device="cuda")}
    for _ in range(2):
        loss.match_loss_raw(dev(d["locations"]), dev(d["confidences"]).view(B, P), dev(d["gt"]), dev(d["num_gt"]),
                            dev(d["priors"]), d["alpha"], want_mask=True, warps=w, out=out)
    torch.cuda.synchronize()
    t = out["mask"].cpu().numpy().view(np.int64)[:B * w * NS].reshape(B, w, NS)
    for lo, hi in ((0, 20), (40, 60), (90, 110), (140, 160), (180, 200)):
        sel = np.where((d["num_gt"] >= lo) & (d["num_gt"] <= hi))[0]
        if len(sel) == 0:
            continue
        tot = t[sel][:, :, :9].mean(0).mean(0)
        print("cfg5-shape warps=%d: %d images with n in [%d,%d], mean cycles by phase: %s  total %.0f; exact evals/image %.0f "
              "(mean n*P %.0f)" %
              (w, len(sel), lo, hi, " ".join("%s=%.0f" % (nm.replace(" ", "_"), tot[k]) for k, nm in enumerate(names)),
               tot.sum(), t[sel][:, :, 9].sum(1).mean(), (d["num_gt"][sel] * P).mean()))
