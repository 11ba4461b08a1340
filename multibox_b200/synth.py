"""Seeded synthetic inputs of the shapes named in BASELINE.json:configs.

There is no dataset or checkpoint access, so the head outputs are random-init
shaped tensors: offsets ~ N(0, 0.1^2) fp32 [B,P,4], logits ~
N(0,1) fp32 [B,P,1] (confidences = sigmoid), ground-truth boxes with two
corners ~ U[0,1]^2 sorted per axis, side >= 0.02, distinct per image, rows past
the image's count zero-filled (reference inputs.py:346-347).  Everything is
generated on the host with numpy so the CPU oracle and the CUDA path see the
same bits.
"""
import numpy as np

from .priors import priors_fp32, num_priors

ASPECT_RATIOS = {
    5: [1., 2., 3., 1. / 2., 1. / 3.],                      # BASELINE.json configs[0]
    7: np.geomspace(1. / 3., 3., 7).tolist(),               # COCO-person-like (K=7)
    11: np.geomspace(0.25, 4., 11).tolist(),                # "11 clustered aspect ratios"
}

# name -> (K, B, M, n_b distribution, seed, alpha)
TRAIN_CONFIGS = {
    "cfg1": dict(K=5, B=32, M=20, dist="uniform", seed=1001, alpha=1000.0),
    "cfg2": dict(K=5, B=32, M=20, dist="uniform", seed=1002, alpha=1000.0),
    "cfg4": dict(K=7, B=1024, M=100, dist="coco_person", seed=1004, alpha=1000.0),
    "cfg5": dict(K=11, B=8192, M=200, dist="uniform", seed=1005, alpha=1000.0),
}
DETECT_CONFIGS = {
    "cfg3": dict(K=5, B=256, keep=200, nms_iou=0.5, seed=1003),
    "cfg5d": dict(K=11, B=8192, keep=200, nms_iou=0.5, seed=1005),
}


def _sigmoid32(z):
    return (np.float32(1.) / (np.float32(1.) + np.exp(-z.astype(np.float32)))).astype(np.float32)


def gt_counts(rng, B, M, dist):
    if dist == "uniform":
        return rng.integers(0, M + 1, size=B).astype(np.int32)
    if dist == "full":
        return np.full(B, M, dtype=np.int32)
    if dist == "coco_person":      # min(M, floor(Exp(mean 4))): ~22 % of images have no GT
        return np.minimum(M, np.floor(rng.exponential(4.0, size=B))).astype(np.int32)
    raise ValueError(dist)


def gt_boxes(rng, counts, M):
    """[B,M,4] f32: x1<x2, y1<y2, sides >= 0.02, rows >= n_b are zero."""
    B = counts.shape[0]
    out = np.zeros((B, M, 4), dtype=np.float32)
    for b in range(B):
        n = int(counts[b])
        if n == 0:
            continue
        boxes = np.zeros((0, 4), dtype=np.float32)
        while boxes.shape[0] < n:
            c = rng.uniform(size=(2 * n + 4, 2, 2))
            lo, hi = c.min(axis=1), c.max(axis=1)
            cand = np.concatenate([lo, hi], axis=1).astype(np.float32)
            ok = ((cand[:, 2] - cand[:, 0]) >= 0.02) & ((cand[:, 3] - cand[:, 1]) >= 0.02)
            boxes = np.unique(np.concatenate([boxes, cand[ok]]), axis=0)
            rng.shuffle(boxes)
        out[b, :n] = boxes[:n]
    return out


def make_train_inputs(K=5, B=32, M=20, dist="uniform", seed=0, alpha=1000.0,
                      edge_cases=False):
    """dict: priors [P,4] f32, locations [B,P,4] f32, logits/confidences [B,P,1]
    f32, gt [B,M,4] f32, num_gt [B] i32, alpha."""
    rng = np.random.default_rng(seed)
    P = num_priors(K)
    priors = priors_fp32(ASPECT_RATIOS[K])
    locations = rng.normal(0., 0.1, size=(B, P, 4)).astype(np.float32)
    logits = rng.normal(0., 1., size=(B, P, 1)).astype(np.float32)
    counts = gt_counts(rng, B, M, dist)
    if edge_cases and B >= 3:
        counts[0], counts[1] = 0, M            # empty image and a full image
        logits[2, ::7] = 30.0                  # saturated confidences
        logits[2, 3::7] = -30.0
    gt = gt_boxes(rng, counts, M)
    return dict(K=K, P=P, B=B, M=M, priors=priors, locations=locations, logits=logits,
                confidences=_sigmoid32(logits), gt=gt, num_gt=counts, alpha=float(alpha))


def make_detect_inputs(K=5, B=256, keep=200, seed=0, nms_iou=0.5, patches=False):
    """dict for the detect path.  With patches=False every item is a whole image
    (restriction [0,0,1,1], offset 0, patch == image, reference detect.py:204,222);
    with patches=True a mix of whole images, crops (restriction 0.1/0.9 on the
    non-edge sides, reference detect.py:50-54) and x-flipped patches."""
    rng = np.random.default_rng(seed)
    P = num_priors(K)
    priors = priors_fp32(ASPECT_RATIOS[K])
    locations = rng.normal(0., 0.1, size=(B, P, 4)).astype(np.float32)
    logits = rng.normal(0., 1., size=(B, P, 1)).astype(np.float32)
    restrictions = np.tile(np.array([0., 0., 1., 1.], np.float32), (B, 1))
    max_to_keep = np.full((B, 1), keep, dtype=np.int32)
    image_dims = np.stack([rng.integers(300, 1200, size=B), rng.integers(300, 1200, size=B)], 1).astype(np.int32)
    patch_dims = image_dims.copy()
    offsets = np.zeros((B, 2), dtype=np.int32)
    is_flipped = np.zeros((B, 1), dtype=np.int32)
    if patches:
        for b in range(B):
            kind = b % 4
            if kind == 1 or kind == 3:   # interior crop: keep only boxes inside [.1,.9]^2 on non-edge sides
                h, w = image_dims[b]
                ph, pw = max(32, h // 2), max(32, w // 2)
                oy, ox = rng.integers(0, h - ph + 1), rng.integers(0, w - pw + 1)
                r = np.array([0.1, 0.1, 0.9, 0.9], np.float32)
                if ox == 0: r[0] = 0.
                if oy == 0: r[1] = 0.
                if ox + pw == w: r[2] = 1.
                if oy + ph == h: r[3] = 1.
                restrictions[b] = r
                patch_dims[b] = (ph, pw)
                offsets[b] = (oy, ox)
                max_to_keep[b] = max(1, keep // 2)
            if kind >= 2:
                is_flipped[b] = 1
    return dict(K=K, P=P, B=B, priors=priors, locations=locations, logits=logits,
                confidences=_sigmoid32(logits), restrictions=restrictions,
                max_to_keep=max_to_keep, offsets=offsets, patch_dims=patch_dims,
                image_dims=image_dims, is_flipped=is_flipped,
                image_ids=np.arange(B, dtype=np.int64), keep=keep, nms_iou=nms_iou)


def split_heads(locations, confidences, K, grids=(8, 6, 4, 3, 2, 1)):
    """Inverse of the reference's layout step (model.py:295-320): cut concatenated [B,P,4] /
    [B,P,1] arrays back into the per-head NHWC conv outputs [B,g,g,K*4] / [B,g,g,K] they would
    have been concatenated from (the 1x1 head has a single box, model.py:281-287)."""
    B = locations.shape[0]
    hl, hc, off = [], [], 0
    for g in grids:
        k = K if g > 1 else 1
        n = g * g * k
        hl.append(np.ascontiguousarray(locations[:, off:off + n].reshape(B, g, g, k * 4)))
        hc.append(np.ascontiguousarray(confidences[:, off:off + n].reshape(B, g, g, k)))
        off += n
    assert off == locations.shape[1]
    return hl, hc


def ragged_gt(gt, num_gt):
    """Padded [B,M,4] + counts -> (gt_flat [N,4], gt_row_offsets [B+1] int32)."""
    off = np.zeros(len(num_gt) + 1, dtype=np.int32)
    off[1:] = np.cumsum(num_gt)
    flat = np.concatenate([gt[b, :num_gt[b]] for b in range(len(num_gt))] + [np.zeros((0, 4), np.float32)], axis=0)
    return np.ascontiguousarray(flat.astype(np.float32)), off
