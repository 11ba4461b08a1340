"""TEST INFRASTRUCTURE -- CPU (numpy/scipy) restatement of the Multibox hot path.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  ``multibox_b200`` never does.

It restates, in plain numpy with the reference's fp32/fp64 operation order and
the reference's loop structure (so that timing it is timing the reference's
algorithm), the following reference code (paths relative to the reference
checkout):

  generate_priors        priors.py:185-314
  compute_assignments    loss.py:8-53        (scipy.optimize.linear_sum_assignment,
                                              loss.py:2,40 -- third-party, pinned
                                              scipy==0.17.0 in requirements.txt:5;
                                              this image has scipy 1.18.1)
  add_loss (+ autodiff)  loss.py:55-117      (TensorFlow graph; restated in numpy:
                                              elementwise fp32 as written, sums
                                              accumulated in fp64)
  filter_proposals       detect.py:74-104
  convert_proposals      detect.py:106-131
  detect loop body       detect.py:408-443
  eval loop body         eval.py:142-175
  sigmoid                model.py:322

Parity pinning: the reference's own tests (model_tests.py) hold no numeric
golden vectors for this path, only the constant 646 and sign/identity
properties.  The restatement is therefore pinned against the reference's own
functions executed from line slices in the build container
(``oracle/ref_slices.py`` + ``oracle/gen_golden.py`` -> ``tests/golden/``).
NMS has no reference implementation at all: ``greedy_nms`` below is this
project's specification -- PARITY UNPINNED for NMS (cross-checked against
torchvision.ops.nms on CPU only).
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

SMALL_EPSILON = 1e-10          # loss.py:6
GRIDS = (8, 6, 4, 3, 2, 1)     # priors.py:196


# ----------------------------------------------------------------------------
# priors.py:185-314
# ----------------------------------------------------------------------------
def _one_prior(center_i, center_j, scale, a, restrict):
    """One prior box in python/np.float64 scalar arithmetic, priors.py:269-310
    (the 1x1 branch priors.py:206-256 is the same arithmetic with a = 1.)."""
    w = scale * np.sqrt(a)
    h = scale / np.sqrt(a)
    x1 = center_j - (w / 2.)
    x2 = center_j + (w / 2.)
    y1 = center_i - (h / 2.)
    y2 = center_i + (h / 2.)
    if restrict:
        right_trim = abs(min(0, x1))
        left_trim = abs(min(0, 1 - x2))
        top_trim = abs(min(0, y1))
        bottom_trim = abs(min(0, 1 - y2))
        trim = max(max(right_trim, left_trim), max(top_trim, bottom_trim))
        if h > w:
            width_trim, height_trim = trim * a, trim
        else:
            width_trim, height_trim = trim, trim / a
        xa, xb = x1 + width_trim, x2 - width_trim
        ya, yb = y1 + height_trim, y2 - height_trim
        x1, x2 = min(xa, xb), max(xa, xb)
        y1, y2 = min(ya, yb), max(ya, yb)
    return [max(x1, 0.), max(y1, 0.), min(x2, 1.), min(y2, 1.)]


def generate_priors(aspect_ratios, min_scale=0.1, max_scale=0.95,
                    restrict_to_image_bounds=True):
    n = len(GRIDS)
    scales = [min_scale + (max_scale - min_scale) * (i - 1) / (n - 1)
              for i in range(1, n + 1)]                      # priors.py:199-200
    out = []
    for grid, scale in zip(GRIDS, scales):
        if grid == 1:
            out.append(_one_prior(0.5, 0.5, scale, 1., restrict_to_image_bounds))
            continue
        for i in range(grid):
            for j in range(grid):
                ci = (i + 0.5) / grid
                cj = (j + 0.5) / grid
                for a in aspect_ratios:
                    out.append(_one_prior(ci, cj, scale, a, restrict_to_image_bounds))
    return out


# ----------------------------------------------------------------------------
# loss.py:8-53
# ----------------------------------------------------------------------------
def log_terms(confidences):
    """loss.py:21-25 on the whole batch: fp32 numpy log of c and of clamp(1-c)."""
    log_c = np.log(confidences)
    v = 1. - confidences
    v[v > 1.] = 1.
    v[v <= 0] = SMALL_EPSILON
    return log_c, np.log(v)


def cost_matrix(loc_abs, log_c, log_1mc, gt, alpha):
    """loss.py:33-35 for one image: float64 [P, n] holding fp32-computed costs.
    Column loop kept as in the reference (it is ~85% of the CPU time)."""
    P = loc_abs.shape[0]
    n = gt.shape[0]
    C = np.zeros((P, n))
    for j in range(n):
        C[:, j] = (alpha / 2.) * (np.linalg.norm(loc_abs - gt[j], axis=1)) ** 2 \
            - log_c + log_1mc
    return C


def compute_assignments(locations, confidences, gt_bboxes, num_gt_bboxes,
                        batch_size, alpha, return_indices=False):
    """loss.py:8-53.  locations [B*P,4] f32 (prior already added), confidences
    [B*P] f32 (epsilon already added), gt_bboxes [B,M,4] f32, num_gt_bboxes [B]
    i32, alpha fp32 scalar.  Returns [mask int32 [B*P], stacked_gt f32 [N,4]];
    with return_indices also matched_gt_idx int32 [B*P] (-1 = unmatched)."""
    B = int(batch_size)
    P = locations.shape[0] // B                       # loss.py:16 (py2 int division)
    mask = np.zeros(B * P, dtype=np.int32)
    gt_idx = np.full(B * P, -1, dtype=np.int32)
    stacked = np.zeros([0, 4], dtype=np.float32)
    log_c, log_1mc = log_terms(confidences)
    for b in range(B):
        lo = b * P
        C = cost_matrix(locations[lo:lo + P], log_c[lo:lo + P], log_1mc[lo:lo + P],
                        gt_bboxes[b][:num_gt_bboxes[b]], alpha)
        rows, cols = linear_sum_assignment(C)         # loss.py:40
        for r, c in zip(rows, cols):
            mask[lo + r] = 1
            gt_idx[lo + r] = c
            stacked = np.concatenate((stacked, gt_bboxes[b][c].reshape([1, 4])))
    stacked = stacked.astype(np.float32)
    if return_indices:
        return [mask, stacked, gt_idx]
    return [mask, stacked]


# ----------------------------------------------------------------------------
# loss.py:55-117 (+ TF autodiff), restated
# ----------------------------------------------------------------------------
def sigmoid(z):
    """model.py:322 stand-in (fp32)."""
    z = np.asarray(z, dtype=np.float32)
    return (np.float32(1.) / (np.float32(1.) + np.exp(-z))).astype(np.float32)


def add_loss(locations, confidences, batched_bboxes, batched_num_bboxes,
             bbox_priors, location_loss_alpha, assignments=None):
    """loss.py:55-117 forward + the gradients TF autodiff would produce
    (SURVEY.md section 8a rows a2-a12).  Inputs: locations [B,P,4] f32,
    confidences [B,P,1] f32 (post-sigmoid), batched_bboxes [B,M,4] f32,
    batched_num_bboxes [B] i32, bbox_priors [P,4] f32.
    `assignments` may inject a precomputed (mask, stacked_gt) pair.
    Returns a dict."""
    locations = np.asarray(locations, dtype=np.float32)
    confidences = np.asarray(confidences, dtype=np.float32)
    B, P = locations.shape[0], locations.shape[1]
    alpha32 = np.float32(location_loss_alpha)
    loc = locations.reshape(-1, 4) + np.tile(np.asarray(bbox_priors, np.float32), (B, 1))  # :67,71
    conf = confidences.reshape(-1) + np.float32(SMALL_EPSILON)                              # :68,74
    if assignments is None:
        mask, stacked, gt_idx = compute_assignments(                                        # :82
            loc, conf, np.asarray(batched_bboxes, np.float32),
            np.asarray(batched_num_bboxes, np.int32), np.int32(B), alpha32,
            return_indices=True)
    else:
        mask, stacked = assignments[0], assignments[1]
        gt_idx = None
    m = mask.astype(bool)
    matched_loc, unmatched_conf = loc[m], conf[~m]                                          # :88-89
    matched_conf = conf[m]
    # sentinels (:94-97) contribute exactly 0 to both sums, so they are omitted.
    diff = matched_loc - stacked                                                            # fp32
    l2 = np.sum((diff * diff).astype(np.float64)) / 2.                                      # tf.nn.l2_loss
    location_loss = float(alpha32) * l2                                                     # :100
    eps32 = np.float32(SMALL_EPSILON)
    neg_arg = (np.float32(1.) - unmatched_conf) + eps32
    with np.errstate(divide="ignore"):
        confidence_loss = -np.sum(np.log(matched_conf).astype(np.float64)) \
            - np.sum(np.log(neg_arg).astype(np.float64))                                    # :101
    d_loc = np.zeros_like(loc)
    d_loc[m] = alpha32 * diff
    d_conf = np.zeros_like(conf)
    d_conf[m] = np.float32(-1.) / matched_conf
    d_conf[~m] = np.float32(1.) / neg_arg
    return {
        "location_loss": np.float32(location_loss),
        "confidence_loss": np.float32(confidence_loss),
        "location_loss_f64": location_loss,
        "confidence_loss_f64": float(confidence_loss),
        "mask": mask, "stacked_gt": stacked, "matched_gt_idx": gt_idx,
        "d_locations": d_loc.reshape(B, P, 4),
        "d_confidences": d_conf.reshape(B, P, 1),
    }


def add_loss_from_logits(locations, logits, *args, **kw):
    """Extension entry (fuses model.py:322): confidences = sigmoid(logits);
    d_logits = d_conf * s * (1 - s)."""
    s = sigmoid(logits)
    out = add_loss(locations, s, *args, **kw)
    out["confidences"] = s
    out["d_logits"] = out["d_confidences"] * s * (np.float32(1.) - s)
    return out


# ----------------------------------------------------------------------------
# detect.py:74-131, 408-443
# ----------------------------------------------------------------------------
def filter_proposals(bboxes, confidences, restrictions=None):
    """detect.py:74-104 (per-box python loop kept: it dominates the CPU time)."""
    if restrictions is None:
        restrictions = [0.1, 0.1, 0.9, 0.9]
    keep_b, keep_c = [], []
    for bbox, conf in zip(bboxes, confidences):
        if bbox[0] < restrictions[0] or bbox[1] < restrictions[1] \
                or bbox[2] > restrictions[2] or bbox[3] > restrictions[3]:
            continue
        keep_b.append(bbox)
        keep_c.append(conf)
    return np.array(keep_b), np.array(keep_c)


def convert_proposals(bboxes, offset, patch_dims, image_dims, is_flipped=0):
    """detect.py:106-131: patch -> image coordinates in float64; offset (y,x),
    dims (h,w)."""
    sx = patch_dims[1] / float(image_dims[1])
    sy = patch_dims[0] / float(image_dims[0])
    ox = offset[1] / float(image_dims[1])
    oy = offset[0] / float(image_dims[0])
    out = bboxes * np.array([sx, sy, sx, sy]) + np.array([ox, oy, ox, oy])
    if is_flipped:
        out[:, [0, 2]] = out[:, [2, 0]]
        out[:, 0] = 1. - out[:, 0]
        out[:, 2] = 1. - out[:, 2]
    return out


def greedy_nms(boxes, iou_threshold):
    """PROJECT SPECIFICATION (no reference counterpart; parity unpinned).
    boxes [k,4] f32 already in descending-score order.  Box i is kept unless an
    already-kept j<i has IoU(i,j) > iou_threshold (strict).  fp32 arithmetic:
    area=(x2-x1)*(y2-y1); inter=max(0,ix2-ix1)*max(0,iy2-iy1);
    iou=inter/((area_i+area_j)-inter).  Returns positions kept (ascending)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    k = boxes.shape[0]
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    dead = np.zeros(k, dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    zero = np.float32(0.)
    for i in range(k):
        if dead[i]:
            continue
        keep.append(i)
        if i + 1 == k:
            break
        r = boxes[i + 1:]
        w = np.maximum(zero, np.minimum(boxes[i, 2], r[:, 2]) - np.maximum(boxes[i, 0], r[:, 0]))
        h = np.maximum(zero, np.minimum(boxes[i, 3], r[:, 3]) - np.maximum(boxes[i, 1], r[:, 1]))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / ((area[i] + area[i + 1:]) - inter)
        dead[i + 1:] |= iou > thr          # NaN (0/0) compares False: not suppressed
    return np.array(keep, dtype=np.int64)


def postprocess(locs, confs, bbox_priors, restrictions, max_to_keep, offsets,
                patch_dims, image_dims, is_flipped, nms_iou=None):
    """detect.py:408-436 per image, with the original prior index carried along
    (the reference never exposes it) and an optional greedy NMS on the kept
    top-k (extension).  Tie rule pinned: stable argsort then reversal, i.e.
    equal confidences come out in descending index order.
    Returns a list (one entry per image) of dicts: boxes f64 [c,4] (converted),
    patch_boxes f32 [c,4], scores f32 [c], prior_idx i64 [c]."""
    out = []
    priors = np.asarray(bbox_priors, np.float32)
    for b in range(locs.shape[0]):
        boxes = np.clip(locs[b] + priors, 0., 1.)                     # :412-413
        conf = confs[b]                                               # [P,1]
        # filter_proposals with the surviving original indices carried along
        fb, fc = filter_proposals(boxes, conf, restrictions[b])       # :416
        r = restrictions[b]
        ok = ~((boxes[:, 0] < r[0]) | (boxes[:, 1] < r[1]) | (boxes[:, 2] > r[2]) | (boxes[:, 3] > r[3]))
        orig = np.nonzero(ok)[0]
        assert orig.shape[0] == fb.shape[0]
        if fb.shape[0] == 0:                                          # :419-420
            out.append({"boxes": np.zeros((0, 4)), "patch_boxes": np.zeros((0, 4), np.float32),
                        "scores": np.zeros((0,), np.float32), "prior_idx": np.zeros((0,), np.int64)})
            continue
        k = int(np.asarray(max_to_keep[b]).reshape(-1)[0])            # :423
        order = np.argsort(fc.ravel(), kind="stable")[::-1][:k]       # :424-425
        fb, fc, orig = fb[order], fc[order], orig[order]              # :426-427
        if nms_iou is not None:
            kept = greedy_nms(fb, nms_iou)
            fb, fc, orig = fb[kept], fc[kept], orig[kept]
        conv = convert_proposals(fb, offsets[b], patch_dims[b], image_dims[b],
                                 is_flipped=int(np.asarray(is_flipped[b]).reshape(-1)[0]))  # :430
        out.append({"boxes": conv, "patch_boxes": fb.astype(np.float32),
                    "scores": fc.ravel().astype(np.float32), "prior_idx": orig})
    return out


def eval_topk(locs, confs, bbox_priors, input_size, image_ids, k=100):
    """eval.py:142-175: decode+clip, scale to pixels, full descending sort
    (stable+reverse tie rule), top-k rows [img_id, x, y, w, h, score, 1]."""
    priors = np.asarray(bbox_priors, np.float32)
    rows = []
    for b in range(locs.shape[0]):
        boxes = np.clip(locs[b] + priors, 0., 1.)                     # :146-147
        scale = np.array([input_size] * 4)                            # :156 (int64 array)
        boxes = boxes * scale                                         # :157 -> float64
        order = np.argsort(confs[b].ravel(), kind="stable")[::-1]     # :162
        sb, sc = boxes[order], confs[b][order]
        for t in range(k):                                            # :167
            x1, y1, x2, y2 = sb[t]
            rows.append([int(image_ids[b]), x1, y1, x2 - x1, y2 - y1, float(sc[t].reshape(-1)[0]), 1])
    return rows


# ----------------------------------------------------------------------------- layout steps either side of the path
def concat_heads(head_locations, head_confidences):
    """reference model.py:295-320 (without the sigmoid of :322): per-head NHWC outputs
    [B,g,g,K*4] / [B,g,g,K] are flattened per image and concatenated in head order, then viewed
    as [B,P,4] / [B,P,1].  This fixes the prior order the priors of priors.py:185-314 follow."""
    B = head_locations[0].shape[0]
    loc = np.concatenate([np.reshape(t, (B, -1)) for t in head_locations], axis=1)
    conf = np.concatenate([np.reshape(t, (B, -1)) for t in head_confidences], axis=1)
    return np.reshape(loc, (B, -1, 4)), np.reshape(conf, (B, -1, 1))


def pad_ragged_gt(gt_flat, gt_row_offsets, max_num_bboxes):
    """reference inputs.py:340-348: every image's boxes zero-padded to MAX_NUM_BBOXES rows."""
    B = len(gt_row_offsets) - 1
    gt = np.zeros((B, max_num_bboxes, 4), dtype=np.float32)
    num = np.zeros((B,), dtype=np.int32)
    for b in range(B):
        lo, hi = int(gt_row_offsets[b]), int(gt_row_offsets[b + 1])
        num[b] = hi - lo
        gt[b, :hi - lo] = gt_flat[lo:hi]
    return gt, num


def merge_patches(per_patch, image_index, num_images, nms_iou=0.5, max_detections=200):
    """PROJECT SPECIFICATION (no reference counterpart; parity unpinned): what
    multibox_b200.patches.merge_patches computes.  per_patch = the list postprocess() returns
    (one dict per patch: boxes f64 [c,4] in image coordinates, scores f32 [c]).  Per image: pool the
    detections of its patches in (patch order, rank) order, stable-argsort-then-reverse by score
    (ties: later candidate first), greedy NMS on the float32 boxes over the pooled candidates (the top
    1024 by score when there are more: the kernel's per-image capacity), then the first max_detections
    of the kept list."""
    out = []
    for i in range(num_images):
        pats = [b for b in range(len(per_patch)) if image_index[b] == i]
        boxes = np.concatenate([per_patch[b]["boxes"].reshape(-1, 4) for b in pats] + [np.zeros((0, 4))], 0)
        scores = np.concatenate([per_patch[b]["scores"].reshape(-1) for b in pats] + [np.zeros((0,), np.float32)], 0)
        src = np.concatenate([np.full(per_patch[b]["scores"].reshape(-1).shape[0], b, np.int32) for b in pats] +
                             [np.zeros((0,), np.int32)], 0)
        order = np.argsort(scores.astype(np.float32), kind="stable")[::-1][:1024]
        boxes, scores, src = boxes[order], scores[order], src[order]
        if nms_iou is not None and len(order):
            keep = greedy_nms(boxes.astype(np.float32), nms_iou)
            boxes, scores, src = boxes[keep], scores[keep], src[keep]
        boxes, scores, src = boxes[:max_detections], scores[:max_detections], src[:max_detections]
        out.append(dict(boxes=boxes, scores=scores.astype(np.float32), source_patch=src))
    return out
