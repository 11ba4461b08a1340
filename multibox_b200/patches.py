"""Multi-patch detection: the callers either side of the detect kernels.

Upstream of the kernels the reference cuts every test image into patches -- the whole image,
optionally its mirror image, and sliding crops -- and gives each patch the metadata the detect
loop body consumes (reference detect.py:20-72 ``extract_patches`` and detect.py:189-271):
``patch_offsets`` (y, x), ``patch_dims`` (h, w), ``patch_is_flipped``, ``patch_bbox_restrictions``
and ``patch_max_to_keep``.  ``extract_patches`` / ``patch_plan`` below reproduce exactly that
metadata on the host (pixel data -- JPEG decode, bilinear resize -- is out of scope: the network
trunk is not part of this path).

Downstream the reference just appends every patch's detections to one JSON list and leaves the
grouping by image and the removal of cross-patch duplicates to whoever reads
``results-dense-*.json`` (detect.py:438-460).  ``merge_patches`` does that step on the GPU with
the same detect kernel (zero priors: decode is the identity): per image, the detections of all of
its patches are pooled, sorted by score and put through greedy NMS.  Like the in-patch NMS this
is an extension with no reference counterpart (parity unpinned; its specification is restated for
the tests next to the in-patch NMS one).
"""
import numpy as np
import torch

from . import detect


def extract_patches(image_height, image_width, patch_dims, strides, non_edge_restriction=0.1):
    """Geometry of reference detect.py:20-72 (same loop order: rows of patches top to bottom, left
    to right).  Returns ``[patch_offsets int32 [n,2] (y,x), patch_restrictions float32 [n,4], n]``;
    a side that touches the image border is unrestricted (0 / 1), any other side is pulled in by
    `non_edge_restriction` (detect.py:50-54)."""
    patch_height, patch_width = int(patch_dims[0]), int(patch_dims[1])
    h_stride, w_stride = int(strides[0]), int(strides[1])
    ys = np.arange(0, int(image_height) - patch_height + 1, h_stride, dtype=np.int64)
    xs = np.arange(0, int(image_width) - patch_width + 1, w_stride, dtype=np.int64)
    if len(ys) == 0 or len(xs) == 0:
        return [np.zeros((0, 2), np.int32), np.zeros((0, 4), np.float32), np.int32(0)]
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    yy, xx = yy.ravel(), xx.ravel()
    lo, hi = np.float32(non_edge_restriction), np.float32(1. - non_edge_restriction)
    r = np.empty((len(yy), 4), dtype=np.float32)
    r[:, 0] = np.where(xx == 0, np.float32(0.), lo)
    r[:, 1] = np.where(yy == 0, np.float32(0.), lo)
    r[:, 2] = np.where(xx + patch_width == image_width, np.float32(1.), hi)
    r[:, 3] = np.where(yy + patch_height == image_height, np.float32(1.), hi)
    return [np.stack([yy, xx], 1).astype(np.int32), r, np.int32(len(yy))]


def patch_plan(image_height, image_width, detection_cfg):
    """Per-patch metadata of ONE image in the order the reference enqueues it (detect.py:203-271):
    original image, flipped original, then every CROPS entry.  `detection_cfg` is the DETECTION
    section of the reference's config (config.yaml.example:62-84) as a dict.  Returns a dict of
    arrays: offsets [n,2] i32, patch_dims [n,2] i32, is_flipped [n,1] i32, restrictions [n,4] f32,
    max_to_keep [n,1] i32, image_dims [n,2] i32."""
    off, dims, flip, restr, keep = [], [], [], [], []
    whole = np.array([[0., 0., 1., 1.]], np.float32)
    if detection_cfg.get("USE_ORIGINAL_IMAGE"):                                   # detect.py:204-222
        off.append(np.zeros((1, 2), np.int32))
        dims.append(np.array([[image_height, image_width]], np.int32))
        flip.append(np.zeros((1, 1), np.int32))
        restr.append(whole)
        keep.append(np.array([[detection_cfg["ORIGINAL_IMAGE_MAX_TO_KEEP"]]], np.int32))
    if detection_cfg.get("USE_FLIPPED_ORIGINAL_IMAGE"):                           # detect.py:224-241
        off.append(np.zeros((1, 2), np.int32))
        dims.append(np.array([[image_height, image_width]], np.int32))
        flip.append(np.ones((1, 1), np.int32))
        restr.append(whole)
        keep.append(np.array([[detection_cfg["FLIPPED_IMAGE_MAX_TO_KEEP"]]], np.int32))
    for crop in detection_cfg.get("CROPS", []) or []:                             # detect.py:244-271
        o, r, n = extract_patches(image_height, image_width, (crop["HEIGHT"], crop["WIDTH"]),
                                  (crop["HEIGHT_STRIDE"], crop["WIDTH_STRIDE"]))
        n = int(n)
        off.append(o)
        dims.append(np.tile(np.array([[crop["HEIGHT"], crop["WIDTH"]]], np.int32), (n, 1)))
        flip.append(np.full((n, 1), 1 if crop.get("FLIP") else 0, np.int32))
        restr.append(r)
        keep.append(np.full((n, 1), crop["MAX_TO_KEEP"], np.int32))

    def cat(parts, width, dtype):
        return np.concatenate(parts, 0) if parts else np.zeros((0, width), dtype)

    out = dict(offsets=cat(off, 2, np.int32), patch_dims=cat(dims, 2, np.int32), is_flipped=cat(flip, 1, np.int32),
               restrictions=cat(restr, 4, np.float32), max_to_keep=cat(keep, 1, np.int32))
    n = out["offsets"].shape[0]
    out["image_dims"] = np.tile(np.array([[image_height, image_width]], np.int32), (n, 1))   # detect.py:274
    return out


def batch_plan(image_dims, detection_cfg):
    """patch_plan for a list of (height, width) images, concatenated in image order, plus
    ``image_index`` [n] (which image a patch belongs to) -- the batch the detect kernel sees."""
    plans = [patch_plan(h, w, detection_cfg) for h, w in image_dims]
    out = {k: np.concatenate([p[k] for p in plans], 0) for k in plans[0]} if plans else {}
    out["image_index"] = np.concatenate([np.full(p["offsets"].shape[0], i, np.int32) for i, p in enumerate(plans)]) \
        if plans else np.zeros((0,), np.int32)
    return out


def merge_patches(post, image_index, num_images, nms_iou=0.5, max_detections=200):
    """Pools the detections of all patches of each image and removes cross-patch duplicates.

    post: the dict detect.postprocess returned for a batch of patches (boxes f64 [Bp,k,4] in IMAGE
    coordinates, scores [Bp,k], count [Bp]); image_index int [Bp]: the image each patch belongs to
    (0 <= . < num_images, patches of an image in any order).  Per image: candidates in (patch
    order, rank) order, sorted by descending score (ties: later candidate first, the kernel's
    documented tie rule), greedy NMS at `nms_iou` on the float32 boxes (None = no NMS) over the pooled
    candidates -- the top 1024 by score when an image has more, the kernel's per-image capacity -- and
    the first `max_detections` of the KEPT list are returned.  Returns boxes f64 [I,kmax,4] (the
    patches' float64 values, gathered), scores f32 [I,kmax], source_patch i32 [I,kmax] (-1 padding),
    count i32 [I].  One host read-back (the largest per-image candidate count sizes the pooled buffer)."""
    boxes, scores, count = post["boxes"], post["scores"], post["count"]
    dev = boxes.device
    Bp, k = scores.shape
    idx = torch.as_tensor(image_index, device=dev).to(torch.int64).view(Bp)
    # candidate slot of every detection: detections of an image packed back to back in (patch
    # order, rank) order -- index plumbing only (stable sort of patches by image + exclusive scan)
    order = torch.argsort(idx, stable=True)
    cnt_sorted = count.to(torch.int64)[order]
    img_sorted = idx[order]
    cum = torch.cumsum(cnt_sorted, 0) - cnt_sorted                     # exclusive scan over sorted patches
    per_image = torch.zeros(num_images, dtype=torch.int64, device=dev).index_add_(0, img_sorted, cnt_sorted)
    img_start = torch.cumsum(per_image, 0) - per_image
    base = torch.empty(Bp, dtype=torch.int64, device=dev)
    base[order] = cum - img_start[img_sorted]                           # first slot of each patch inside its image
    Pm = max(1, int(per_image.max().item()) if Bp else 1)
    Pm = (Pm + 3) // 4 * 4
    slot = base.view(Bp, 1) + torch.arange(k, device=dev).view(1, k)
    valid = torch.arange(k, device=dev).view(1, k) < count.view(Bp, 1)
    flat = (idx.view(Bp, 1) * Pm + slot)[valid]
    cand_boxes64 = torch.zeros((num_images * Pm, 4), dtype=torch.float64, device=dev)
    cand_scores = torch.full((num_images * Pm,), float("-inf"), dtype=torch.float32, device=dev)
    cand_patch = torch.full((num_images * Pm,), -1, dtype=torch.int32, device=dev)
    cand_boxes64[flat] = boxes[valid]
    cand_scores[flat] = scores[valid]
    cand_patch[flat] = torch.arange(Bp, device=dev, dtype=torch.int32).view(Bp, 1).expand(Bp, k)[valid]
    kpool = max(1, min(Pm, detect.K_MAX_LIMIT))          # candidates that enter the suppression
    kmax = max(1, min(int(max_detections), kpool))
    # empty slots carry a -inf score: mbx_detect never takes them as proposals; priors=None: identity decode
    merged = detect.postprocess(cand_boxes64.view(num_images, Pm, 4).to(torch.float32),
                                cand_scores.view(num_images, Pm, 1), None, nms_iou=nms_iou, k_max=kpool,
                                want_patch_boxes=False)
    merged = {k: (v[:, :kmax].contiguous() if v.dim() > 1 else torch.clamp(v, max=kmax)) for k, v in merged.items()}
    pi = merged["prior_idx"].to(torch.int64).clamp_(min=0)
    kept = merged["prior_idx"] >= 0
    out_boxes = torch.gather(cand_boxes64.view(num_images, Pm, 4), 1, pi.unsqueeze(-1).expand(-1, -1, 4))
    out_boxes = torch.where(kept.unsqueeze(-1), out_boxes, torch.zeros_like(out_boxes))
    src = torch.gather(cand_patch.view(num_images, Pm), 1, pi)
    src = torch.where(kept, src, torch.full_like(src, -1))
    return dict(boxes=out_boxes, scores=merged["scores"], source_patch=src, candidate_idx=merged["prior_idx"],
                count=merged["count"])
