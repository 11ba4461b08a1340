"""TEST INFRASTRUCTURE -- container-only loader for the reference's own functions.

Executes *line slices* of the read-only reference checkout (never copies them
into this repo) so that the numpy/C restatements in ``oracle/`` can be pinned
against the reference's real code, and so that ``oracle/gen_golden.py`` can
write golden vectors under ``tests/golden/``.

``/root/reference`` does not exist on the GPU box: nothing under ``tests/``
(-m gpu), ``bench.py`` or ``__graft_entry__.smoke()`` may import this module at
run time.  It is used only by ``oracle/gen_golden.py`` and by the CPU tests
that are skipped when the checkout is absent.

Slices:
  * priors.py:185-314   generate_priors            -- verbatim
  * loss.py:6-53        compute_assignments        -- one shim: the Python-2
                        integer division at loss.py:16 becomes ``//``
  * detect.py:74-131    filter_proposals / convert_proposals -- verbatim
  * detect.py:20-72     extract_patches (patch geometry; called with an all-zero image) -- verbatim
  * eval.py:142-175     the per-image loop of eval() (decode, clip, scale, sort, top-100 rows),
                        executed VERBATIM (dedented) by ``eval_loop_body`` with one shim: the unstable
                        ``np.argsort(x)`` at eval.py:162 gets ``kind='stable'`` (tie order pinned)
The inline loop body detect.py:408-436 is not a function; its statement order
is followed by ``detect_loop_body`` below, calling the verbatim functions, with
``np.asscalar`` -> ``.item()`` and ``np.argsort(kind='stable')`` pinned (numpy's
default sort is unstable, so the reference leaves tie order unspecified).
"""
import os

import numpy as np
from scipy.optimize import linear_sum_assignment

REFERENCE_ROOT = os.environ.get("MBX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "loss.py"))


def _slice(fname, first, last):
    with open(os.path.join(REFERENCE_ROOT, fname)) as f:
        lines = f.readlines()
    return "".join(lines[first - 1:last])


_cache = {}


def load():
    """Returns a dict with the reference's own function objects."""
    if _cache:
        return _cache
    ns = {"np": np, "linear_sum_assignment": linear_sum_assignment}
    exec(compile(_slice("priors.py", 185, 314), "ref:priors.py:185-314", "exec"), ns)
    src = _slice("loss.py", 6, 53)
    shim_from = "num_predictions = locations.shape[0] / batch_size"
    assert src.count(shim_from) == 1, "reference loss.py:16 changed"
    src = src.replace(shim_from, "num_predictions = locations.shape[0] // batch_size")
    exec(compile(src, "ref:loss.py:6-53(+//)", "exec"), ns)
    exec(compile(_slice("detect.py", 74, 131), "ref:detect.py:74-131", "exec"), ns)
    exec(compile(_slice("detect.py", 20, 72), "ref:detect.py:20-72", "exec"), ns)
    for k in ("generate_priors", "compute_assignments", "filter_proposals",
              "convert_proposals", "extract_patches", "SMALL_EPSILON"):
        _cache[k] = ns[k]
    return _cache


def detect_loop_body(locs, confs, bbox_priors, patch_offsets, patch_dims,
                     patch_is_flipped, patch_bbox_restrictions, patch_max_to_keep,
                     image_height_widths, image_ids):
    """Statement-for-statement walk of detect.py:408-443 around the verbatim
    filter_proposals / convert_proposals.  Returns the detection_results list."""
    ref = load()
    detection_results = []
    for b in range(locs.shape[0]):
        img_id = int(image_ids[b].item())                                   # :410
        predicted_bboxes = locs[b] + bbox_priors                            # :412
        predicted_bboxes = np.clip(predicted_bboxes, 0., 1.)                # :413
        predicted_confs = confs[b]                                          # :414
        filtered_bboxes, filtered_confs = ref["filter_proposals"](          # :416
            predicted_bboxes, predicted_confs, patch_bbox_restrictions[b])
        if filtered_bboxes.shape[0] == 0:                                   # :419
            continue
        num_preds_to_keep = patch_max_to_keep[b].item()                     # :423
        sorted_idxs = np.argsort(filtered_confs.ravel(), kind="stable")[::-1]  # :424
        sorted_idxs = sorted_idxs[:num_preds_to_keep]                       # :425
        filtered_bboxes = filtered_bboxes[sorted_idxs]                      # :426
        filtered_confs = filtered_confs[sorted_idxs]                        # :427
        converted_bboxes = ref["convert_proposals"](                        # :430
            bboxes=filtered_bboxes, offset=patch_offsets[b],
            patch_dims=patch_dims[b], image_dims=image_height_widths[b],
            is_flipped=patch_is_flipped[b])
        for k in range(converted_bboxes.shape[0]):                          # :438
            detection_results.append({
                "image_id": img_id,
                "bbox": converted_bboxes[k].tolist(),
                "score": float(filtered_confs[k].item()),
            })
    return detection_results


def eval_loop_body(locs, confs, bbox_priors, input_size, image_ids):
    """Executes reference eval.py:142-175 verbatim (the ``for b in range(cfg.BATCH_SIZE)`` loop up to the
    ``pred_annotations.append`` of the top-100 rows) in a namespace that provides the names the loop
    reads.  One shim: ``np.argsort(predicted_confs.ravel())`` -> ``kind='stable'``.  Returns
    pred_annotations: rows [img_id, x1, y1, w, h, score, 1]."""
    import textwrap
    src = _slice("eval.py", 142, 175)
    shim_from = "np.argsort(predicted_confs.ravel())[::-1]"
    assert src.count(shim_from) == 1, "reference eval.py:162 changed"
    src = textwrap.dedent(src.replace(shim_from, "np.argsort(predicted_confs.ravel(), kind='stable')[::-1]"))

    class _Cfg:
        BATCH_SIZE = int(locs.shape[0])
        INPUT_SIZE = int(input_size)

    B = int(locs.shape[0])
    ns = {"np": np, "cfg": _Cfg, "locs": locs, "confs": confs, "bbox_priors": bbox_priors, "image_ids": image_ids,
          "all_gt_bboxes": np.zeros((B, 1, 4), np.float32), "all_gt_num_bboxes": np.zeros((B,), np.int32),
          "all_gt_areas": np.zeros((B, 1), np.float32), "pred_annotations": []}
    exec(compile(src, "ref:eval.py:142-175(+stable)", "exec"), ns)
    return ns["pred_annotations"]
