"""Prior-box generation -- host side, one-time, float64.

Mirrors the interface of the reference's ``priors.generate_priors``
(reference priors.py:185-314): same arguments, same defaults, same return type
(a python list of ``[x1, y1, x2, y2]`` float lists, length ``129*K + 1``), same
ordering (grids 8,6,4,3,2 row-major cell order with the aspect ratios innermost,
then the single square 1x1 prior), bit-identical float64 values.

It is vectorised over (grid cell, aspect ratio) instead of the reference's
four nested python loops; every arithmetic step is an IEEE-754 correctly
rounded float64 op in the same order, so results are bit-equal
(tests/test_priors.py checks this against the golden file generated from the
reference's own code).
"""
import numpy as np

GRIDS = (8, 6, 4, 3, 2, 1)   # one entry per detection head (reference priors.py:196)


def num_priors(num_aspect_ratios):
    """129*K + 1 (K=5 -> 646, the constant pinned by reference model_tests.py:15)."""
    return sum(g * g for g in GRIDS[:-1]) * int(num_aspect_ratios) + 1


def _boxes_for_grid(grid, scale, ratios, restrict):
    """All priors of one grid as float64 [grid*grid*len(ratios), 4]."""
    a = np.asarray(ratios, dtype=np.float64)
    w = scale * np.sqrt(a)                       # reference priors.py:269
    h = scale / np.sqrt(a)                       # :270
    centres = (np.arange(grid, dtype=np.float64) + 0.5) / grid   # :264-265
    ci = centres[:, None, None]                  # row index i -> y
    cj = centres[None, :, None]                  # col index j -> x
    shape = (grid, grid, a.shape[0])
    x1 = np.broadcast_to(cj - (w / 2.), shape).copy()
    x2 = np.broadcast_to(cj + (w / 2.), shape).copy()
    y1 = np.broadcast_to(ci - (h / 2.), shape).copy()
    y2 = np.broadcast_to(ci + (h / 2.), shape).copy()
    if restrict:
        # overhang on each side (:278-281), largest one (:283-286)
        trim = np.maximum(np.maximum(np.abs(np.minimum(0., x1)), np.abs(np.minimum(0., 1 - x2))),
                          np.maximum(np.abs(np.minimum(0., y1)), np.abs(np.minimum(0., 1 - y2))))
        tall = np.broadcast_to(h > w, shape)
        width_trim = np.where(tall, trim * a, trim)      # :288-293
        height_trim = np.where(tall, trim, trim / a)
        xa, xb = x1 + width_trim, x2 - width_trim
        ya, yb = y1 + height_trim, y2 - height_trim
        x1, x2 = np.minimum(xa, xb), np.maximum(xa, xb)  # :300-303
        y1, y2 = np.minimum(ya, yb), np.maximum(ya, yb)
    box = np.stack([np.maximum(x1, 0.), np.maximum(y1, 0.),
                    np.minimum(x2, 1.), np.minimum(y2, 1.)], axis=-1)   # :305-310
    return box.reshape(-1, 4)


def generate_priors_array(aspect_ratios, min_scale=0.1, max_scale=0.95,
                          restrict_to_image_bounds=True):
    """float64 ndarray [129*K+1, 4]."""
    n = len(GRIDS)
    out = []
    for k, grid in enumerate(GRIDS):
        # same expression as the reference (:200) so the scale bits agree
        scale = min_scale + (max_scale - min_scale) * ((k + 1) - 1) / (n - 1)
        ratios = [1.] if grid == 1 else list(aspect_ratios)
        out.append(_boxes_for_grid(grid, scale, ratios, restrict_to_image_bounds))
    return np.concatenate(out, axis=0)


def generate_priors(aspect_ratios, min_scale=0.1, max_scale=0.95,
                    restrict_to_image_bounds=True):
    """Drop-in for reference priors.py:185: returns a list of 4-element lists."""
    return generate_priors_array(aspect_ratios, min_scale, max_scale,
                                 restrict_to_image_bounds).tolist()


def priors_fp32(aspect_ratios, **kw):
    """What the reference's drivers feed the graph: ``np.array(priors).astype(np.float32)``
    (reference train.py:368-370)."""
    return generate_priors_array(aspect_ratios, **kw).astype(np.float32)
