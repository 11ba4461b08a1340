"""-m gpu tests of the literal tf.py_func drop-in (multibox_b200/native_boundary.py; reference
loss.py:8-53,81-82): HOST numpy arrays in, [int32 mask, float32 stacked_gt] numpy arrays out, checked
against the golden vectors written from the reference's own code and against the oracle."""
import os

import numpy as np
import pytest

from multibox_b200 import detect, native_boundary, synth, _lib
from oracle import c_oracle, np_oracle
from gpu_util import boundary_inputs, dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_py_func_drop_in_against_reference_golden(cuda_device, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "match_%s.npz" % name))
    d = synth.make_train_inputs(**synth.TRAIN_CONFIGS[name])
    loc, conf = boundary_inputs(d)
    # exactly the py_func parameter list of reference loss.py:81 (batch_size / alpha as 0-d arrays)
    out = native_boundary.compute_assignments(loc, conf, d["gt"], d["num_gt"], np.int32(d["B"]), np.float32(d["alpha"]))
    assert isinstance(out, list) and len(out) == 2
    mask, stacked = out
    assert isinstance(mask, np.ndarray) and mask.dtype == np.int32 and mask.shape == (d["B"] * d["P"],)
    assert isinstance(stacked, np.ndarray) and stacked.dtype == np.float32 and stacked.shape == (int(d["num_gt"].sum()), 4)
    assert np.array_equal(np.nonzero(mask)[0], g["matched_flat_idx"])
    assert np.array_equal(stacked, g["stacked_gt"])


def test_py_func_drop_in_shapes_errors_and_reuse(cuda_device):
    for B, K, M, dist in ((3, 5, 20, "full"), (5, 7, 100, "coco_person"), (3, 5, 20, "uniform")):
        d = synth.make_train_inputs(K=K, B=B, M=M, dist=dist, seed=40 + B + K)
        loc, conf = boundary_inputs(d)
        m0, s0, _ = c_oracle.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
        for _ in range(2):      # the staging buffers are reused on the second call
            m, s = native_boundary.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
            assert np.array_equal(m, m0) and np.array_equal(s, s0)
    # empty ground truth: (0, 4) float32, all-zero mask
    d["num_gt"][:] = 0
    m, s = native_boundary.compute_assignments(loc, conf, d["gt"], d["num_gt"], B, d["alpha"])
    assert m.sum() == 0 and s.shape == (0, 4) and s.dtype == np.float32
    # scipy's failures (reference loss.py:40) come back as the same exception type
    d = synth.make_train_inputs(K=5, B=2, M=20, dist="full", seed=5)
    loc, conf = boundary_inputs(d)
    bad = loc.copy()
    bad[700, 2] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        native_boundary.compute_assignments(bad, conf, d["gt"], d["num_gt"], 2, 1000.0)
    with pytest.raises(ValueError, match="infeasible"):
        native_boundary.compute_assignments(loc, np.zeros_like(conf), d["gt"], d["num_gt"], 2, 1000.0)
    with pytest.raises(ValueError, match="expected locations"):
        native_boundary.compute_assignments(loc[:-1], conf, d["gt"], d["num_gt"], 2, 1000.0)
    # the inputs are not retained or modified (TF owns them)
    loc2, conf2 = loc.copy(), conf.copy()
    native_boundary.compute_assignments(loc2, conf2, d["gt"], d["num_gt"], 2, 1000.0)
    assert np.array_equal(loc2, loc) and np.array_equal(conf2, conf)


def test_k_max_beyond_capacity_is_an_argument_error(cuda_device):
    """The reference has no cap on max_to_keep; this kernel family handles 1024 detections per image.
    More is an error (MBX_E_TOO_LARGE), never a silent clamp."""
    q = synth.make_detect_inputs(K=11, B=2, keep=200, seed=3)
    args = (dev(q["locations"]), dev(q["confidences"]), dev(q["priors"]))
    with pytest.raises(_lib.MultiboxLibraryError, match="1024"):
        detect.postprocess(*args, k_max=1025)
    import torch
    mk = torch.full((2, 1), 1400, dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.MultiboxLibraryError, match="1024"):
        detect.postprocess(*args, max_to_keep=mk)
    out = detect.postprocess(*args, k_max=1024)
    assert out["scores"].shape == (2, 1024)
