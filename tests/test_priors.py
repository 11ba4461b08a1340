"""CPU tests: the product's host-side prior generator (multibox_b200/priors.py)
is bit-identical in float64 to the reference's generate_priors (golden file
written from reference priors.py:185-314)."""
import os

import numpy as np
import pytest

from multibox_b200 import priors


@pytest.mark.parametrize("K", [5, 7, 11])
def test_generate_priors_bit_equal(golden_dir, K):
    g = np.load(os.path.join(golden_dir, "priors.npz"))
    ours = priors.generate_priors(g["ratios%d" % K].tolist())
    assert isinstance(ours, list) and len(ours) == 129 * K + 1 == priors.num_priors(K)
    assert all(isinstance(b, list) and len(b) == 4 for b in ours[:3])
    assert np.array_equal(np.array(ours, dtype=np.float64), g["K%d" % K])


def test_generate_priors_unrestricted(golden_dir):
    g = np.load(os.path.join(golden_dir, "priors.npz"))
    ours = priors.generate_priors(g["ratios5"].tolist(), 0.2, 0.9, False)
    assert np.array_equal(np.array(ours), g["K5_unrestricted_0.2_0.9"])


def test_known_values_k5():
    p = priors.priors_fp32([1, 2, 3, 1 / 2., 1 / 3.])
    assert p.dtype == np.float32 and p.shape == (646, 4)          # reference model_tests.py:15
    np.testing.assert_allclose(p[0], [.0125, .0125, .1125, .1125], rtol=0, atol=1e-7)
    np.testing.assert_allclose(p[-1], [.025, .025, .975, .975], rtol=0, atol=1e-7)
    assert (p >= 0).all() and (p <= 1).all()
    assert ((p[:, 2] - p[:, 0]) > 0).all() and ((p[:, 3] - p[:, 1]) > 0).all()
