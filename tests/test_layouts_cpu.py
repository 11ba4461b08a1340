"""CPU tests of the layout steps either side of the hot path: the oracle's restatement of the
reference's head concat (model.py:295-320) and GT padding (inputs.py:340-348), and the host
helpers that produce the per-head / ragged formats the CUDA entry points consume."""
import numpy as np

from multibox_b200 import loss, synth
from multibox_b200.priors import num_priors
from oracle import np_oracle


def test_head_priors_follow_the_reference_grids():
    # 129*K + 1 priors (reference config: K=5 -> 646, pinned by model_tests.py:15)
    for K in (5, 7, 11):
        hp = loss.head_priors(K)
        assert hp == [64 * K, 36 * K, 16 * K, 9 * K, 4 * K, 1]
        assert sum(hp) == num_priors(K) == 129 * K + 1


def test_split_and_concat_heads_round_trip():
    d = synth.make_train_inputs(K=5, B=3, M=20, seed=1)
    hl, hc = synth.split_heads(d["locations"], d["logits"], 5)
    assert [t.shape for t in hl] == [(3, 8, 8, 20), (3, 6, 6, 20), (3, 4, 4, 20), (3, 3, 3, 20), (3, 2, 2, 20),
                                     (3, 1, 1, 4)]
    assert [t.shape for t in hc] == [(3, 8, 8, 5), (3, 6, 6, 5), (3, 4, 4, 5), (3, 3, 3, 5), (3, 2, 2, 5), (3, 1, 1, 1)]
    loc, conf = np_oracle.concat_heads(hl, hc)
    assert np.array_equal(loc, d["locations"]) and np.array_equal(conf, d["logits"])
    # prior p of head h, cell (i,j), box k sits at flat index (i*g + j)*K + k: the order
    # priors.generate_priors emits (priors.py:261-267: cells row-major, ratios innermost)
    g, K = 8, 5
    i, j, k = 3, 6, 2
    p = (i * g + j) * K + k
    assert np.array_equal(hl[0][1, i, j, 4 * k:4 * k + 4], d["locations"][1, p])


def test_ragged_and_padded_gt_round_trip():
    d = synth.make_train_inputs(K=5, B=9, M=20, seed=2, edge_cases=True)
    flat, off = synth.ragged_gt(d["gt"], d["num_gt"])
    assert off[0] == 0 and off[-1] == flat.shape[0] == d["num_gt"].sum()
    gt, num = np_oracle.pad_ragged_gt(flat, off, 20)
    assert np.array_equal(num, d["num_gt"])
    for b in range(9):
        assert np.array_equal(gt[b, :num[b]], d["gt"][b, :num[b]])
        assert not gt[b, num[b]:].any()
