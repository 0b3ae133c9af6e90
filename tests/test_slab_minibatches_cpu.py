"""The zero-copy minibatch claim, checked on the CPU against the oracle restatement of the reference's
sort_training_data / flatten_batch (clean_pufferl.py:452-482), which is itself pinned to reference outputs
(tests/test_oracle_golden.py): the slab-major minibatches hold exactly the reference minibatches' rows."""
import numpy as np
import pytest

from oracle import experience as oexp
from pufferlib_b200 import clean_pufferl


@pytest.mark.parametrize('n,h,mb,bptt', [(64, 128, 2048, 16), (16, 64, 256, 8), (4, 32, 64, 8), (7, 24, 84, 4),
                                         (16384, 128, 524288, 16), (64, 32, 1024, 8), (3, 8, 24, 8)])
def test_slab_minibatches_are_the_reference_minibatches(n, h, mb, bptt):
    b = n * h
    ora = oexp.Experience(b, bptt, mb, (1,), np.float32)
    nm = ora.num_minibatches
    layout = clean_pufferl.slab_layout(n, h, nm, bptt)
    assert layout is not None
    g_, r_ = layout
    assert g_ * r_ == mb
    if b <= 1 << 16:                   # the reference's python sort is O(B log B) on tuples: small cases only
        ora.sort_keys = [(e, t) for t in range(h) for e in range(n)]
        ora.sort_training_data()
        ref_rows = ora.b_idxs_flat                              # [n_mb, mb_size] arrival rows, reference order
    else:                              # same permutation, written arithmetically (sorted f = e*h + t -> row t*n + e)
        f = np.arange(b, dtype=np.int64)
        idxs = (f % h) * n + f // h
        ref_rows = idxs.reshape(ora.minibatch_rows, nm, bptt).transpose(1, 0, 2).reshape(nm, mb)
    ours = clean_pufferl.slab_row_index(n, h, nm, bptt)
    for m in range(nm):
        assert np.array_equal(np.sort(ours[m]), np.sort(ref_rows[m]))
        # each slab is one contiguous run of the time-major buffer
        runs = ours[m].reshape(g_, r_)
        assert np.array_equal(runs, runs[:, :1] + np.arange(r_)[None, :])
        # and the position map used by the GPU tests: slab position (g, j, e) <-> reference position (e, g, j)
        assert np.array_equal(ours[m].reshape(g_, bptt, n).transpose(2, 0, 1).reshape(-1), ref_rows[m])
    assert np.array_equal(np.sort(ours.reshape(-1)), np.arange(b))


@pytest.mark.parametrize('n,h,mb,bptt', [(33, 12, 36, 4), (8, 32, 64, 16), (100, 10, 250, 5), (256, 128, 4096, 32),
                                         (4, 12, 16, 8)])
def test_slab_layout_refuses_shapes_where_minibatches_interleave_envs(n, h, mb, bptt):
    nm = (n * h) // mb
    assert clean_pufferl.slab_layout(n, h, nm, bptt) is None
