"""Index arithmetic of the device-driven sharded exchange, restated on the host (oracle.shard_exchange_*): every
inbox is tiled exactly once, and what comes back for a unique key is that key's (owner, local row)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ref_numpy as R


@settings(max_examples=60, deadline=None)
@given(world=st.integers(1, 8), rows=st.integers(1, 50), seed=st.integers(0, 10_000), n=st.integers(0, 200))
def test_round_trip_returns_each_keys_owner_and_row(world, rows, seed, n):
    rng = np.random.default_rng(seed)
    keys = [rng.integers(0, world * rows, size=rng.integers(0, n + 1)) for _ in range(world)]
    uniq, inbox, landing = R.shard_exchange_simulate(keys, world, rows)
    for o in range(world):
        assert (inbox[o] >= 0).all() and (inbox[o] < rows).all()            # no gap, no overlap, local rows only
    for s in range(world):
        assert (landing[s] >= 0).all()                                        # every unique key was served
        np.testing.assert_array_equal(landing[s][:, 0] * rows + landing[s][:, 1], uniq[s])


def test_offsets_match_the_closed_forms_on_a_fixed_matrix():
    sizes = np.array([[2, 0, 3], [1, 4, 0], [0, 0, 5]])
    bounds = np.concatenate([np.zeros((3, 1), int), np.cumsum(sizes, 1)], 1)
    dst, src, inbox, n_r = R.shard_exchange_offsets(bounds, 1)
    assert dst.tolist() == [2, 1, 0]            # where each rank's bucket for owner 1 starts in its own unique list
    assert src.tolist() == [0, 0, 4, 4] and n_r == 4
    assert inbox.tolist() == [2, 0, 3]          # rank 1's buckets land after rank 0's in every inbox
