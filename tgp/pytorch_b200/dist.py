"""Row sharding of a global minibatch across ranks: the host-side arithmetic.

Rank r owns the contiguous slice [r*MB/G, (r+1)*MB/G) of the global batch (the same explicit index tensor on every
rank); the ELL scale stays N / MB_global on every rank.  The exchange itself lives in `functional.allreduce_packed`
(one sum all-reduce of the tril-packed pre-chain buffer per step) and `functional.synced_module_call` (gradients of the
input-dependent flow MLPs); `sparse_MF_SP._global_scale` and `bench.py` use the helpers below.
"""


def shard_bounds(n_rows, world):
    """[lo_0, lo_1, ..., lo_world] with lo_r = floor(r * n_rows / world)."""
    return [(r * n_rows) // world for r in range(world + 1)]


def local_slice(n_rows, rank, world):
    b = shard_bounds(n_rows, world)
    return slice(b[rank], b[rank + 1])


def global_scale(N, n_rows_global):
    """N / MB of sparse_MF_SP.ELL (reference sparse_MF_SP.py:626) with MB = rows of the GLOBAL minibatch."""
    return float(N) / float(n_rows_global)
