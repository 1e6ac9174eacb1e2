"""Row sharding of a global minibatch across ranks: host-side helpers.

Rank r owns the contiguous slice [r*MB/G, (r+1)*MB/G) of the global batch (the same explicit index tensor on every
rank); the ELL scale stays N / MB_global on every rank; the only exchange is a sum all-reduce of the packed pre-chain
buffer.  These helpers are pure host logic (CPU-testable with gloo); the kernels never see ranks.
"""
import torch


def shard_bounds(n_rows, world):
    """[lo_0, lo_1, ..., lo_world] with lo_r = floor(r * n_rows / world)."""
    return [(r * n_rows) // world for r in range(world + 1)]


def local_slice(n_rows, rank, world):
    b = shard_bounds(n_rows, world)
    return slice(b[rank], b[rank + 1])


def global_scale(N, n_rows_global):
    return float(N) / float(n_rows_global)


def pack(tensors):
    """Flattens tensors into one contiguous FP64 vector (the object that is all-reduced once per step)."""
    return torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])


def unpack(buf, like):
    out, o = [], 0
    for t in like:
        n = t.numel()
        out.append(buf[o:o + n].reshape(t.shape))
        o += n
    return out


def allreduce_sum_(buf):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    return buf
