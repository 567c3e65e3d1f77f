"""Host-side scatter of independent LZFSE streams over the GPUs of one box.

Streams share nothing (no dictionary, tables are per block, the window is per frame), so multi-GPU
is a partition of the descriptor array: contiguous ranges balanced by bytes moved.  No collective."""
import numpy as np


def shard_ranges(weights, world_size):
    """Split stream indices [0, n) into `world_size` contiguous ranges with near-equal total weight.

    weights[i] = bytes stream i moves (compressed + uncompressed).  Returns a list of (lo, hi)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world_size <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world_size, 1) - 1)
    cum = np.cumsum(w)
    total = cum[-1] if n else 0.0
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        i = int(np.searchsorted(cum, target, side="left"))
        # choose the cut (after i or after i+1 streams) closest to the target
        if i < n and abs(cum[i] - target) <= abs((cum[i - 1] if i > 0 else 0.0) - target):
            i += 1
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]
