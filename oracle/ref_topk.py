"""CPU oracle of the top-k serving epilogue (test infrastructure only: imported by tests/ and bench.py's
checks, never by the product path). The reference has no top-k code: this restates the obvious
definition on top of `full_predictions` (model.py:82-84) - per row the k highest scores, best
first, ties by ascending column, optionally without the columns the row holds as inputs."""
import numpy as np


def topk(scores, k, seen=None):
    """scores [B, N]; seen: list of per-row column arrays to drop (or None).
    Returns (cols int32 [B, k], vals float32 [B, k]); missing slots are -1 / -inf."""
    scores = np.array(scores, dtype=np.float32, copy=True)
    B, N = scores.shape
    if seen is not None:
        for b, cols in enumerate(seen):
            scores[b, np.asarray(cols, dtype=np.int64)] = -np.inf
    out_c = np.full((B, k), -1, dtype=np.int32)
    out_v = np.full((B, k), -np.inf, dtype=np.float32)
    cols = np.arange(N)
    for b in range(B):
        order = np.lexsort((cols, -scores[b].astype(np.float64)))[:k]
        vals = scores[b, order]
        ok = vals != -np.inf
        out_c[b, :len(order)] = np.where(ok, order, -1)
        out_v[b, :len(order)] = vals
    return out_c, out_v
