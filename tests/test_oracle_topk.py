"""The top-k oracle against brute force (CPU)."""
import numpy as np

from oracle import ref_topk


def test_topk_oracle_orders_by_score_then_column():
    rs = np.random.RandomState(0)
    scores = rs.randint(0, 6, size=(5, 40)).astype(np.float32)          # many ties
    seen = [rs.choice(40, size=7, replace=False) for _ in range(5)]
    cols, vals = ref_topk.topk(scores, 12, seen)
    for b in range(5):
        pairs = sorted(((-scores[b, c], c) for c in range(40) if c not in set(seen[b].tolist())))[:12]
        assert [c for _, c in pairs] == cols[b].tolist()
        assert [-v for v, _ in pairs] == vals[b].tolist()


def test_topk_oracle_pads_when_k_exceeds_the_catalogue():
    scores = np.array([[2.0, 1.0, 3.0]], dtype=np.float32)
    cols, vals = ref_topk.topk(scores, 5, [[1]])
    assert cols.tolist() == [[2, 0, -1, -1, -1]]
    assert vals[0, :2].tolist() == [3.0, 2.0] and np.all(np.isneginf(vals[0, 2:]))
