"""Oracle: hybrid fusion.  TEST INFRASTRUCTURE.

PINNED against the reference itself: tests/golden/fuse_golden.json holds outputs
of the unmodified ``HybridRetriever._fuse`` (legalrag/retrieval/hybrid_retriever.py:
389-551, helpers ``_minmax`` :24-30 and ``_rrf_with_breakdown`` :33-56) executed in
this container by oracle/make_golden.py; tests/test_oracle.py checks this
restatement against every case in that file.

The restatement works on integer doc ids and float64, exactly as the reference's
Python does.  One deliberate difference: the reference orders equal fused scores
by Python ``set`` iteration order (hybrid_retriever.py:460,484,526 -- it varies
with PYTHONHASHSEED); here ties are ordered by ascending id so results are
reproducible.  Parity checks treat tied runs as sets.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

METHODS = ("weighted_sum", "rrf", "wrrf", "rrf_norm_blend")
CHANNELS = ("dense", "bm25", "colbert")


def minmax(scores: Sequence[float]) -> List[float]:
    """hybrid_retriever.py:24-30."""
    if not scores:
        return []
    lo, hi = min(scores), max(scores)
    if hi - lo < 1e-12:
        return [0.0 for _ in scores]
    return [(float(s) - lo) / (hi - lo) for s in scores]


def rrf_with_breakdown(rank_lists: Dict[str, List[int]], *, k: int = 60,
                       weights: Optional[Dict[str, float]] = None):
    """hybrid_retriever.py:33-56."""
    totals: Dict[int, float] = {}
    contrib: Dict[int, Dict[str, float]] = {}
    weights = weights or {}
    for channel, ids in rank_lists.items():
        w = float(weights.get(channel, 1.0))
        for rank, cid in enumerate(ids, start=1):
            v = w * (1.0 / (k + rank))
            totals[cid] = totals.get(cid, 0.0) + v
            contrib.setdefault(cid, {})[channel] = v
    return totals, contrib


def fuse(dense: Sequence[Tuple[int, float]], bm25: Sequence[Tuple[int, float]],
         colbert: Sequence[Tuple[int, float]], *, method: str = "rrf_norm_blend",
         w_dense: float = 0.6, w_bm25: float = 0.4, w_colbert: float = 0.35,
         rrf_k: int = 60, alpha: float = 0.5) -> List[Dict]:
    """hybrid_retriever.py:389-551 on (id, score) lists.

    Returns dicts {id, score, rrf_norm, weighted_sum, dense_norm, bm25_norm,
    colbert_norm, channel_contrib} sorted by (score desc, id asc).
    """
    method = method.lower()
    weights = {"dense": float(w_dense), "bm25": float(w_bm25), "colbert": float(w_colbert)}
    lists = {"dense": list(dense), "bm25": list(bm25), "colbert": list(colbert)}
    # :409-411 stable sort desc
    for ch in lists:
        lists[ch] = sorted(lists[ch], key=lambda p: float(p[1]), reverse=True)
    rank_lists = {ch: [p[0] for p in lists[ch]] for ch in CHANNELS}

    norm_map: Dict[str, Dict[int, float]] = {}
    for ch in CHANNELS:                     # :432-443
        vals = minmax([float(p[1]) for p in lists[ch]])
        norm_map[ch] = {p[0]: float(vals[i]) for i, p in enumerate(lists[ch])}

    if method == "wrrf":                    # :446-449
        rrf_total, rrf_raw = rrf_with_breakdown(rank_lists, k=rrf_k, weights=weights)
    else:
        rrf_total, rrf_raw = rrf_with_breakdown(rank_lists, k=rrf_k)

    rrf_norm_map: Dict[int, float] = {}
    if rrf_total:                           # :452-457
        items = list(rrf_total.items())
        vals = minmax([float(v) for _, v in items])
        for i, (cid, _) in enumerate(items):
            rrf_norm_map[cid] = float(vals[i])

    all_ids = set(rrf_total)
    for m in norm_map.values():
        all_ids |= set(m)

    def rrf_alloc(cid, mass):               # :471-476
        raw = rrf_raw.get(cid, {}) or {}
        total = float(rrf_total.get(cid, 0.0))
        if mass <= 0.0 or total <= 1e-18:
            return {}
        return {ch: mass * float(v) / total for ch, v in raw.items()}

    out = []
    for cid in all_ids:                     # :484-514
        norms = {ch: float(norm_map[ch].get(cid, 0.0)) for ch in weights}
        w_terms = {ch: weights[ch] * norms[ch] for ch in weights}
        wsum = sum(w_terms.values())
        rrf_norm = float(rrf_norm_map.get(cid, 0.0))
        contrib = {ch: 0.0 for ch in weights}
        if method == "weighted_sum":
            score = float(wsum)
            contrib.update(w_terms)
        elif method in ("rrf", "wrrf"):
            score = float(rrf_norm)
            contrib.update(rrf_alloc(cid, score))
        else:
            score = alpha * rrf_norm + (1.0 - alpha) * wsum
            for ch, v in w_terms.items():
                contrib[ch] += (1.0 - alpha) * v
            for ch, v in rrf_alloc(cid, alpha * rrf_norm).items():
                contrib[ch] = contrib.get(ch, 0.0) + v
        out.append({
            "id": cid, "score": float(score), "rrf_norm": rrf_norm, "weighted_sum": float(wsum),
            "dense_norm": norms["dense"], "bm25_norm": norms["bm25"], "colbert_norm": norms["colbert"],
            "channel_contrib": contrib,
        })
    out.sort(key=lambda r: (-r["score"], r["id"]))
    return out


def fuse_topk(dense, bm25, colbert, k: int, min_final: float = float("-inf"), **kw) -> List[Tuple[int, float]]:
    """fuse -> ``score >= min_final_score`` filter (hybrid_retriever.py:309-310) -> first k."""
    rows = [r for r in fuse(dense, bm25, colbert, **kw) if r["score"] >= min_final]
    return [(r["id"], r["score"]) for r in rows[:k]]
