"""Dense channel behind the reference's `DenseRetriever(cfg, store=None).search(query, top_k)` interface
(legalrag/retrieval/dense_retriever.py:13-60): embed the query, flat inner-product top-k over the corpus, one
`RetrievalHit` per returned row.

The engine is batch-first: `search_batch` embeds every query, runs ONE scan of the corpus for all of them
(`VectorStore.index` is the faiss-shaped front of `lrag_dense_topk_bf16`) and turns the `[nq, k]` result block into
hits; `search` is the one-query case of it.  Behaviour kept from the reference: `top_k` is clamped to >= 1, rows whose
index is -1 (padding when k > ntotal) or past the chunk list are dropped, `rank` is the 1-based position in the
index's result row (dropped rows leave gaps), `score == semantic_score`, `source == "retriever"`.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ..schemas import RetrievalHit
from .vector_store import VectorStore


class DenseRetriever:
    def __init__(self, cfg: object, store: Optional[VectorStore] = None) -> None:
        self.cfg = cfg
        self.store = store if store is not None else VectorStore.from_config(cfg)

    # ---- result block -> hits --------------------------------------------------------------
    def _row_hits(self, row_scores: np.ndarray, row_ids: np.ndarray) -> List[RetrievalHit]:
        chunks = self.store.chunks
        keep = np.flatnonzero((row_ids >= 0) & (row_ids < len(chunks)))
        out: List[RetrievalHit] = []
        for pos in keep.tolist():
            value = float(row_scores[pos])
            out.append(RetrievalHit(chunk=chunks[int(row_ids[pos])], score=value, semantic_score=value,
                                    rank=pos + 1, source="retriever"))
        return out

    # ---- public API ------------------------------------------------------------------------
    def search_batch(self, queries: Sequence[str], top_k: int) -> List[List[RetrievalHit]]:
        """Every query of the batch against the corpus in one scan (the reference answers one query per call)."""
        self.store.load()
        depth = max(1, int(top_k))
        block_scores, block_ids = self.store.index.search(self.store._embed(list(queries), is_query=True), depth)
        block_scores, block_ids = np.asarray(block_scores), np.asarray(block_ids)
        return [self._row_hits(block_scores[j], block_ids[j]) for j in range(block_ids.shape[0])]

    def search(self, query: str, top_k: int) -> List[RetrievalHit]:
        return self.search_batch([query], top_k)[0]
