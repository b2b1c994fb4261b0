"""BM25Retriever (replaces legalrag/retrieval/bm25_retriever.py:17-76): bm25.pkl is unpickled (with a shim
when rank_bm25 is not installed), converted once to device-resident CSR postings, and every search is the
liblrag BM25 kernel -- scoring and the full ranking the reference does in Python."""
from __future__ import annotations

import logging
import threading
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import engine
from ..bm25_index import Bm25HostIndex
from ..schemas import LawChunk, chunk_from_obj
from . import artifacts, encoders


def compact_known_terms(q_indptr: np.ndarray, q_term: np.ndarray, cap: int):
    """Drops the out-of-vocabulary tokens (-1) of every query, then keeps at most `cap` tokens per query (the first ones).
    Tokens outside the vocabulary score nothing in rank_bm25 (`self.doc_freqs[...].get(q) or 0`, bm25_retriever.py:74), so
    they must not take up slots of the kernel's per-query term table -- jieba emits whitespace and punctuation as tokens
    of their own.  -> (q_indptr int64 [nq + 1], q_term int32, longest kept query, longest query before the cap)."""
    q_indptr = np.asarray(q_indptr, dtype=np.int64)
    q_term = np.asarray(q_term, dtype=np.int32)
    known = q_term >= 0
    seen = np.concatenate([[0], np.cumsum(known)]).astype(np.int64)
    lens = seen[q_indptr[1:]] - seen[q_indptr[:-1]]
    q_term = q_term[known]
    q_indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    longest = int(lens.max()) if lens.size else 0
    if longest > cap:
        keep = np.concatenate([np.arange(a, min(b, a + cap)) for a, b in zip(q_indptr[:-1], q_indptr[1:])]).astype(np.int64)
        q_term = q_term[keep]
        lens = np.minimum(lens, cap)
        q_indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return q_indptr, q_term, min(longest, cap), longest


class BM25Retriever:
    def __init__(self, cfg, tokenizer: Optional[Callable[[str], List[str]]] = None):
        self.cfg = cfg
        self.bm25_path = Path(cfg.retrieval.bm25_index_file)
        self.device = torch.device(getattr(cfg, "device", None) or "cuda")
        self.tokenizer = tokenizer or encoders.default_query_tokenizer()
        self._loaded = False
        self._load_lock = threading.Lock()
        self._bm25_mtime: Optional[float] = None
        self.bm25 = None                          # the unpickled BM25Okapi (or its stand-in)
        self.chunks: List[LawChunk] = []
        self.host_index: Optional[Bm25HostIndex] = None
        self.device_index: Optional[engine.Bm25DeviceIndex] = None

    def load(self) -> None:
        if not self.bm25_path.exists():
            raise RuntimeError(f"[BM25] index not found: {self.bm25_path}. "
                               f"Run: python build_index.py (or python build_index.py --only-bm25)")
        current_mtime = self.bm25_path.stat().st_mtime
        if self._loaded and self._bm25_mtime == current_mtime:
            return
        with self._load_lock:                    # threads that arrive together load once and share the snapshot
            if not (self._loaded and self._bm25_mtime == current_mtime):
                self._load_locked(current_mtime)

    def _load_locked(self, current_mtime: float) -> None:
        obj = artifacts.read_bm25_pickle(self.bm25_path)
        bm25 = obj.get("bm25")
        if bm25 is None:
            raise RuntimeError(f"[BM25] invalid index file (missing 'bm25'): {self.bm25_path}")
        try:
            chunks = [chunk_from_obj(c) for c in obj.get("chunks", [])]
        except RuntimeError as e:
            raise RuntimeError(f"[BM25] {e}") from e
        host = Bm25HostIndex.from_okapi(bm25)
        dev = host.to_device(self.device)
        # publish the finished snapshot in one step
        self.bm25, self.chunks, self.host_index, self.device_index = bm25, chunks, host, dev
        self._loaded, self._bm25_mtime = True, current_mtime

    def search_ids(self, token_lists: Sequence[Sequence[str]], top_k: int):
        """Pre-tokenised queries -> (scores [nq, k], doc rows [nq, k]) on the device."""
        self.load()
        host, dev = self.host_index, self.device_index
        k = max(1, min(int(top_k), engine.LRAG_MAX_K))
        qi, qt, _ = host.encode_queries([list(t) for t in token_lists])
        qi, qt, mx, longest = compact_known_terms(qi, qt, engine.LRAG_BM25_MAX_QUERY_TERMS)
        if longest > engine.LRAG_BM25_MAX_QUERY_TERMS:
            # rank_bm25 scores every token (bm25_retriever.py:74); the kernel's per-query term table holds 128
            logging.getLogger("legal_rag_b200.retrieval").warning(
                "[BM25] a query of %d in-vocabulary tokens is scored on its first %d (kernel limit LRAG_BM25_MAX_QUERY_TERMS); "
                "the reference would score all of them", longest, engine.LRAG_BM25_MAX_QUERY_TERMS)
        return engine.bm25_topk(dev, torch.from_numpy(qi).to(self.device), torch.from_numpy(qt).to(self.device), mx, k)

    def search(self, query: str, top_k: int) -> List[Tuple[LawChunk, float]]:
        self.load()
        n = len(self.chunks)
        s, i = self.search_ids([self.tokenizer(query)], top_k)
        s, i = s[0].tolist(), i[0].tolist()
        return [(self.chunks[d], float(sc)) for sc, d in zip(s, i) if 0 <= d < n][: int(top_k)]

    def search_batch(self, queries: Sequence[str], top_k: int) -> List[List[Tuple[LawChunk, float]]]:
        self.load()
        n = len(self.chunks)
        s, i = self.search_ids([self.tokenizer(q) for q in queries], top_k)
        s, i = s.tolist(), i.tolist()
        return [[(self.chunks[d], float(sc)) for sc, d in zip(rs, ri) if 0 <= d < n] for rs, ri in zip(s, i)]
