"""Oracle: exact flat inner-product top-k.  TEST INFRASTRUCTURE.

parity unpinned: faiss-cpu 1.13.2 (requirements.txt:7, pyproject.toml:36) is not
vendored or installed.  Restates ``faiss.IndexFlat(METRIC_INNER_PRODUCT).search``
as called at legalrag/retrieval/dense_retriever.py:42 and
legalrag/retrieval/vector_store.py:169: fp32 ``D, I = topk_k(Q @ X.T)``,
descending, padded with ``(-FLT_MAX, -1)`` when k > ntotal.  faiss leaves tie
order unspecified (binary heap); this oracle fixes it to (score desc, id asc),
which is also the product's rule, and the parity helper in tests/parity.py only
compares ids at positions the scores decide.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

FLT_LOWEST = np.float32(-3.4028234663852886e38)


def topk_rows(S: np.ndarray, k: int, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Row-wise exact top-k of a score matrix with (score desc, id asc) order."""
    nq, N = S.shape
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, N)
    if kk == 0:
        return D, I
    for r in range(nq):
        s = S[r]
        if kk < N:
            # keep everything >= the kk-th largest so ties at the cut are resolved by id
            kth = np.partition(s, N - kk)[N - kk]
            cand = np.nonzero(s >= kth)[0]
        else:
            cand = np.arange(N)
        order = cand[np.lexsort((cand, -s[cand].astype(np.float64)))][:kk]
        D[r, :kk] = s[order]
        I[r, :kk] = order + id_base
    return D, I


def flat_ip_topk(Q: np.ndarray, X: np.ndarray, k: int, id_base: int = 0,
                 chunk: int = 262144) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 GEMM + exact top-k, streaming the corpus in row chunks."""
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    nq = Q.shape[0]
    N = X.shape[0]
    best_s = np.full((nq, 0), 0, dtype=np.float32)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    for s0 in range(0, max(N, 1), chunk):
        Xc = np.asarray(X[s0:s0 + chunk], dtype=np.float32)
        if Xc.shape[0] == 0:
            break
        S = Q @ Xc.T
        d, i = topk_rows(S, min(k, Xc.shape[0]), id_base=s0)
        best_s = np.concatenate([best_s, d], axis=1)
        best_i = np.concatenate([best_i, i], axis=1)
        if best_s.shape[1] > 4 * k:
            best_s, best_i = _reduce(best_s, best_i, k)
    best_s, best_i = _reduce(best_s, best_i, k)
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    kk = best_s.shape[1]
    D[:, :kk] = best_s
    I[:, :kk] = np.where(best_i >= 0, best_i + id_base, -1)
    return D, I


def _reduce(s: np.ndarray, i: np.ndarray, k: int):
    kk = min(k, s.shape[1])
    out_s = np.empty((s.shape[0], kk), dtype=np.float32)
    out_i = np.empty((s.shape[0], kk), dtype=np.int64)
    for r in range(s.shape[0]):
        order = np.lexsort((i[r], -s[r].astype(np.float64)))[:kk]
        out_s[r] = s[r, order]
        out_i[r] = i[r, order]
    return out_s, out_i


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k-way merge of per-shard candidate lists [nq, G*k]; id == -1 is padding."""
    nq = scores.shape[0]
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        valid = np.nonzero(ids[r] >= 0)[0]
        order = valid[np.lexsort((ids[r, valid], -scores[r, valid].astype(np.float64)))][:k]
        D[r, :len(order)] = scores[r, order]
        I[r, :len(order)] = ids[r, order]
    return D, I
