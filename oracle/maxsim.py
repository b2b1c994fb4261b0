"""Oracle: exact ColBERT MaxSim late interaction.  TEST INFRASTRUCTURE.

parity unpinned: colbert-ai 0.2.22 (requirements.txt:12, pyproject.toml:45) is not
vendored or installed.  Restates ``colbert.modeling.colbert.colbert_score`` /
``colbert_score_reduce`` (cosine variant) as reached from
legalrag/retrieval/colbert_retriever.py:152:

    score(q, d) = sum_{i < Lq} max_{j < doclen[d]} <q_i, d_j>

on L2-normalised token vectors, padded doc tokens masked out of the max (the
upstream code fills them with -9999).  The real library additionally scores
4-bit-residual-decompressed tokens over a PLAID-pruned candidate set; the oracle
is exact MaxSim by definition.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

PAD_FILL = -9999.0


def maxsim_scores(Q: np.ndarray, D: np.ndarray, doclen: Optional[np.ndarray], cand: np.ndarray) -> np.ndarray:
    """Q [nq,Lq,dim] f32, D [Nd,Ld,dim] f32, cand [nq,C] rows (-1 = skip) -> [nq,C] f32.

    Skipped candidates score -inf.
    """
    Q = np.asarray(Q, dtype=np.float32)
    nq, C = cand.shape
    Ld = D.shape[1]
    out = np.full((nq, C), -np.inf, dtype=np.float32)
    for q in range(nq):
        rows = cand[q]
        ok = np.nonzero(rows >= 0)[0]
        if ok.size == 0:
            continue
        Dq = np.asarray(D[rows[ok]], dtype=np.float32)            # [c, Ld, dim]
        S = np.einsum("ld,ctd->clt", Q[q], Dq, optimize=True)    # [c, Lq, Ld]
        if doclen is not None:
            mask = np.arange(Ld)[None, :] >= np.asarray(doclen)[rows[ok]][:, None]
            S = np.where(mask[:, None, :], np.float32(PAD_FILL), S)
        out[q, ok] = S.max(axis=2).sum(axis=1)
    return out


def rerank_topk(Q, D, doclen, cand, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of the candidate scores, (score desc, id asc); pads with (-inf, -1)."""
    S = maxsim_scores(Q, D, doclen, cand)
    nq, C = cand.shape
    out_s = np.full((nq, k), -np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        ok = np.nonzero(cand[q] >= 0)[0]
        order = ok[np.lexsort((cand[q, ok], -S[q, ok].astype(np.float64)))][:k]
        out_s[q, :len(order)] = S[q, order]
        out_i[q, :len(order)] = cand[q, order]
    return out_s, out_i
