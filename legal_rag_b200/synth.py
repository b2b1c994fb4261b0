"""Seeded synthetic inputs shaped like BASELINE.json's configs, generated on the device so that the
10^7..10^8-row corpora never cross PCIe.  (The reference's own seed convention is 42,
scripts/generate_synthetic_data.py:36; per-config seeds follow SURVEY.md section 8d.)
"""
from __future__ import annotations

import torch


def unit_rows_bf16(n: int, d: int, seed: int, device, chunk: int = 1 << 20) -> torch.Tensor:
    """[n, d] bf16, rows L2-normalised in fp32 before rounding (BGE embeddings are unit norm,
    legalrag/retrieval/vector_store.py:154)."""
    out = torch.empty((n, d), dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, d), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        out[s:e] = x.to(torch.bfloat16)
    return out


def unit_tokens_bf16(n: int, L: int, d: int, seed: int, device, chunk: int = 1 << 14) -> torch.Tensor:
    out = torch.empty((n, L, d), dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, L, d), generator=g, device=device, dtype=torch.float32)
        out[s:e] = torch.nn.functional.normalize(x, dim=2).to(torch.bfloat16)
    return out


def zipf_cdf(vocab: int, device) -> torch.Tensor:
    p = 1.0 / torch.arange(1, vocab + 1, device=device, dtype=torch.float64)
    return torch.cumsum(p / p.sum(), 0)


def bm25_synthetic_index(n_docs: int, vocab: int, seed: int, device, *, mean_len: float = 40.0, sigma: float = 0.6,
                         min_len: int = 4, max_len: int = 512, chunk_docs: int = 2_000_000,
                         k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25, id_base: int = 0,
                         keep_tf: bool = False):
    """SURVEY 8d config 4, built on the device: Zipf(s=1) term ids, doc length
    clip(round(lognormal(ln mean_len, sigma)), min_len, max_len), tf = multiplicity; term-major CSR with
    ascending doc ids and fp32 impacts from rank_bm25's idf / avgdl rules (oracle/bm25.py).

    Two streaming passes over seeded doc chunks (pass 1 counts df, pass 2 scatters postings to their
    final offsets), so peak memory is the index itself plus one chunk.  Returns (Bm25DeviceIndex, stats)."""
    import math
    from .engine import Bm25DeviceIndex
    cdf = zipf_cdf(vocab, device)
    mu = math.log(mean_len)
    n_chunks = (n_docs + chunk_docs - 1) // chunk_docs

    def gen_chunk(c):
        c0, c1 = c * chunk_docs, min(n_docs, (c + 1) * chunk_docs)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + c)
        lens = torch.exp(torch.randn(c1 - c0, generator=g, device=device) * sigma + mu).round().clamp_(min_len, max_len).to(torch.int64)
        ntok = int(lens.sum())
        u = torch.rand(ntok, generator=g, device=device, dtype=torch.float64)
        term = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
        del u
        doc = torch.repeat_interleave(torch.arange(c1 - c0, device=device), lens)
        key = term * (c1 - c0) + doc
        del term, doc
        key, _ = torch.sort(key)
        uniq, tf = torch.unique_consecutive(key, return_counts=True)
        del key
        t = torch.div(uniq, c1 - c0, rounding_mode="floor")
        d = uniq - t * (c1 - c0)
        return c0, lens, t, d, tf

    df_chunks = torch.zeros((n_chunks, vocab), dtype=torch.int32, device=device)
    total_tokens = 0
    for c in range(n_chunks):
        _, lens, t, _, _ = gen_chunk(c)
        df_chunks[c] = torch.bincount(t, minlength=vocab).to(torch.int32)
        total_tokens += int(lens.sum())
    df = df_chunks.sum(0, dtype=torch.int64)
    indptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(df, 0)
    nnz = int(indptr[-1])
    avgdl = total_tokens / n_docs
    dff = df.to(torch.float64)
    present = dff > 0
    idf = torch.zeros(vocab, dtype=torch.float64, device=device)
    idf[present] = torch.log(n_docs - dff[present] + 0.5) - torch.log(dff[present] + 0.5)
    average_idf = float(idf[present].sum() / max(1, int(present.sum())))
    idf[present & (idf < 0)] = epsilon * average_idf
    chunk_base = torch.cumsum(df_chunks.to(torch.int64), 0) - df_chunks.to(torch.int64)      # postings of earlier chunks, per term

    doc_id = torch.empty(nnz, dtype=torch.int32, device=device)
    impact = torch.empty(nnz, dtype=torch.float32, device=device)
    tf_all = torch.empty(nnz, dtype=torch.int32, device=device) if keep_tf else None
    len_all = torch.empty(n_docs, dtype=torch.int32, device=device) if keep_tf else None
    for c in range(n_chunks):
        c0, lens, t, d, tf = gen_chunk(c)
        local_ptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
        local_ptr[1:] = torch.cumsum(df_chunks[c].to(torch.int64), 0)
        rank = torch.arange(t.numel(), device=device) - local_ptr[t]
        pos = indptr[t] + chunk_base[c][t] + rank
        f = tf.to(torch.float64)
        dl = lens[d].to(torch.float64)
        imp = idf[t] * (f * (k1 + 1) / (f + k1 * (1 - b + b * dl / avgdl)))
        doc_id[pos] = (d + c0).to(torch.int32)
        impact[pos] = imp.to(torch.float32)
        if keep_tf:
            tf_all[pos] = tf.to(torch.int32)
            len_all[c0:c0 + lens.numel()] = lens.to(torch.int32)
        del t, d, tf, rank, pos, f, dl, imp
    nonneg = bool((idf >= 0).all())
    index = Bm25DeviceIndex(indptr, doc_id, impact, n_docs, nonneg, id_base)
    stats = {"n_docs": n_docs, "vocab": vocab, "nnz": nnz, "tokens": total_tokens, "avgdl": avgdl, "average_idf": average_idf,
             "df": df, "tf": tf_all, "doc_len": len_all}
    return index, stats


def bm25_synthetic_queries(nq: int, vocab: int, seed: int, device, min_terms: int = 2, max_terms: int = 8):
    """nq queries of U{min..max} iid Zipf terms -> (q_indptr int64 [nq+1], q_term int32, max terms)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.randint(min_terms, max_terms + 1, (nq,), generator=g, device=device)
    q_indptr = torch.zeros(nq + 1, dtype=torch.int64, device=device)
    q_indptr[1:] = torch.cumsum(lens, 0)
    u = torch.rand(int(lens.sum()), generator=g, device=device, dtype=torch.float64)
    q_term = torch.searchsorted(zipf_cdf(vocab, device), u).clamp_(max=vocab - 1).to(torch.int32)
    return q_indptr, q_term, max_terms
