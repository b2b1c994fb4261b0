"""Self-check of the benchmark workloads, outside the timed region: a sample of the batch's queries is recomputed per
rank with plain torch (fp32 matmul for the dense scan, fp64 scatter-adds for BM25, fp32 einsum for MaxSim, fp64
min-max fusion) from the same device-resident stores, merged across ranks with torch collectives and sorts, and compared
with what the engine returned for those queries.  None of liblrag runs on the checking side, and `oracle/` is not
imported here (bench.py may execute the oracle only in its CPU legs).

Parity rule (tests/parity.py): a position whose checker score is separated from both neighbours by more than
tau must carry the checker's id; every score the two sides share an id for agrees within tau relative.
"""
from __future__ import annotations

import torch


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def sort_lists(s: torch.Tensor, i: torch.Tensor, k: int):
    """rows of (score, id) -> first k by (score desc, id asc); id < 0 entries last."""
    s = torch.where(i >= 0, s, torch.full_like(s, float("-inf")))
    o1 = torch.argsort(i, dim=1, stable=True)
    s1, i1 = torch.gather(s, 1, o1), torch.gather(i, 1, o1)
    o2 = torch.argsort(s1, dim=1, descending=True, stable=True)
    return torch.gather(s1, 1, o2)[:, :k].contiguous(), torch.gather(i1, 1, o2)[:, :k].contiguous()


def gather_lists(s: torch.Tensor, i: torch.Tensor, k: int):
    """all ranks' [n, k] lists -> merged [n, k] (replicated)."""
    dist = _dist()
    if dist is None:
        return sort_lists(s, i, k)
    w = dist.get_world_size()
    gs = [torch.empty_like(s) for _ in range(w)]
    gi = [torch.empty_like(i) for _ in range(w)]
    dist.all_gather(gs, s.contiguous())
    dist.all_gather(gi, i.contiguous())
    return sort_lists(torch.cat(gs, 1), torch.cat(gi, 1), k)


def dense_lists(X: torch.Tensor, Q: torch.Tensor, k: int, id_base: int, chunk: int = 1 << 20):
    """fp32 flat inner product of Q [n, d] bf16 against every row of the shard X [N, d] bf16, top-k by (score desc, id asc)."""
    n = Q.shape[0]
    best_s = torch.full((n, 0), 0.0, device=X.device)
    best_i = torch.zeros((n, 0), dtype=torch.int64, device=X.device)
    Qf = Q.float()
    for lo in range(0, X.shape[0], chunk):
        hi = min(X.shape[0], lo + chunk)
        S = Qf @ X[lo:hi].float().t()
        kk = min(k, hi - lo)
        s, j = torch.topk(S, kk, dim=1)
        # everything tied with the chunk's kk-th score must be seen for the id tie-break; scores of random unit vectors do not tie
        best_s, best_i = sort_lists(torch.cat([best_s, s], 1), torch.cat([best_i, j + lo + id_base], 1), k)
    return gather_lists(best_s, best_i, k)


def bm25_lists(index, q_indptr: torch.Tensor, q_term: torch.Tensor, rows, k: int):
    """Okapi scores of the sampled queries in fp64: one scatter-add of `impact * multiplicity` per query term over the
    shard's postings, stable descending sort (zero-score documents follow in id order, bm25_retriever.py:75)."""
    N, dev = index.n_docs, index.doc_id.device
    qi = q_indptr.tolist()
    out_s, out_i = [], []
    for r in rows:
        terms = q_term[qi[r]:qi[r + 1]].tolist()
        acc = torch.zeros(N, dtype=torch.float64, device=dev)
        for t in terms:                                   # a repeated token is scored once per occurrence
            if t < 0 or t >= index.vocab:
                continue
            a, b = int(index.indptr[t]), int(index.indptr[t + 1])
            if b > a:
                acc.index_add_(0, index.doc_id[a:b].long(), index.impact[a:b].double())
        s, j = torch.sort(acc, descending=True, stable=True)
        out_s.append(s[:k].float())
        out_i.append(j[:k] + index.id_base)
    return gather_lists(torch.stack(out_s), torch.stack(out_i), k)


def minmax_weighted_sum(lists, weights, k: int):
    """The reference's `weighted_sum` fusion (hybrid_retriever.py:24-30, 432-457, 505-533) in fp64 for a few rows: per channel
    min-max over the returned list (all zeros when the list is flat), score = sum of weight x norm over the channels that
    returned the doc; sorted (score desc, id asc)."""
    n = lists[0][0].shape[0]
    out_s, out_i = [], []
    for r in range(n):
        tot = {}
        for (s, i), w in zip(lists, weights):
            ids = [int(x) for x in i[r].tolist() if x >= 0]
            sc = [float(x) for x, y in zip(s[r].tolist(), i[r].tolist()) if y >= 0]
            if not ids:
                continue
            lo, hi = min(sc), max(sc)
            for d_, v in zip(ids, sc):
                nv = 0.0 if hi - lo < 1e-12 else (v - lo) / (hi - lo)
                tot[d_] = tot.get(d_, 0.0) + w * nv
        items = sorted(tot.items(), key=lambda t: (-t[1], t[0]))[:k]
        items += [(-1, float("-inf"))] * (k - len(items))
        out_i.append([a for a, _ in items])
        out_s.append([b for _, b in items])
    dev = lists[0][0].device
    return torch.tensor(out_s, dtype=torch.float64, device=dev), torch.tensor(out_i, dtype=torch.int64, device=dev)


def maxsim_candidates(tokens: torch.Tensor, Qtok: torch.Tensor, cand: torch.Tensor, tok_row_base: int, tok_rows_total: int, k: int):
    """sum_i max_j <q_i, d_j> in fp32 for the candidates whose token rows this rank owns, max-reduced over ranks."""
    n, C = cand.shape
    rows = torch.where(cand >= 0, cand % max(1, tok_rows_total), cand) - tok_row_base
    own = (cand >= 0) & (rows >= 0) & (rows < tokens.shape[0])
    sc = torch.full((n, C), float("-inf"), device=tokens.device)
    for r in range(n):
        sel = own[r].nonzero().flatten()
        if sel.numel():
            D = tokens[rows[r, sel]].float()                                   # [c, Ld, dim]
            sim = torch.einsum("ld,ctd->clt", Qtok[r].float(), D)
            sc[r, sel] = sim.amax(2).sum(1)
    dist = _dist()
    if dist is not None:
        dist.all_reduce(sc, op=dist.ReduceOp.MAX)
    return sort_lists(sc, torch.where(sc > float("-inf"), cand, torch.full_like(cand, -1)), k)


def compare(got_s: torch.Tensor, got_i: torch.Tensor, ref_s: torch.Tensor, ref_i: torch.Tensor, tau: float, floor: float = 1e-6):
    """-> (decided positions, mismatched ids among them, max relative score error over ids both sides returned)."""
    gs, gi = got_s.double().cpu(), got_i.cpu()
    rs, ri = ref_s.double().cpu(), ref_i.cpu()
    n, k = gi.shape                      # the checker's rows may be longer than k: the k-th hit then has a known successor
    decided = mism = 0
    max_rel = 0.0
    for r in range(n):
        ref = {int(a): float(b) for a, b in zip(ri[r].tolist(), rs[r].tolist()) if a >= 0}
        for a, b in zip(gi[r].tolist(), gs[r].tolist()):
            if a >= 0 and a in ref:
                max_rel = max(max_rel, abs(b - ref[a]) / max(abs(ref[a]), floor))
        row_s, row_i = rs[r].tolist(), ri[r].tolist()
        for c in range(min(k, len(row_i))):                 # positions of the engine's list
            if row_i[c] < 0:
                continue
            gap = tau * max(abs(row_s[c]), floor)
            sep_prev = c == 0 or row_s[c - 1] - row_s[c] > gap
            sep_next = c + 1 >= len(row_s) or row_i[c + 1] < 0 or row_s[c] - row_s[c + 1] > gap
            if sep_prev and sep_next:
                decided += 1
                mism += int(gi[r, c]) != row_i[c]
    return decided, mism, max_rel
