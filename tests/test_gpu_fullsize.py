"""Parity at BASELINE.json's full sizes (configs[1] and configs[3]) and the long-query BM25 paths.

The literal oracle cannot finish at these sizes, so the checks are (a) a plain PyTorch fp32 restatement of the
same arithmetic, evaluated on the device for a subset of the batch (the floating-point kernels' reference),
and (b) size-independent properties: sorted and duplicate-free rows, identical results on a second run,
a sharded run merged == the unsharded run."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import bm25 as obm25
from tests.parity import check_topk_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from legal_rag_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return engine


def _rows_sorted_and_unique(s, i):
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert (np.diff(s, axis=1) <= 0).all(), "scores not descending"
    ties = np.diff(s, axis=1) == 0
    assert (np.diff(i, axis=1)[ties] > 0).all(), "ties not broken by ascending id"
    assert all(len(set(r.tolist())) == len(r) for r in i[:64]), "duplicate ids in a row"


@pytest.mark.timeout(900)
def test_dense_config1_full_size(eng):
    """configs[1]: 10M x 1024 bf16, 4096 queries, top-100."""
    from legal_rag_b200 import synth
    N, d, nq, k = 10_000_000, 1024, 4096, 100
    X = synth.unit_rows_bf16(N, d, 2, "cuda")
    Q = synth.unit_rows_bf16(nq, d, 3, "cuda", chunk=nq)
    s, i = eng.dense_topk(X, Q, k)
    _rows_sorted_and_unique(s, i)
    s2, i2 = eng.dense_topk(X, Q, k)
    assert torch.equal(i, i2) and torch.equal(s, s2), "second run differs"
    # fp32 restatement on the same bf16 inputs for 32 queries spread over the batch, 1M rows at a time
    sub = torch.arange(0, nq, nq // 32, device="cuda")[:32]
    Qs = Q[sub].float()
    best_s = torch.full((32, 0), 0.0, device="cuda"); best_i = torch.zeros((32, 0), dtype=torch.int64, device="cuda")
    for lo in range(0, N, 1_000_000):
        S = Qs @ X[lo:lo + 1_000_000].float().T
        cs, ci = torch.topk(S, k + 20, dim=1)
        best_s = torch.cat([best_s, cs], 1); best_i = torch.cat([best_i, ci + lo], 1)
        o = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, :k + 20]
        best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
    check_topk_parity(s[sub].cpu().numpy(), i[sub].cpu().numpy(), best_s.cpu().numpy(), best_i.cpu().numpy(), k, 1e-4,
                      what="dense-10M", floor=0.1)
    # two shards merged == the unsharded scan (the multi-GPU path's invariant), on a slice of the batch
    h = N // 2
    a = eng.dense_topk(X[:h], Q[:256], k, id_base=0)
    b = eng.dense_topk(X[h:], Q[:256], k, id_base=h)
    ms, mi = eng.topk_merge(torch.cat([a[0], b[0]], 1), torch.cat([a[1], b[1]], 1), k)
    assert torch.equal(mi, i[:256]) and torch.equal(ms, s[:256])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("ragged", [False, True])
def test_maxsim_config2_full_size(eng, ragged):
    """configs[2]: 1M docs x 128 tokens x 128-d token store (32.8 GB), 1024 queries x 32 tokens, 1000 candidates each -> top-100
    (colbert_retriever.py:152, 174-182); `ragged` draws document lengths from U{32..128} to exercise the masking.  The
    literal oracle cannot hold the store; the check is a plain torch fp32 restatement of the same arithmetic on the device
    for 32 queries spread over the batch, plus sortedness and a repeat run."""
    from legal_rag_b200 import synth
    Nd, Ld, Lq, nq, C, k = 1_000_000, 128, 32, 1024, 1000, 100
    D = synth.unit_tokens_bf16(Nd, Ld, 128, 5, "cuda")
    Q = synth.unit_tokens_bf16(nq, Lq, 128, 7, "cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    cand = torch.argsort(torch.rand((nq, 4 * C), generator=g, device="cuda"), dim=1)[:, :C]            # distinct per row ...
    cand = (cand * (Nd // (4 * C)) + torch.randint(0, Nd // (4 * C), (nq, C), generator=g, device="cuda")).long()   # ... over the store
    cand[5, 17] = -1                                                                                     # a skipped slot
    doclen = None
    if ragged:
        g6 = torch.Generator(device="cuda"); g6.manual_seed(6)
        doclen = torch.randint(32, Ld + 1, (Nd,), generator=g6, device="cuda", dtype=torch.int32)
    s, i = eng.maxsim_rerank(D, doclen, Q, cand, k)
    _rows_sorted_and_unique(s, i)
    s2, i2 = eng.maxsim_rerank(D, doclen, Q, cand, k)
    assert torch.equal(i, i2) and torch.equal(s, s2), "second run differs"
    sub = list(range(0, nq, nq // 32))[:32]
    ref_s = torch.full((len(sub), C), float("-inf"), device="cuda")
    for n, q in enumerate(sub):
        rows = cand[q]
        ok = rows >= 0
        sim = torch.einsum("ld,ctd->clt", Q[q].float(), D[rows[ok]].float())                            # [c, Lq, Ld]
        if doclen is not None:
            sim = sim.masked_fill(torch.arange(Ld, device="cuda")[None, None, :] >= doclen[rows[ok]].long()[:, None, None], -9999.0)
        ref_s[n, ok] = sim.amax(2).sum(1)
    o = torch.argsort(ref_s, dim=1, descending=True, stable=True)[:, :k + 20]
    check_topk_parity(s[sub].cpu().numpy(), i[sub].cpu().numpy(), torch.gather(ref_s, 1, o).cpu().numpy(),
                      torch.gather(cand[sub], 1, o).cpu().numpy(), k, 1e-4, what=f"maxsim-1M-ragged{int(ragged)}", floor=1.0)


def _torch_bm25_reference(index, q_indptr, q_term, rows, k):
    """fp64 scatter-add of `multiplicity * impact` over the posting lists of each query, then top-k by
    (score desc, id asc) -- the reference's get_scores + stable sort, evaluated on the device."""
    N = index.n_docs
    out_s, out_i = [], []
    qi, qt = q_indptr.cpu().numpy(), q_term.cpu().numpy()
    for q in rows:
        sc = torch.zeros(N, dtype=torch.float64, device="cuda")
        for t in qt[qi[q]:qi[q + 1]]:
            if t < 0 or t >= index.vocab:
                continue
            lo, hi = int(index.indptr[t]), int(index.indptr[t + 1])
            sc.index_add_(0, index.doc_id[lo:hi].long(), index.impact[lo:hi].double())
        cs, ci = torch.topk(sc, k, sorted=True)
        key = torch.argsort(ci, stable=True)                      # stable two-pass sort: id asc, then score desc
        cs, ci = cs[key], ci[key]
        key = torch.argsort(cs, descending=True, stable=True)
        out_s.append(cs[key].cpu().numpy()); out_i.append((ci[key] + index.id_base).cpu().numpy())
    return np.stack(out_s), np.stack(out_i)


@pytest.mark.timeout(900)
def test_bm25_config3_full_size(eng):
    """configs[3]: 50M docs, Zipf(1) vocabulary of 500k, 8192 queries of 2-8 terms, top-100."""
    from legal_rag_b200 import synth
    N, V, nq, k = 50_000_000, 500_000, 8192, 100
    index, st = synth.bm25_synthetic_index(N, V, 10, "cuda")
    q_indptr, q_term, mx = synth.bm25_synthetic_queries(nq, V, 11, "cuda")
    s, i = eng.bm25_topk(index, q_indptr, q_term, mx, k)
    _rows_sorted_and_unique(s, i)
    s2, i2 = eng.bm25_topk(index, q_indptr, q_term, mx, k)
    assert torch.equal(i, i2) and torch.equal(s, s2), "second run differs (fixed-point sums must be order-independent)"
    rows = list(range(0, nq, nq // 24))[:24]
    O_s, O_i = _torch_bm25_reference(index, q_indptr, q_term, rows, k + 20)
    check_topk_parity(s[rows].cpu().numpy(), i[rows].cpu().numpy(), O_s, O_i, k, 1e-3, what="bm25-50M")
    # the same queries as a small batch take the split-chain path (many chains per query): same answer
    sub = torch.tensor(rows[:8], device="cuda")
    lens = (q_indptr[1:] - q_indptr[:-1])[sub]
    qi8 = torch.zeros(9, dtype=torch.int64, device="cuda"); qi8[1:] = torch.cumsum(lens, 0)
    qt8 = torch.cat([q_term[int(q_indptr[r]):int(q_indptr[r + 1])] for r in rows[:8]])
    s8, i8 = eng.bm25_topk(index, qi8, qt8, mx, k)
    assert torch.equal(i8, i[sub]) and torch.equal(s8, s[sub])


def test_bm25_long_queries_many_terms(eng):
    """Queries of 33-128 tokens: more than one lane batch of terms in the copy warps, bounds groups smaller than an item,
    repeated tokens, and k at the API maximum."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(123)
    N, V = 150_000, 3000
    lens = np.clip(np.round(rng.lognormal(np.log(10), 0.5, N)), 2, 40).astype(np.int64)
    p = 1.0 / np.arange(1, V + 1); p /= p.sum()
    flat = rng.choice(V, size=int(lens.sum()), p=p)
    off = np.concatenate([[0], np.cumsum(lens)])
    docs = [flat[off[j]:off[j + 1]] for j in range(N)]
    host = Bm25HostIndex.from_token_ids(docs, V)
    csr = obm25.CsrBM25.from_token_ids(docs, V)
    queries = [rng.choice(V, size=n, p=p).tolist() for n in (33, 40, 64, 100, 128, 128)]
    queries[4] = (queries[4][:30] * 5)[:128]                   # heavy repetition: multiplicities up to 5
    queries.append(rng.choice(V, size=128, replace=False).tolist())   # 128 distinct terms
    dev = host.to_device("cuda")
    qi, qt, mx = host.encode_queries(queries)
    for k in (100, 1024):
        s, i = eng.bm25_topk(dev, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, k)
        m = k + 30
        O = [csr.search(q, m) for q in queries]
        check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), np.stack([o[0] for o in O]), np.stack([o[1] for o in O]), k, 1e-3,
                          what=f"bm25-long-k{k}")
    with pytest.raises(Exception, match="at most 128"):
        eng.bm25_topk(dev, torch.tensor([0, 129]).cuda(), torch.zeros(129, dtype=torch.int32).cuda(), 129, 10)


def test_bm25_tiny_corpus_and_k_larger_than_n(eng):
    from legal_rag_b200.bm25_index import Bm25HostIndex
    corpus = [["alpha", "beta"], ["beta", "gamma", "gamma"], ["delta"]]
    host = Bm25HostIndex.from_tokens(corpus)
    lit = obm25.BM25Okapi(corpus)
    dev = host.to_device("cuda")
    queries = [["gamma"], ["beta", "delta"], ["nope"], []]
    qi, qt, mx = host.encode_queries(queries)
    s, i = eng.bm25_topk(dev, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), max(mx, 1), 7)
    for q, toks in enumerate(queries):
        os_, oi = obm25.search(lit, toks, 3)
        got_i, got_s = i[q].cpu().numpy(), s[q].cpu().numpy()
        assert (got_i[3:] == -1).all()
        if dev.nonneg:
            assert got_i[:3].tolist() == oi.tolist(), (toks, got_i, oi)
            np.testing.assert_allclose(got_s[:3], os_, rtol=1e-3, atol=1e-6)


def test_bm25_negative_idf_corpus_spanning_slabs_and_items(eng):
    """Every term occurs in most documents, so rank_bm25's epsilon-floored idf is negative for all of them: matched
    documents score below untouched ones, no slab may be skipped and zero-score documents are ranked, across several
    slabs, work items and (with few queries) split chains."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(9)
    N, V = 40_000, 8
    docs = [np.nonzero(rng.random(V) < 0.8)[0] for _ in range(N)]
    docs = [d if d.size else np.array([0]) for d in docs]
    for j in range(0, N, 97):
        docs[j] = np.array([V - 1])                 # documents that match few queries keep score 0 often
    host = Bm25HostIndex.from_token_ids(docs, V)
    csr = obm25.CsrBM25.from_token_ids(docs, V)
    dev = host.to_device("cuda")
    assert not dev.nonneg
    queries = [[0], [1, 2], [3, 3, 4], [5], [0, 1, 2, 3, 4, 5, 6], [V + 3]]
    qi, qt, mx = host.encode_queries(queries)
    for item_slabs, reps in ((0, 1), (1, 120)):
        eng.bm25_set_item_slabs(item_slabs)
        try:
            s, i = eng.bm25_topk(dev, torch.from_numpy(np.concatenate([qi[:-1] + r * qi[-1] for r in range(reps)] + [[reps * qi[-1]]])).cuda(),
                                 torch.from_numpy(np.tile(qt, reps)).cuda(), mx, 50)
        finally:
            eng.bm25_set_item_slabs(0)
        s, i = s[:len(queries)].cpu().numpy(), i[:len(queries)].cpu().numpy()
        O = [csr.search(q, 80) for q in queries]
        check_topk_parity(s, i, np.stack([o[0] for o in O]), np.stack([o[1] for o in O]), 50, 1e-3, what=f"bm25-negidf-items{item_slabs}")
