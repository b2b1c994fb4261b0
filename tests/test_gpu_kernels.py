"""GPU parity tests: every liblrag kernel, called through the C ABI (via legal_rag_b200.engine),
against the CPU oracle on the same seeded inputs.  Tolerances are the north star's: 1e-2 relative for
bf16 dense / MaxSim, 1e-3 for fp32 BM25 and fusion; ids exact wherever the oracle's scores decide.
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import fuse as ofuse
from oracle import maxsim as omaxsim
from tests.parity import check_topk_parity

pytestmark = pytest.mark.gpu

TAU_BF16 = 1e-2
TAU_FP32 = 1e-3


@pytest.fixture(scope="module")
def eng():
    from legal_rag_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return engine


def _unit(rng, shape):
    x = rng.standard_normal(shape).astype(np.float32)
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _bf16(x):
    """(device bf16 tensor, the same values back as fp32 numpy)."""
    t = torch.from_numpy(x).cuda().to(torch.bfloat16)
    return t, t.float().cpu().numpy()


# ---------------------------------------------------------------- select / merge
def test_topk_select_matches_oracle(eng):
    rng = np.random.default_rng(0)
    # rows longer than 2^18 scores take the sliced single-pass path (slices + merge)
    for nq, N, k in [(3, 1000, 100), (5, 37, 64), (2, 70000, 1024), (4, 100, 100), (1, 1, 1), (3, 700_001, 100), (2, 1_100_000, 1024),
                      # balanced runs of trips: runs that cross row ends, rows of a few trips, a 4-score last trip, whole-trip rows
                      (300, 20_000, 100), (700, 8196, 10), (40, 131_072, 256), (9, 1_000_004, 1024), (1100, 8192, 3)]:
        S = rng.standard_normal((nq, N)).astype(np.float32)
        S[:, ::7] = S[:, :1]                      # plenty of exact ties -> lower id first
        s, i = eng.topk_select(torch.from_numpy(S).cuda(), k, id_base=1000)
        D, I = odense.topk_rows(S, k, id_base=1000)
        np.testing.assert_array_equal(i.cpu().numpy(), I)
        np.testing.assert_array_equal(s.cpu().numpy(), D)


def test_topk_select_one_warp_per_row(eng):
    """Many medium rows take the warp-per-row kernel (aligned rows, no column ids, k <= 256, 2048 <= N < 131072, >= 64 rows):
    ragged row ends, a row stride longer than the row, runs of -inf longer than a trip, rows with fewer finite scores than
    k, ascending rows (every trip raises the threshold), exact ties, k = 1 and k = 256."""
    rng = np.random.default_rng(21)
    W = rng.standard_normal((200, 140_000)).astype(np.float32)
    W[:, ::5] = W[:, :1]
    W[1, 1000:60_000] = -np.inf
    W[2, :] = np.sort(W[2, :])
    W[3, 30:] = -np.inf
    W[4, :] = 0.5
    Wd = torch.from_numpy(W).cuda()
    for c0, N, k in [(0, 20_000, 100), (0, 2048, 256), (4, 131_071, 100), (8, 131_070, 1), (0, 5_001, 37), (12, 65_537, 256), (0, 70_000, 200)]:
        s, i = eng.topk_select(Wd[:, c0:c0 + N], k, id_base=7)
        D, I = odense.topk_rows(W[:, c0:c0 + N], k, id_base=7)
        np.testing.assert_array_equal(i.cpu().numpy(), I, err_msg=f"c0={c0} N={N} k={k}")
        np.testing.assert_array_equal(s.cpu().numpy(), D, err_msg=f"c0={c0} N={N} k={k}")


def test_topk_select_column_window_of_a_wider_matrix(eng):
    """Row stride != row length: 16-byte aligned rows whose length is not a multiple of four scores (the ragged end of the
    streamed path), a window starting at an unaligned column (plain-load path), short last slices, runs of -inf."""
    rng = np.random.default_rng(3)
    W = rng.standard_normal((3, 800_000)).astype(np.float32)
    W[:, ::5] = W[:, :1]
    W[1, 1000:300_000] = -np.inf
    W[2, :2048] = np.sort(W[2, :2048])        # first trip ascending: the bootstrap threshold sits at its very end
    W[0, 50:] = -np.inf                       # fewer finite scores than k: the k-th best is -inf, lowest ids first
    Wd = torch.from_numpy(W).cuda()
    for c0, N, k in [(0, 799_997, 100), (0, 262_146, 7), (4, 524_289 + 2048, 256), (3, 700_000, 100), (0, 800_000, 1000), (8, 270_001, 1),
                     (0, 300_000, 257)]:
        s, i = eng.topk_select(Wd[:, c0:c0 + N], k, id_base=5)
        D, I = odense.topk_rows(W[:, c0:c0 + N], k, id_base=5)
        np.testing.assert_array_equal(i.cpu().numpy(), I, err_msg=f"c0={c0} N={N} k={k}")
        np.testing.assert_array_equal(s.cpu().numpy(), D, err_msg=f"c0={c0} N={N} k={k}")


def test_topk_select_with_col_ids_and_skips(eng):
    rng = np.random.default_rng(1)
    for nq, C, k, pool in [(4, 300, 50, 10000), (2, 600_000, 100, 5_000_000), (300, 20_000, 64, 100_000)]:      # the last two are streamed
        S = rng.standard_normal((nq, C)).astype(np.float32)
        ids = np.stack([rng.permutation(pool)[:C] for _ in range(nq)]).astype(np.int64)
        ids[:, 5::11] = -1
        S[0, :40] = 9.25                          # ties resolved by id, not by column
        s, i = eng.topk_select(torch.from_numpy(S).cuda(), k, col_id=torch.from_numpy(ids).cuda())
        for q in range(nq):
            ok = np.nonzero(ids[q] >= 0)[0]
            order = ok[np.lexsort((ids[q, ok], -S[q, ok].astype(np.float64)))][:k]
            np.testing.assert_array_equal(i[q].cpu().numpy(), ids[q, order])
            np.testing.assert_array_equal(s[q].cpu().numpy(), S[q, order])


def test_topk_merge_matches_oracle(eng):
    rng = np.random.default_rng(2)
    nq, G, k = 7, 8, 100
    sc = rng.standard_normal((nq, G * k)).astype(np.float32)
    ids = np.stack([rng.permutation(10 ** 6)[:G * k] for _ in range(nq)]).astype(np.int64)
    ids[:, -37:] = -1
    sc[:, 3] = sc[:, 4]
    s, i = eng.topk_merge(torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda(), k)
    D, I = odense.merge_topk(sc, ids, k)
    np.testing.assert_array_equal(i.cpu().numpy(), I)
    np.testing.assert_array_equal(s.cpu().numpy(), D)
    # fewer valid entries than k -> padding
    s, i = eng.topk_merge(torch.from_numpy(sc[:, :10]).cuda(), torch.from_numpy(ids[:, :10]).cuda(), 16)
    assert (i[:, 10:] == -1).all() and (i[:, :10] >= 0).all()


def test_merge_and_select_keep_64_bit_ids(eng):
    """Ids beyond 2^32 (and rows whose ids span more than 2^32) come back whole from the merge, the in-place shard merge
    and the column-id select: they never pass through the 32-bit tie field of the selection keys (round-1 finding)."""
    rng = np.random.default_rng(12)
    nq, G, k = 5, 4, 50
    sc = rng.standard_normal((nq, G * k)).astype(np.float32)
    big = (1 << 33) + 5
    # narrow rows far above 2^32, and wide rows that mix small and huge ids
    ids_narrow = big + np.stack([rng.permutation(10 ** 6)[:G * k] for _ in range(nq)]).astype(np.int64)
    ids_wide = ids_narrow.copy()
    ids_wide[:, ::3] -= big
    for ids in (ids_narrow, ids_wide):
        ids = ids.copy()
        ids[:, -9:] = -1
        sc2 = sc.copy()
        sc2[:, 3] = sc2[:, 4] = sc2[:, 5]                    # a three-way tie: ascending id order, whatever the width
        s, i = eng.topk_merge(torch.from_numpy(sc2).cuda(), torch.from_numpy(ids).cuda(), k)
        D, I = odense.merge_topk(sc2, ids, k)
        np.testing.assert_array_equal(i.cpu().numpy(), I)
        np.testing.assert_array_equal(s.cpu().numpy(), D)
        # the same lists as G shards of k entries, laid out [shard, query, k] like an all-gather leaves them
        gs = torch.from_numpy(np.ascontiguousarray(sc2.reshape(nq, G, k).transpose(1, 0, 2))).cuda()
        gi = torch.from_numpy(np.ascontiguousarray(ids.reshape(nq, G, k).transpose(1, 0, 2))).cuda()
        s, i = eng.topk_merge_shards(gs, gi, k)
        np.testing.assert_array_equal(i.cpu().numpy(), I)
        np.testing.assert_array_equal(s.cpu().numpy(), D)
    # column ids of the select (MaxSim candidate ranking): short rows and streamed long rows
    for N in (300, 200_000):
        S = rng.standard_normal((3, N)).astype(np.float32)
        cid = big * 3 + np.stack([rng.permutation(N) for _ in range(3)]).astype(np.int64) * 7
        cid[:, 5] = -1
        s, i = eng.topk_select(torch.from_numpy(S).cuda(), 40, col_id=torch.from_numpy(cid).cuda())
        Sm = np.where(cid >= 0, S, -np.inf)
        order = np.argsort(-Sm, axis=1, kind="stable")[:, :40]
        np.testing.assert_array_equal(i.cpu().numpy(), np.take_along_axis(cid, order, 1))
        np.testing.assert_array_equal(s.cpu().numpy(), np.take_along_axis(S, order, 1))


# ---------------------------------------------------------------- dense
DENSE_SHAPES = [
    # (N, d, nq, k)
    (591, 768, 1, 100),        # config 1, the reference's one-query-per-call pattern
    (591, 768, 200, 100),      # config 1 batched, 2 query blocks
    (5000, 1024, 19, 10),
    (5000, 1024, 20, 100),     # faiss switches strategy at nq = 20
    (40000, 1024, 300, 100),   # several doc tiles per CTA, 3 query blocks, ragged last tile
    (100, 64, 5, 200),         # k > N: padding
    (300, 136, 130, 7),        # d not a multiple of 64: TMA zero-fills the K tail
]


@pytest.mark.parametrize("N,d,nq,k", DENSE_SHAPES)
def test_dense_cuda_core_reference_kernel(eng, N, d, nq, k):
    rng = np.random.default_rng(N + d + nq)
    Xd, Xr = _bf16(_unit(rng, (N, d)))
    Qd, Qr = _bf16(_unit(rng, (nq, d)))
    s, i = eng.dense_topk(Xd, Qd, k, id_base=7, reference_kernel=True)
    D, I = odense.flat_ip_topk(Qr, Xr, min(k + 20, max(N, 1)), id_base=7)
    check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), D, I, k, 1e-4, what="dense-ref")


@pytest.mark.parametrize("N,d,nq,k", DENSE_SHAPES)
def test_dense_tcgen05_matches_oracle(eng, N, d, nq, k):
    rng = np.random.default_rng(N + d + nq)
    X32, Q32 = _unit(rng, (N, d)), _unit(rng, (nq, d))
    Xd, Xr = _bf16(X32)
    Qd, Qr = _bf16(Q32)
    s, i = eng.dense_topk(Xd, Qd, k, id_base=7)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    m = min(k + 20, N)
    # oracle A: fp32 math on the bf16-rounded inputs (isolates kernel error)
    D, I = odense.flat_ip_topk(Qr, Xr, m, id_base=7)
    check_topk_parity(s, i, D, I, k, 1e-4, what="dense-A")
    # oracle B: the reference's numerics (fp32 inputs), north-star tolerance
    D, I = odense.flat_ip_topk(Q32, X32, m, id_base=7)
    check_topk_parity(s, i, D, I, k, TAU_BF16, what="dense-B", floor=0.1)


def test_dense_duplicate_vectors_tie_to_lower_id(eng):
    rng = np.random.default_rng(5)
    X = _unit(rng, (2000, 256))
    X[1500] = X[3]; X[77] = X[3]
    Xd, Xr = _bf16(X)
    Qd, Qr = _bf16(X[[3, 10]])
    s, i = eng.dense_topk(Xd, Qd, 10)
    assert i[0, :3].tolist() == [3, 77, 1500]
    assert s[0, 0] == s[0, 1] == s[0, 2]


def test_dense_rejects_bad_arguments(eng):
    X = torch.zeros((10, 60), dtype=torch.bfloat16, device="cuda")
    Q = torch.zeros((1, 60), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 8"):
        eng.dense_topk(X, Q, 5)
    X = torch.zeros((10, 64), dtype=torch.bfloat16, device="cuda")
    Q = torch.zeros((1, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        eng.dense_topk(X, Q, 5000)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        eng.dense_topk(X.cpu(), Q, 5)


# ---------------------------------------------------------------- BM25
def _ucc(golden_dir):
    z = np.load(os.path.join(golden_dir, "ucc_corpus.npz"))
    lens, flat = z["doc_len"], z["tokens"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    return [flat[off[i]:off[i + 1]] for i in range(len(lens))], len(z["vocab"])


def _window_queries(rng, docs, n):
    out = []
    for _ in range(n):
        d = rng.integers(0, len(docs))
        L = int(rng.integers(3, 9))
        st = int(rng.integers(0, max(1, len(docs[d]) - L)))
        out.append(docs[d][st:st + L].tolist())
    return out


def _run_bm25(eng, host, queries, k, lo=0, hi=None):
    dev = host.to_device("cuda", lo, hi)
    qi, qt, mx = host.encode_queries(queries)
    return eng.bm25_topk(dev, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, k)


def test_bm25_ucc_matches_literal_oracle(eng, golden_dir):
    from legal_rag_b200.bm25_index import Bm25HostIndex
    docs, V = _ucc(golden_dir)
    host = Bm25HostIndex.from_token_ids(docs, V)
    lit = obm25.BM25Okapi([[str(t) for t in d] for d in docs])
    rng = np.random.default_rng(44)
    queries = _window_queries(rng, docs, 64)
    queries[0] = queries[0] + queries[0][:2]          # repeated tokens count once per occurrence
    queries[1] = [V + 5, -1]                          # only OOV tokens: every doc scores 0 -> ids 0..k-1
    queries[2] = []                                   # empty query
    queries[3] = [int(docs[10][0])]                   # single rare-ish term: fewer than k matches -> zero-score fill
    k = 100
    s, i = _run_bm25(eng, host, queries, k)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    O_s = np.zeros((len(queries), 591)); O_i = np.zeros((len(queries), 591), dtype=np.int64)
    for q, toks in enumerate(queries):
        sc = lit.get_scores([str(t) for t in toks if 0 <= t < V] + ["<oov>"] * sum(1 for t in toks if not 0 <= t < V))
        O_s[q], O_i[q] = obm25.topk_reference(sc, 591)
    check_topk_parity(s, i, O_s, O_i, k, TAU_FP32, what="bm25-ucc")
    assert i[1].tolist() == list(range(k)) and (s[1] == 0).all()
    assert i[2].tolist() == list(range(k)) and (s[2] == 0).all()


def test_bm25_synthetic_multi_slab_and_split(eng):
    """Corpus spanning several 16K-doc slabs and batches; small nq forces the doc-range split path."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(10)
    N, V = 600_000, 5000
    lens = np.clip(np.round(rng.lognormal(np.log(12), 0.6, N)), 2, 64).astype(np.int64)
    p = 1.0 / np.arange(1, V + 1); p /= p.sum()
    flat = rng.choice(V, size=int(lens.sum()), p=p)
    off = np.concatenate([[0], np.cumsum(lens)])
    docs = [flat[off[i]:off[i + 1]] for i in range(N)]
    host = Bm25HostIndex.from_token_ids(docs, V)
    csr = obm25.CsrBM25.from_token_ids(docs, V)
    np.testing.assert_allclose(host.idf, csr.idf, rtol=1e-12)
    queries = [rng.choice(V, size=int(rng.integers(2, 9)), p=p).tolist() for _ in range(12)]
    queries.append([V - 1, V - 2])                    # rare terms only: sparse slabs are skipped
    queries.append([int(np.argmin(np.diff(host.indptr) + (np.diff(host.indptr) == 0) * 10 ** 9))])  # rarest term
    # (k, item slabs, query repetitions): 14 queries over 600k docs split every query into chains of one
    # item; 1-slab items with 700 queries make every chain hand its state over 37 times
    for k, item_slabs, reps in ((100, 0, 1), (1000, 0, 1), (100, 1, 50), (100, 3, 50)):
        eng.bm25_set_item_slabs(item_slabs)
        try:
            s, i = _run_bm25(eng, host, queries * reps, k)
        finally:
            eng.bm25_set_item_slabs(0)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        if reps > 1:      # identical queries give identical rows whichever CTA ran their items
            for r in range(1, reps):
                sl = slice(r * len(queries), (r + 1) * len(queries))
                np.testing.assert_array_equal(i[sl], i[:len(queries)])
                np.testing.assert_array_equal(s[sl], s[:len(queries)])
            s, i = s[:len(queries)], i[:len(queries)]
        m = k + 50
        O_s = np.zeros((len(queries), m)); O_i = np.zeros((len(queries), m), dtype=np.int64)
        for q, toks in enumerate(queries):
            O_s[q], O_i[q] = csr.search(toks, m)
        check_topk_parity(s, i, O_s, O_i, k, TAU_FP32, what=f"bm25-synth-k{k}-items{item_slabs}")
    # a shard [lo, hi) with global statistics returns global ids
    lo, hi = 200_000, 450_000
    s, i = _run_bm25(eng, host, queries[:4], 100, lo, hi)
    for q, toks in enumerate(queries[:4]):
        sc = csr.get_scores(toks)[lo:hi]
        order = np.lexsort((np.arange(hi - lo), -sc))[:150]
        check_topk_parity(s[q:q + 1].cpu().numpy(), i[q:q + 1].cpu().numpy(), sc[order][None], (order + lo)[None], 100,
                          TAU_FP32, what="bm25-shard")


def test_bm25_dense_rows_change_nothing_but_the_speed(eng):
    """Terms that occur in most documents are walked through a dense impact row instead of their posting list
    (lrag_bm25_topk_dense): same fixed-point sums, so bit-identical scores and ids -- at several slab / item / split shapes,
    with a range end inside a chunk, and with a repeated dense term in a query."""
    from legal_rag_b200 import synth
    for N, V, nq, k, mean_len in ((70_001, 300, 300, 100, 40.0), (5_000, 50, 7, 20, 30.0), (300_000, 2_000, 40, 50, 24.0)):
        index, st = synth.bm25_synthetic_index(N, V, 77, "cuda", mean_len=mean_len)
        q_indptr, q_term, mx = synth.bm25_synthetic_queries(nq, V, 78, "cuda")
        q_term[:2] = 0                                           # the head term twice in the first query
        for item_slabs in (0, 1):
            eng.bm25_set_item_slabs(item_slabs)
            try:
                index._dense_built, index.dense_term, index.dense_rows = True, None, None      # posting lists only
                s0, i0 = eng.bm25_topk(index, q_indptr, q_term, mx, k)
                n_rows = index.build_dense_rows()
                assert n_rows >= 1 and index.dense_rows.shape[1] % 32 == 0
                s1, i1 = eng.bm25_topk(index, q_indptr, q_term, mx, k)
            finally:
                eng.bm25_set_item_slabs(0)
            assert torch.equal(i0, i1) and torch.equal(s0, s1), f"dense rows changed the result (N={N}, item_slabs={item_slabs})"
    # impacts that may be negative: every slab is ranked in full, zero-score docs included; rows and posting lists still agree
    index, st = synth.bm25_synthetic_index(30_000, 200, 79, "cuda", mean_len=30.0)
    index.impact[::7] *= -1.0
    index.nonneg = False
    q_indptr, q_term, mx = synth.bm25_synthetic_queries(25, 200, 80, "cuda")
    index._dense_built, index.dense_term, index.dense_rows = True, None, None
    s0, i0 = eng.bm25_topk(index, q_indptr, q_term, mx, 64)
    assert index.build_dense_rows(min_density=0.3) >= 1
    s1, i1 = eng.bm25_topk(index, q_indptr, q_term, mx, 64)
    assert torch.equal(i0, i1) and torch.equal(s0, s1), "dense rows changed the result (negative impacts)"


def test_bm25_negative_average_idf_ranks_every_document(eng):
    """3-doc corpus whose epsilon-floored idf is negative (tests/test_oracle.py's hand-computed case):
    zero-score documents then outrank matched ones and no slab may be skipped."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    corpus = [["a", "b", "b", "c"], ["a", "c"], ["a", "d", "d", "d"]]
    host = Bm25HostIndex.from_tokens(corpus)
    lit = obm25.BM25Okapi(corpus)
    dev = host.to_device("cuda")
    assert not dev.nonneg
    queries = [["a"], ["b", "b", "zzz", "a"], ["c", "a"], ["zzz"]]
    qi, qt, mx = host.encode_queries(queries)
    s, i = eng.bm25_topk(dev, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, 5)
    for q, toks in enumerate(queries):
        os_, oi = obm25.search(lit, toks, 3)
        assert i[q, :3].tolist() == oi.tolist(), (toks, i[q], oi)
        np.testing.assert_allclose(s[q, :3].cpu().numpy(), os_, rtol=TAU_FP32, atol=1e-6)
        assert (i[q, 3:] == -1).all()


# ---------------------------------------------------------------- MaxSim
@pytest.mark.parametrize("Nd,Ld,Lq,nq,C,k", [(300, 128, 32, 5, 200, 100), (64, 32, 7, 3, 64, 10),
                                              (50, 224, 32, 2, 50, 20), (2000, 128, 32, 40, 1000, 100)])
def test_maxsim_matches_oracle(eng, Nd, Ld, Lq, nq, C, k):
    rng = np.random.default_rng(Nd + Ld)
    D32, Q32 = _unit(rng, (Nd, Ld, 128)), _unit(rng, (nq, Lq, 128))
    Dd, Dr = _bf16(D32)
    Qd, Qr = _bf16(Q32)
    doclen = rng.integers(1, Ld + 1, Nd).astype(np.int32)
    doclen[:3] = Ld
    cand = np.stack([rng.permutation(Nd)[:C] if C <= Nd else rng.integers(0, Nd, C) for _ in range(nq)]).astype(np.int64)
    cand[:, 3::17] = -1
    if C >= 64:                      # skipped slots never enter the pipeline: a run of them, a whole 32-slot item, a whole query
        cand[0, 5:31] = -1
        cand[1 % nq, 32:64] = -1
        cand[nq - 1, :] = -1
        cand[0, 40] = Nd + 5         # past the store = skipped
    sc = eng.maxsim_scores(Dd, torch.from_numpy(doclen).cuda(), Qd, torch.from_numpy(cand).cuda()).cpu().numpy()
    cand_o = np.where(cand >= Nd, -1, cand)          # the oracle knows only -1 as "skip"
    ref = omaxsim.maxsim_scores(Qr, Dr, doclen, cand_o)
    assert np.isneginf(sc[(cand < 0) | (cand >= Nd)]).all()
    ok = (cand >= 0) & (cand < Nd)
    np.testing.assert_allclose(sc[ok], ref[ok], rtol=2e-4, atol=2e-4)          # oracle A: same bf16 inputs
    ref32 = omaxsim.maxsim_scores(Q32, D32, doclen, cand_o)
    # oracle B: fp32 inputs.  A sum of 32 signed terms can sit near zero while each term carries bf16
    # input-rounding error, hence the absolute floor
    np.testing.assert_allclose(sc[ok], ref32[ok], rtol=TAU_BF16, atol=1e-2)
    s, i = eng.maxsim_rerank(Dd, torch.from_numpy(doclen).cuda(), Qd, torch.from_numpy(cand).cuda(), k)
    o_s, o_i = omaxsim.rerank_topk(Qr, Dr, doclen, cand_o, min(C, k + 20))
    check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), o_s, o_i, k, 1e-3, what="maxsim-rerank")
    # no doclen = every token counts
    sc = eng.maxsim_scores(Dd, None, Qd, torch.from_numpy(cand).cuda()).cpu().numpy()
    ref = omaxsim.maxsim_scores(Qr, Dr, None, cand_o)
    np.testing.assert_allclose(sc[ok], ref[ok], rtol=2e-4, atol=2e-4)


# ---------------------------------------------------------------- fusion
def _pad(pairs, kc):
    s = np.full(kc, -3.4028234663852886e38, dtype=np.float32)
    i = np.full(kc, -1, dtype=np.int64)
    for j, (d, v) in enumerate(pairs):
        i[j], s[j] = d, v
    return s, i


def _fuse_gpu(eng, cases, k, **kw):
    kc = max(1, max(len(c[ch]) for c in cases for ch in ("dense", "bm25", "colbert")))
    chans = {}
    for ch in ("dense", "bm25", "colbert"):
        S = np.stack([_pad(c[ch], kc)[0] for c in cases]); I = np.stack([_pad(c[ch], kc)[1] for c in cases])
        chans[ch] = (torch.from_numpy(S).cuda(), torch.from_numpy(I).cuda())
    return eng.fuse_topk(chans["dense"], chans["bm25"], chans["colbert"], k=k, breakdown=True, **kw)


def test_fuse_matches_reference_golden(eng, golden_dir):
    """tests/golden/fuse_golden.json = outputs of the reference's own HybridRetriever._fuse."""
    with open(os.path.join(golden_dir, "fuse_golden.json")) as f:
        golden = json.load(f)
    checked = 0
    for name, case in golden["cases"].items():
        inp = {ch: [tuple(p) for p in case["inputs"][ch]] for ch in ("dense", "bm25", "colbert")}
        # the kernel computes on fp32 channel scores; feed the oracle the same rounded values
        if not any(inp.values()):
            continue
        for run in case["runs"]:
            k = max(1, len(run["out"]))
            s, i, bd = _fuse_gpu(eng, [inp], k, method=run["method"], w_dense=run["w_dense"], w_bm25=run["w_bm25"],
                                 w_colbert=run["w_colbert"], rrf_k=run["rrf_k"], alpha=run["alpha"])
            s, i, bd = s[0].cpu().numpy(), i[0].cpu().numpy(), bd[0].cpu().numpy()
            ref = run["out"]
            O_s = np.array([[r["score"] for r in ref]]); O_i = np.array([[r["id"] for r in ref]])
            check_topk_parity(s[None], i[None], O_s, O_i, k, TAU_FP32, what=f"fuse-{name}-{run['method']}")
            by_id = {r["id"]: r for r in ref}
            for j in range(k):
                r = by_id[int(i[j])]
                got = dict(zip(eng.BREAKDOWN_FIELDS, bd[j]))
                for key in ("rrf_norm", "weighted_sum", "dense_norm", "bm25_norm", "colbert_norm"):
                    assert got[key] == pytest.approx(r[key], rel=TAU_FP32, abs=1e-5), (name, run["method"], key)
                for ch in ("dense", "bm25", "colbert"):
                    assert got[f"contrib_{ch}"] == pytest.approx(r["channel_contrib"].get(ch, 0.0), rel=TAU_FP32, abs=1e-5)
            checked += 1
    assert checked >= 40


def test_fuse_batch_matches_oracle_and_filters(eng):
    rng = np.random.default_rng(3)
    cases = []
    for q in range(64):
        def lst(n, lo, hi):
            ids = rng.permutation(400)[:n]
            sc = np.sort(rng.uniform(lo, hi, n).astype(np.float32))[::-1]
            return [(int(a), float(b)) for a, b in zip(ids, sc)]
        cases.append({"dense": lst(100, 0.2, 0.9), "bm25": lst(int(rng.integers(0, 101)), 0, 40), "colbert": lst(100, 10, 30)})
    for method in ofuse.METHODS:
        for min_final in (float("-inf"), 0.2):
            s, i, _ = _fuse_gpu(eng, cases, 100, method=method, min_final=min_final)
            s, i = s.cpu().numpy(), i.cpu().numpy()
            for q, c in enumerate(cases):
                rows = [r for r in ofuse.fuse(c["dense"], c["bm25"], c["colbert"], method=method) if r["score"] >= min_final]
                m = min(len(rows), 130)
                O_s = np.array([[r["score"] for r in rows[:m]]]); O_i = np.array([[r["id"] for r in rows[:m]]])
                if m == 0:
                    assert (i[q] == -1).all()
                    continue
                # entries within tau of the min_final cut may legitimately fall on either side
                near = [r for r in ofuse.fuse(c["dense"], c["bm25"], c["colbert"], method=method)
                        if abs(r["score"] - min_final) < 1e-6]
                if near:
                    continue
                check_topk_parity(s[q:q + 1], i[q:q + 1], O_s, O_i, 100, TAU_FP32, what=f"fuse-batch-{method}")


def test_fuse_long_lists_wide_ids_and_ties(eng):
    """kc = 1024 (the hash table at its highest load), ids above 2^32 (the sort key holds entry indices, not ids, so the
    id order of equal scores must still hold), and channels full of equal scores."""
    rng = np.random.default_rng(33)
    cases = []
    for q in range(6):
        pool = rng.permutation(5000)[:2500].astype(np.int64) + (1 << 33) * (q % 2)
        def lst(n, levels):
            ids = rng.permutation(pool)[:n]
            sc = np.sort(rng.integers(0, levels, n).astype(np.float32))[::-1]       # few distinct values: long tie runs
            return [(int(a), float(b)) for a, b in zip(ids, sc)]
        cases.append({"dense": lst(1024, 7), "bm25": lst(int(rng.integers(500, 1025)), 5), "colbert": lst(1024 if q < 3 else 0, 3)})
    for method in ofuse.METHODS:
        s, i, bd = _fuse_gpu(eng, cases, 1500, method=method)
        s, i, bd = s.cpu().numpy(), i.cpu().numpy(), bd.cpu().numpy()
        for q, c in enumerate(cases):
            rows = ofuse.fuse(c["dense"], c["bm25"], c["colbert"], method=method)
            m = min(len(rows), 1500)
            got = [(int(a), float(b)) for a, b in zip(i[q, :m], s[q, :m])]
            assert len({a for a, _ in got}) == m and (i[q, m:] == -1).all(), "ids are unique and the tail is padding"
            want = {r["id"]: r["score"] for r in rows}
            if m == len(rows):
                assert {a for a, _ in got} == set(want)
            for a, b in got:
                assert abs(b - want[a]) <= 1e-5 * max(1.0, abs(want[a])), (method, q, a, b, want[a])
            # output order: score descending, equal fp32 scores by ascending id
            for (a0, b0), (a1, b1) in zip(got[:-1], got[1:]):
                assert b0 > b1 or (b0 == b1 and a0 < a1), (method, q, (a0, b0), (a1, b1))
            # nothing better was left out
            if m < len(rows):
                assert min(b for _, b in got) >= sorted(want.values(), reverse=True)[m - 1] - 1e-5


def test_bm25_device_built_scale_model_matches_oracle(eng):
    """SURVEY 8d config 4's 50 000-doc / 5 000-term scale model, index built on the device by the bench's
    generator, scored by the kernel, checked against the fp64 oracle on the same postings."""
    from legal_rag_b200 import synth
    index, st = synth.bm25_synthetic_index(50_000, 5_000, 10, "cuda", chunk_docs=16_000, keep_tf=True)
    csr = obm25.CsrBM25(index.indptr.cpu().numpy(), index.doc_id.cpu().numpy().astype(np.int64), st["tf"].cpu().numpy().astype(np.int64),
                        st["doc_len"].cpu().numpy())
    assert csr.avgdl == pytest.approx(st["avgdl"])
    # the device builder's impacts are the oracle's per-posting contributions
    t_of = np.repeat(np.arange(5_000), np.diff(csr.indptr))
    f = csr.tf.astype(np.float64)
    imp = csr.idf[t_of] * (f * 2.5 / (f + 1.5 * (0.25 + 0.75 * csr.doc_len[csr.doc_id] / csr.avgdl)))
    np.testing.assert_allclose(index.impact.cpu().numpy(), imp, rtol=2e-7)
    qi, qt, mx = synth.bm25_synthetic_queries(256, 5_000, 11, "cuda")
    s, i = eng.bm25_topk(index, qi, qt, mx, 100)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    qi, qt = qi.cpu().numpy(), qt.cpu().numpy()
    O = [csr.search(qt[qi[j]:qi[j + 1]].tolist(), 150) for j in range(256)]
    check_topk_parity(s, i, np.stack([o[0] for o in O]), np.stack([o[1] for o in O]), 100, TAU_FP32, what="bm25-scale-model")


# ---------------------------------------------------------------- batched hybrid pipeline (one shard)
def test_hybrid_shard_matches_oracle_pipeline(eng):
    """engine.HybridShard on one GPU: per-channel top-kc -> candidate union -> MaxSim -> three-channel fusion,
    against the same pipeline assembled from the oracle pieces."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    from oracle import maxsim as omaxsim
    rng = np.random.default_rng(77)
    N, d, V, nq, Ld, Lq, kc, k = 4000, 128, 800, 12, 32, 8, 20, 10
    Xd, Xr = _bf16(_unit(rng, (N, d)))
    Qd, Qr = _bf16(_unit(rng, (nq, d)))
    docs = [rng.integers(0, V, int(rng.integers(5, 30))) for _ in range(N)]
    host = Bm25HostIndex.from_token_ids(docs, V)
    csr = obm25.CsrBM25.from_token_ids(docs, V)
    queries = [rng.integers(0, V, int(rng.integers(2, 6))).tolist() for _ in range(nq)]
    qi, qt, mx = host.encode_queries(queries)
    Td, Tr = _bf16(_unit(rng, (N, Ld, 128)))
    Qtd, Qtr = _bf16(_unit(rng, (nq, Lq, 128)))
    shard = eng.HybridShard(Xd, host.to_device("cuda"), Td, None, id_base=0, tok_row_base=0, tok_rows_total=N)
    for method in ("weighted_sum", "rrf_norm_blend"):
        s, i = shard.search_device(Qd, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, Qtd, k=k, kc=kc, method=method)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        D, I = odense.flat_ip_topk(Qr, Xr, kc)
        for q in range(nq):
            dl = list(zip(I[q].tolist(), D[q].tolist()))
            bs, bi = csr.search(queries[q], kc)
            bl = list(zip(bi.tolist(), bs.tolist()))
            cand = np.array([[r["id"] for r in ofuse.fuse(dl, bl, [], method=method)]])
            cs = omaxsim.maxsim_scores(Qtr[q:q + 1], Tr, None, cand)[0]
            order = np.lexsort((cand[0], -cs))[:kc]
            cl = [(int(cand[0][o]), float(cs[o])) for o in order]
            ref = ofuse.fuse(dl, bl, cl, method=method)
            m = min(len(ref), k + 8)
            check_topk_parity(s[q:q + 1], i[q:q + 1], np.array([[r["score"] for r in ref[:m]]]), np.array([[r["id"] for r in ref[:m]]]),
                              k, 2e-3, what=f"hybrid-{method}")
    # the host-buffer entry (pinned inputs; the query token matrix travels on a copy stream while the scans run) returns what
    # search_device returns, call after call (the upload buffer is recycled by the allocator between calls)
    host_in = [t.cpu().pin_memory() for t in (Qd, torch.from_numpy(qi), torch.from_numpy(qt), Qtd)]
    for mode in ("rerank", "scan"):
        want_s, want_i = shard.search_device(Qd, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, Qtd, k=k, kc=kc,
                                             colbert_mode=mode)
        for _ in range(4):
            junk = torch.randn(nq * Lq * 128, device="cuda")        # something else takes and returns memory in between
            del junk
            got_s, got_i = shard.search(host_in[0], host_in[1], host_in[2], mx, host_in[3], k=k, kc=kc, colbert_mode=mode)
            assert torch.equal(got_s, want_s.cpu()) and torch.equal(got_i, want_i.cpu())
    # ColBERT as a first-stage channel over the whole corpus (the reference's own arrangement, hybrid_retriever.py:299)
    s, i = shard.search_device(Qd, torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, Qtd, k=k, kc=kc, method="rrf_norm_blend",
                               colbert_mode="scan")
    s, i = s.cpu().numpy(), i.cpu().numpy()
    all_docs = np.arange(N, dtype=np.int64)[None, :]
    for q in range(nq):
        dl = list(zip(I[q].tolist(), D[q].tolist()))
        bs, bi = csr.search(queries[q], kc)
        cs = omaxsim.maxsim_scores(Qtr[q:q + 1], Tr, None, all_docs)[0]
        order = np.lexsort((all_docs[0], -cs.astype(np.float64)))[:kc]
        ref = ofuse.fuse(dl, list(zip(bi.tolist(), bs.tolist())), [(int(o), float(cs[o])) for o in order], method="rrf_norm_blend")
        m = min(len(ref), k + 8)
        check_topk_parity(s[q:q + 1], i[q:q + 1], np.array([[r["score"] for r in ref[:m]]]), np.array([[r["id"] for r in ref[:m]]]),
                          k, 2e-3, what="hybrid-colbert-scan")


def test_side_by_side_scans_equal_the_serial_step(eng):
    """HybridShard.dense_sms > 0: the dense scan confined to that many SMs and the BM25 scan on the others at the same time
    (lrag_sm_reserve / *_part entry points) return bit for bit what the two scans return one after the other, for several
    splits, repeated calls (the start counter accumulates) and a BM25 grid smaller than its share of the machine."""
    rng = np.random.default_rng(5)
    N, d, V, nq, kc, k = 60_000, 256, 5000, 300, 50, 20
    Xd, _ = _bf16(_unit(rng, (N, d)))
    Qd, _ = _bf16(_unit(rng, (nq, d)))
    from legal_rag_b200 import synth
    index, _ = synth.bm25_synthetic_index(N, V, 3, "cuda", mean_len=20.0)
    qi, qt, mx = synth.bm25_synthetic_queries(nq, V, 4, "cuda")
    shard = eng.HybridShard(Xd, index)
    ref = shard.search_device(Qd, qi, qt, mx, None, k=k, kc=kc)
    d0 = eng.dense_topk(Xd, Qd, kc)
    b0 = eng.bm25_topk(index, qi, qt, mx, kc)
    for D in (68, 8, 140, 64):
        shard.dense_sms = D
        for _ in range(3):
            s, i = shard.search_device(Qd, qi, qt, mx, None, k=k, kc=kc)
            assert torch.equal(s, ref[0]) and torch.equal(i, ref[1]), f"dense on {D} SMs changed the fused result"
        d1 = eng.dense_topk(Xd, Qd, kc, max_ctas=D)
        b1 = eng.bm25_topk(index, qi, qt, mx, kc, max_sms=148 - D)
        assert torch.equal(d0[0], d1[0]) and torch.equal(d0[1], d1[1]), f"dense scan on {D} SMs changed its result"
        assert torch.equal(b0[0], b1[0]) and torch.equal(b0[1], b1[1]), f"BM25 scan on {148 - D} SMs changed its result"
    timed = shard.tune_partition(Qd, qi, qt, mx, None, candidates=(0, 68, 100, 1000), steps=1, k=k, kc=kc)
    assert {0, 68, 100} <= set(timed) and 1000 not in timed and shard.dense_sms in timed and all(ms > 0 for ms in timed.values())
    s, i = shard.search_device(Qd, qi, qt, mx, None, k=k, kc=kc)
    assert torch.equal(s, ref[0]) and torch.equal(i, ref[1])
    torch.cuda.synchronize()


def test_sm_reservation_times_out_when_its_partner_never_comes(eng):
    """lrag_sm_reserve holds its SMs until the counter reaches the target -- or the timeout passes: a reservation whose BM25
    launch failed must not hold a part of the machine for ever; one whose target is already reached exits at once."""
    import time
    from legal_rag_b200 import _native
    lib = _native.init(0)
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert lib.lrag_sm_reserve(8, counter.data_ptr(), 5, 20, stream) == 0          # never reached: 20 ms
    torch.cuda.synchronize()
    waited = time.perf_counter() - t0
    assert 0.005 < waited < 1.0, waited
    counter.fill_(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert lib.lrag_sm_reserve(8, counter.data_ptr(), 5, 2000, stream) == 0        # already reached
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.5
    assert lib.lrag_sm_reserve(1000, counter.data_ptr(), 5, 20, stream) != 0       # more SMs than the device has


# ---------------------------------------------------------------- gathered inner products (graph-expansion scoring)
@pytest.mark.parametrize("N,d,nq,C", [(500, 768, 1, 800), (64, 64, 3, 10), (3000, 1024, 17, 33), (10, 8, 2, 5)])
def test_gather_scores_match_numpy(eng, N, d, nq, C):
    rng = np.random.default_rng(N + C)
    Xd, Xr = _bf16(_unit(rng, (N, d)))
    Qd, Qr = _bf16(_unit(rng, (nq, d)))
    rows = rng.integers(0, N, (nq, C)).astype(np.int64)
    rows[0, 0] = -1
    rows[-1, -1] = N              # out of range -> -inf
    got = eng.gather_scores(Xd, Qd, torch.from_numpy(rows).cuda()).cpu().numpy()
    want = np.einsum("qd,qcd->qc", Qr, Xr[np.clip(rows, 0, N - 1)])
    ok = (rows >= 0) & (rows < N)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-4, atol=1e-5)
    assert np.isneginf(got[~ok]).all()


# ---------------------------------------------------------------- batched full-corpus MaxSim
@pytest.mark.parametrize("Nd,Ld,Lq,nq,k", [(700, 128, 32, 9, 100), (1000, 64, 7, 4, 10), (333, 32, 32, 5, 20), (97, 256, 20, 3, 50),
                                            (5000, 128, 32, 64, 100)])
def test_maxsim_scan_matches_oracle(eng, Nd, Ld, Lq, nq, k):
    rng = np.random.default_rng(Nd + Ld + nq)
    Dd, Dr = _bf16(_unit(rng, (Nd, Ld, 128)))
    Qd, Qr = _bf16(_unit(rng, (nq, Lq, 128)))
    doclen = rng.integers(1, Ld + 1, Nd).astype(np.int32)
    doclen[:5] = Ld
    doclen[5] = 0                                          # an empty document scores Lq * -9999
    cand = np.tile(np.arange(Nd, dtype=np.int64), (nq, 1))
    want = omaxsim.maxsim_scores(Qr, Dr, doclen, cand)
    got = eng.maxsim_scan_scores(Dd, torch.from_numpy(doclen).cuda(), Qd).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)
    # no doclen array = every document is full length
    got_full = eng.maxsim_scan_scores(Dd, None, Qd).cpu().numpy()
    np.testing.assert_allclose(got_full, omaxsim.maxsim_scores(Qr, Dr, None, cand), rtol=2e-4, atol=2e-4)
    # ranked, in query groups small enough to force several launches; equals the gather kernel on all candidates
    s, i = eng.maxsim_scan_topk(Dd, torch.from_numpy(doclen).cuda(), Qd, k, id_base=11, max_score_bytes=4 * Nd * 8)
    O_s, O_i = omaxsim.rerank_topk(Qr, Dr, doclen, cand, min(k + 10, Nd))
    check_topk_parity(s.cpu().numpy(), i.cpu().numpy() - 11 * (i.cpu().numpy() >= 0), O_s, O_i, k, 1e-4, what="maxsim-scan", floor=1.0)
    s2, i2 = eng.maxsim_rerank(Dd, torch.from_numpy(doclen).cuda(), Qd, torch.from_numpy(cand).cuda(), k, id_base=11)
    check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), s2.cpu().numpy(), i2.cpu().numpy(), min(k, Nd), 1e-5, what="scan-vs-gather", floor=1.0)


def test_maxsim_scan_chunking_changes_nothing(eng, monkeypatch):
    """The scan's work split (doc tiles per chunk, capped so that the chunks in flight stay in L2: LRAG_SCAN_L2_MB) is a
    scheduling choice: the score matrix is bit-identical for tiny, default and unlimited chunks, on a store large enough
    for the cap to bind, and a sample of it matches the oracle."""
    rng = np.random.default_rng(8)
    Nd, Ld, Lq, nq = 12_000, 128, 32, 64
    Dd, Dr = _bf16(_unit(rng, (Nd, Ld, 128)))
    Qd, Qr = _bf16(_unit(rng, (nq, Lq, 128)))
    doclen = rng.integers(1, Ld + 1, Nd).astype(np.int32)
    dl = torch.from_numpy(doclen).cuda()
    outs = []
    for mb in ("0", None, "4096"):
        if mb is None:
            monkeypatch.delenv("LRAG_SCAN_L2_MB", raising=False)
        else:
            monkeypatch.setenv("LRAG_SCAN_L2_MB", mb)
        outs.append(eng.maxsim_scan_scores(Dd, dl, Qd))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    rows = rng.choice(Nd, 300, replace=False).astype(np.int64)
    want = omaxsim.maxsim_scores(Qr[:3], Dr, doclen, np.tile(rows, (3, 1)))
    np.testing.assert_allclose(outs[1][:3].cpu().numpy()[:, rows], want, rtol=2e-4, atol=2e-4)


def test_maxsim_scan_rejects_row_lengths_that_do_not_tile(eng):
    D = torch.zeros((10, 96, 128), dtype=torch.bfloat16, device="cuda")
    Q = torch.zeros((2, 32, 128), dtype=torch.bfloat16, device="cuda")
    assert not eng.maxsim_scan_supported(96)
    with pytest.raises(Exception, match="must be 32, 64, 128 or 256"):
        eng.maxsim_scan_scores(D, None, Q)


# ---------------------------------------------------------------- BASELINE.json configs[0], end to end
def test_config0_ucc_hybrid_matches_reference_path(eng):
    """UCC corpus, random unit embeddings (seed 42), 1024 synthetic queries (seeds 43/44), dense + BM25 top-100 and
    weighted fusion 0.6/0.4 -- every query against the literal oracle pipeline (flat-IP, BM25Okapi.get_scores + stable sort, _fuse)."""
    import argparse
    import bench
    wl = bench.UccWorkload(argparse.Namespace(nq=1024, k=100), 0, 1, torch.device("cuda", 0))
    wl.setup()
    s, i = wl.step()
    s, i = s.cpu().numpy(), i.cpu().numpy()
    docs, V, X, Q, queries = wl._corpus()
    Xr = torch.from_numpy(X).to(torch.bfloat16).float().numpy()      # the engine's inputs are the bf16-rounded embeddings
    Qr = torch.from_numpy(Q).to(torch.bfloat16).float().numpy()
    lit = obm25.BM25Okapi([[str(t) for t in d] for d in docs])
    for method in ("weighted_sum",):
        for q in range(0, 1024, 4):
            ds, di = odense.flat_ip_topk(Qr[q:q + 1], Xr, 100)
            bs, bi = obm25.search(lit, [str(t) for t in queries[q]], 100)
            ref = ofuse.fuse(list(zip(di[0].tolist(), ds[0].tolist())), list(zip(bi.tolist(), bs.tolist())), [], method=method,
                             w_dense=0.6, w_bm25=0.4)
            m = min(len(ref), 120)
            check_topk_parity(s[q:q + 1], i[q:q + 1], np.array([[r["score"] for r in ref[:m]]]), np.array([[r["id"] for r in ref[:m]]]),
                              100, 2e-3, what=f"config0-{method}")


def test_graphed_single_query_equals_eager_calls(eng):
    """configs[0]'s online shape: one query at a time through a captured CUDA graph; replays with new inputs (short, long,
    empty and out-of-vocabulary term lists) return exactly what the eager kernel calls return."""
    import argparse
    import bench
    wl = bench.UccWorkload(argparse.Namespace(nq=16, k=100), 0, 1, torch.device("cuda", 0))
    wl.setup()
    docs, V, X, Q, queries = wl._corpus()
    queries[3] = []                                   # no tokens: BM25 contributes zero scores only
    queries[5] = [-1, -1, queries[5][0]]              # out-of-vocabulary tokens
    queries[7] = (queries[7] * 8)[:32]                # as long as the graph allows, repeated tokens
    g = eng.GraphedHybridQuery(wl.X, wl.index, k=100, nq=1, max_terms=32)
    for q in list(range(16)) + [2, 0]:                # replays in any order, the same query twice
        s, i = g.search(torch.from_numpy(Q[q:q + 1]), [queries[q]])
        qi = torch.tensor([0, len(queries[q])], dtype=torch.int64).cuda()
        qt = torch.tensor(queries[q] if queries[q] else [], dtype=torch.int32).cuda()
        Qd = torch.from_numpy(Q[q:q + 1]).cuda().to(torch.bfloat16)
        d = eng.dense_topk(wl.X, Qd, 100)
        b = eng.bm25_topk(wl.index, qi, qt, max(1, len(queries[q])), 100)
        es, ei = eng.fuse_topk(d, b, None, k=100, method="weighted_sum", w_dense=0.6, w_bm25=0.4)
        np.testing.assert_array_equal(i.numpy(), ei.cpu().numpy(), err_msg=f"query {q}")
        np.testing.assert_array_equal(s.numpy(), es.cpu().numpy(), err_msg=f"query {q}")
    with pytest.raises(eng.LragError):
        g.search(torch.from_numpy(Q[:1]), [[1] * 33])


def test_bm25_kernel_reproduces_the_upstream_readme_example(eng):
    """The rank_bm25 README's published answer (see tests/test_oracle.py) through the CUDA path."""
    from legal_rag_b200.bm25_index import Bm25HostIndex
    corpus = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
    host = Bm25HostIndex.from_tokens([d.split(" ") for d in corpus])
    qi, qt, mx = host.encode_queries(["windy London".split(" ")])
    s, i = eng.bm25_topk(host.to_device("cuda"), torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), mx, 3)
    assert i[0].tolist() == [1, 0, 2]
    np.testing.assert_allclose(s[0].cpu().numpy(), [0.93729472, 0.0, 0.0], rtol=1e-6, atol=1e-7)
