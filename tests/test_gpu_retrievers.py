"""GPU tests of the reference-shaped classes: artifacts are written by this package's builders with
deterministic stand-in encoders, searched through DenseRetriever / BM25Retriever / ColBERTRetriever /
HybridRetriever exactly as RagPipeline would, and every stage is compared with the CPU oracle."""
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

QUESTIONS = [
    "What remedies does a buyer have when the seller fails to deliver conforming goods?",
    "perfection of a security interest in deposit accounts",
    "negotiable instrument holder in due course",
    "Letter of credit issuer obligations and the independence principle",
    "warranty of merchantability disclaimer",
]


@pytest.fixture(scope="module")
def world(golden_dir, tmp_path_factory):
    from legal_rag_b200.config import AppConfig
    from legal_rag_b200.retrieval import builders, encoders
    from legal_rag_b200.schemas import LawChunk
    z = np.load(os.path.join(golden_dir, "ucc_corpus.npz"))
    lens, flat, vocab, ids = z["doc_len"], z["tokens"].astype(np.int64), z["vocab"], z["ids"]
    off = np.concatenate([[0], np.cumsum(lens)])
    texts = [" ".join(vocab[flat[off[i]:off[i + 1]]]) for i in range(len(lens))]
    chunks = [LawChunk(id=str(ids[i]), law_name="UCC", article_no=str(i), article_id=str(i), text=texts[i], lang="en")
              for i in range(len(texts))]
    root = tmp_path_factory.mktemp("index")
    cfg = AppConfig()
    r = cfg.retrieval
    r.faiss_index_file, r.faiss_meta_file = str(root / "faiss" / "faiss.index"), str(root / "faiss" / "faiss_meta.jsonl")
    r.bm25_index_file = str(root / "bm25.pkl")
    r.colbert_index_path, r.colbert_meta_file = str(root / "colbert"), str(root / "colbert" / "colbert_meta.jsonl")
    r.colbert_doc_maxlen = 96
    dense_enc, tok_enc = encoders.HashingDenseEncoder(768), encoders.HashingTokenEncoder(doc_maxlen=96)
    encoders.register_dense_encoder(lambda name, dev: dense_enc)
    encoders.register_token_encoder(lambda name, dev: tok_enc)
    builders.build_faiss_index(cfg, chunks, flat=False)          # HNSW container, like the reference writes
    builders.build_bm25_index(cfg, chunks)
    builders.build_colbert_index(cfg, chunks)
    yield {"cfg": cfg, "chunks": chunks, "texts": texts, "dense_enc": dense_enc, "tok_enc": tok_enc, "root": root}
    encoders.register_dense_encoder(None)
    encoders.register_token_encoder(None)


def _oracle_channels(world, question, k):
    from legal_rag_b200.retrieval import encoders
    texts, chunks = world["texts"], world["chunks"]
    X = world["dense_enc"].encode(texts)
    q = world["dense_enc"].encode_queries([question])
    dS, dI = odense.flat_ip_topk(q, X, k)
    lit = obm25.BM25Okapi([encoders.tokenize_en(t) for t in texts])
    bS, bI = obm25.search(lit, encoders.default_query_tokenizer()(question), k)
    D = np.zeros((len(texts), 96, 128), np.float32); dl = np.zeros(len(texts), np.int64)
    for i, t in enumerate(texts):
        m = world["tok_enc"].encode_doc(t)
        D[i, :len(m)] = m; dl[i] = len(m)
    Qm = world["tok_enc"].encode_query(question)[None]
    cS, cI = omaxsim.rerank_topk(Qm, D, dl, np.arange(len(texts))[None], k)
    return (dS[0], dI[0]), (bS, bI), (cS[0], cI[0])


def test_channels_match_oracle_through_the_retriever_classes(world):
    from legal_rag_b200.retrieval import HybridRetriever
    hr = HybridRetriever(world["cfg"])
    assert hr.colbert is not None and hr.graph is None
    row = {c.id: i for i, c in enumerate(world["chunks"])}
    k = 100
    for qn in QUESTIONS[:3]:
        (dS, dI), (bS, bI), (cS, cI) = _oracle_channels(world, qn, k + 20)
        hits = hr.search_dense(qn, k)
        assert [h.rank for h in hits] == list(range(1, k + 1)) and hits[0].score_breakdown["channel"] == ["dense"]
        assert hits[0].semantic_score == hits[0].score and hits[0].source == "retriever"
        check_topk_parity(np.array([[h.score for h in hits]]), np.array([[row[h.chunk.id] for h in hits]]), dS[None], dI[None], k,
                          1e-2, what="cls-dense", floor=0.1)
        hits = hr.search_bm25(qn, k)
        assert len(hits) == k                                   # zero-score documents are returned too
        check_topk_parity(np.array([[h.score for h in hits]]), np.array([[row[h.chunk.id] for h in hits]]), bS[None], bI[None], k,
                          1e-3, what="cls-bm25")
        hits = hr.search_colbert(qn, k)
        check_topk_parity(np.array([[h.score for h in hits]]), np.array([[row[h.chunk.id] for h in hits]]), cS[None], cI[None], k,
                          1e-2, what="cls-colbert", floor=1.0)
        assert hits[0].score_breakdown["channel"] == ["colbert"] and "colbert_raw" in hits[0].score_breakdown


@pytest.mark.parametrize("method", ["rrf_norm_blend", "weighted_sum", "rrf", "wrrf"])
def test_fuse_and_search_match_oracle(world, method):
    from legal_rag_b200.retrieval import HybridRetriever
    cfg = world["cfg"]
    cfg.retrieval.fusion_method = method
    cfg.retrieval.top_k = 100
    hr = HybridRetriever(cfg)
    row = {c.id: i for i, c in enumerate(world["chunks"])}
    qn = QUESTIONS[0]
    d, b, c = hr.search_dense(qn, 100), hr.search_bm25(qn, 100), hr.search_colbert(qn, 100)
    # _fuse breaks exact score ties by first appearance (dense, bm25, colbert lists); number the chunks that way
    row = {}
    for h in d + b + c:
        row.setdefault(h.chunk.id, len(row))
    pairs = lambda hs: [(row[h.chunk.id], float(np.float32(h.score))) for h in hs]   # noqa: E731
    ref = ofuse.fuse(pairs(d), pairs(b), pairs(c), method=method, w_dense=0.6, w_bm25=0.4, w_colbert=0.35, rrf_k=60, alpha=0.5)
    fused = hr._fuse(dense_hits=d, bm25_hits=b, colbert_hits=c)
    assert len(fused) == len(ref) and [h.rank for h in fused] == list(range(1, len(ref) + 1))
    check_topk_parity(np.array([[h.score for h in fused]]), np.array([[row[h.chunk.id] for h in fused]]),
                      np.array([[r["score"] for r in ref]]), np.array([[r["id"] for r in ref]]), len(ref), 1e-3, what=f"cls-fuse-{method}")
    by = {r["id"]: r for r in ref}
    for h in fused[:50]:
        r, sb = by[row[h.chunk.id]], h.score_breakdown
        assert sb["fusion_method"] == method and sb["rrf_k"] == 60 and set(sb["channel_contrib"]) == {"dense", "bm25", "colbert"}
        for key in ("rrf_norm", "weighted_sum", "dense_norm", "bm25_norm", "colbert_norm"):
            assert sb[key] == pytest.approx(r[key], rel=1e-3, abs=1e-5)
        for ch in ("dense", "bm25", "colbert"):
            assert sb["channel_contrib"][ch] == pytest.approx(r["channel_contrib"][ch], rel=1e-3, abs=1e-5)
    # full search: min_final_score filter (0.2) + slice; positional (question, llm, top_k, decision) order
    out = hr.search(qn, None, 10, None)
    want = [r for r in ref if r["score"] >= 0.2][:12]
    assert 0 < len(out) <= 10 and all(h.score >= 0.2 for h in out)
    check_topk_parity(np.array([[h.score for h in out]]), np.array([[row[h.chunk.id] for h in out]]),
                      np.array([[r["score"] for r in want]]), np.array([[r["id"] for r in want]]), len(out), 1e-3, what="cls-search")
    cfg.retrieval.fusion_method, cfg.retrieval.top_k = "rrf_norm_blend", 10


def test_search_batch_equals_per_question_search(world):
    from legal_rag_b200.retrieval import HybridRetriever
    cfg = world["cfg"]
    hr = HybridRetriever(cfg)
    row = {c.id: i for i, c in enumerate(world["chunks"])}
    batch = hr.search_batch(QUESTIONS, top_k=10)
    for qn, got in zip(QUESTIONS, batch):
        one = hr.search(qn, None, 10)
        check_topk_parity(np.array([[h.score for h in got]]), np.array([[row[h.chunk.id] for h in got]]),
                          np.array([[h.score for h in one]]), np.array([[row[h.chunk.id] for h in one]]), len(one), 1e-3, what="cls-batch")


@pytest.mark.parametrize("method", ["rrf_norm_blend", "weighted_sum"])
def test_search_fast_path_equals_general_path(world, method):
    """Without a ColBERT channel / graph / reranker, search() answers a question with one captured CUDA graph
    (HybridRetriever._search_graphed); the hits, scores and score_breakdown are those of the general per-channel path."""
    import copy
    from legal_rag_b200.retrieval import HybridRetriever
    cfg = copy.deepcopy(world["cfg"])
    cfg.retrieval.enable_colbert = False
    cfg.retrieval.fusion_method = method
    cfg.retrieval.top_k = 50
    cfg.retrieval.min_final_score = 0.05
    fast, slow = HybridRetriever(cfg), HybridRetriever(cfg)
    slow._search_graphed = lambda *a, **k: None
    for qn in QUESTIONS:
        for top_k in (10, 50):
            a, b = fast.search(qn, None, top_k), slow.search(qn, None, top_k)
            assert fast.fast_path_used and not slow.fast_path_used
            assert len(a) == len(b) > 0
            sa = {h.chunk.id: h for h in a}
            for hb in b:
                ha = sa.get(hb.chunk.id)
                if ha is None:                       # only a tie at the cut may differ
                    assert abs(hb.score - b[-1].score) < 1e-6
                    continue
                assert abs(ha.score - hb.score) < 1e-6
                assert ha.source == hb.source == "retriever"
                assert set(ha.score_breakdown) == set(hb.score_breakdown)
                assert sorted(ha.score_breakdown["channel"]) == sorted(hb.score_breakdown["channel"])
                for key in ("rrf_norm", "weighted_sum", "dense_norm", "bm25_norm"):
                    assert abs(ha.score_breakdown[key] - hb.score_breakdown[key]) < 1e-6
            assert [h.rank for h in a] == list(range(1, len(a) + 1))
            assert all(x.score >= y.score for x, y in zip(a, a[1:]))


def test_search_fast_path_from_several_host_threads(world):
    """the reference's API calls HybridRetriever.search from Starlette's worker threads.  The captured pipeline
    works on static buffers, so concurrent callers take turns on it (GraphedHybridQuery.lock) and the first one captures it
    while the others wait; every thread must get the single-threaded answer."""
    import copy
    import threading
    from legal_rag_b200.retrieval import HybridRetriever
    cfg = copy.deepcopy(world["cfg"])
    cfg.retrieval.enable_colbert = False
    hr, ref = HybridRetriever(cfg), HybridRetriever(cfg)
    want = {qn: [(h.chunk.id, h.score) for h in ref.search(qn, None, 10)] for qn in QUESTIONS}
    assert ref.fast_path_used
    errors = []

    def worker(seed):
        try:
            for it in range(25):
                qn = QUESTIONS[(seed + it) % len(QUESTIONS)]
                got = [(h.chunk.id, h.score) for h in hr.search(qn, None, 10)]
                if got != want[qn]:
                    errors.append(f"thread {seed} iteration {it}: {got[:3]} != {want[qn][:3]}")
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {seed}: {e!r}")

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert hr.fast_path_used and len(hr._graphs) == 1


def test_bm25_long_question_with_unknown_tokens_is_scored_whole(world):
    """rank_bm25's get_scores has no token limit (bm25_retriever.py:74); the kernel's term table holds 128 IN-VOCABULARY
    tokens.  Out-of-vocabulary tokens (jieba's whitespace / punctuation tokens, mostly) score nothing and must not use up
    that budget: a 300-token question with 60 known tokens ranks exactly like its known tokens alone."""
    from legal_rag_b200.retrieval import BM25Retriever
    bm = BM25Retriever(world["cfg"], tokenizer=lambda q: q.split(" "))
    known = (QUESTIONS[0] + " " + QUESTIONS[1] + " " + QUESTIONS[3]).lower().replace("?", "").split(" ")
    known = (known * 3)[:60]
    rng = np.random.default_rng(5)
    noisy = []
    for w in known:
        noisy.extend([" ", "\u3000", f"zzqx{rng.integers(1 << 30)}", "?"])
        noisy.append(w)
    assert len(noisy) == 300
    s0, i0 = bm.search_ids([known], 50)
    s1, i1 = bm.search_ids([noisy, [], ["zzqx-only-unknown"], known], 50)
    assert torch.equal(s0[0], s1[0]) and torch.equal(i0[0], i1[0])
    assert torch.equal(s0[0], s1[3]) and torch.equal(i0[0], i1[3])
    assert float(s1[1].abs().max()) == 0.0 and float(s1[2].abs().max()) == 0.0     # the reference ranks all-zero scores too


def test_incremental_add_is_searchable_and_persisted(world):
    from legal_rag_b200.retrieval import VectorStore, artifacts, builders
    from legal_rag_b200.schemas import LawChunk
    cfg = world["cfg"]
    store = VectorStore.from_config(cfg)
    store.load()
    n0 = store.index.ntotal
    new = LawChunk(id="new::1", law_name="UCC", article_no="x", article_id="x", lang="en",
                   text="zebra quagga okapi unmistakable marker sentence about striped equids")
    assert builders.IncrementalDenseBuilder(cfg, store).add_chunks([new]) == 1
    assert store.index.ntotal == n0 + 1 and len(store.chunks) == n0 + 1
    hits = store.search("zebra quagga okapi striped equids", 3)
    assert hits[0][0].id == "new::1"
    X, info = artifacts.read_faiss_index(cfg.retrieval.faiss_index_file)      # rewritten as a flat file, row count grown
    assert X.shape[0] == n0 + 1 and len(artifacts.read_meta_jsonl(cfg.retrieval.faiss_meta_file)) == n0 + 1
    # ids that are already indexed are skipped (incremental_dense_builder.py:51-58)
    assert builders.IncrementalDenseBuilder(cfg, store).add_chunks([new, new]) == 0
    assert store.index.ntotal == n0 + 1


def test_graph_retriever_scoring_matches_reference_golden(golden_dir, tmp_path):
    """retrieval.GraphRetriever (gathered inner products on the GPU) against the hits the reference's own
    GraphRetriever.search produced for the same nodes and vectors (tests/golden/graph_golden.json)."""
    import json
    from types import SimpleNamespace
    from legal_rag_b200.retrieval import GraphRetriever, GpuFlatIndex
    from legal_rag_b200.schemas import LawChunk
    g = json.load(open(os.path.join(golden_dir, "graph_golden.json")))
    vecs, qvec = np.array(g["vectors"], dtype=np.float32), np.array(g["qvec"], dtype=np.float32)
    chunks = [LawChunk(**c) for c in g["chunks"]]
    nodes = [SimpleNamespace(**n) for n in g["nodes"]]

    class Store:                                  # the slice of VectorStore the channel uses
        def __init__(self):
            self.index, self.chunks = GpuFlatIndex.from_numpy(vecs), chunks
        def load(self):
            pass
        def _embed(self, texts, is_query=False):
            return qvec[None, :]

    class Graph:
        def walk(self, **kw):
            return nodes

    for case in g["cases"]:
        rcfg = SimpleNamespace(graph_walk_depths={"default": 2}, graph_limit=800, graph_rel_types=None, graph_min_conf=0.0,
                               graph_depth_gamma=case["gamma"])
        gr = GraphRetriever(SimpleNamespace(retrieval=rcfg), graph=Graph(), store=Store())
        hits = gr.search("the question", [SimpleNamespace(chunk=chunks[0])], lang=case["lang"], top_k=case["top_k"])
        want = case["hits"]
        row = {c.id: i for i, c in enumerate(chunks)}
        # the golden list is already cut at top_k, so the last tie run may only be matched as a subset
        check_topk_parity(np.array([[h.score for h in hits]]), np.array([[row[h.chunk.id] for h in hits]]),
                          np.array([[h["score"] for h in want]]), np.array([[row[h["id"]] for h in want]]), len(want), 1e-2,
                          what="graph", floor=0.1)
        assert all(h.source == "graph" and h.rank == r for r, h in enumerate(hits, start=1))
        by_id = {h["id"]: h for h in want}
        for h in hits:
            if h.chunk.id in by_id:
                b = by_id[h.chunk.id]["breakdown"]
                assert h.score_breakdown["depth_decay"] == pytest.approx(b["depth_decay"], rel=1e-9)
                assert h.score_breakdown["relation_weight"] == b["relation_weight"] and h.score_breakdown["edge_conf"] == b["edge_conf"]
                assert h.score_breakdown["graph_depth"] == b["graph_depth"]
                assert h.score_breakdown["semantic"] == pytest.approx(b["semantic"], abs=5e-3)     # bf16 corpus rows


def test_engine_is_reentrant_across_host_threads():
    """SURVEY 8b threading: `search` is called concurrently from Starlette's thread pool.  Four host threads, each on its
    own CUDA stream, hammer the three channel kernels; every result must equal the single-threaded one."""
    import threading
    from legal_rag_b200 import engine
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(21)
    N, d, V, nq = 30_000, 256, 2000, 24
    X = torch.from_numpy(rng.standard_normal((N, d)).astype(np.float32)).cuda().to(torch.bfloat16)
    Q = torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)).cuda().to(torch.bfloat16)
    docs = [rng.integers(0, V, int(rng.integers(5, 40))) for _ in range(N)]
    host = Bm25HostIndex.from_token_ids(docs, V)
    dev = host.to_device("cuda")
    qi, qt, mx = host.encode_queries([rng.integers(0, V, 5).tolist() for _ in range(nq)])
    qi, qt = torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda()
    T = torch.from_numpy(rng.standard_normal((2000, 64, 128)).astype(np.float32)).cuda().to(torch.bfloat16)
    Qt = torch.from_numpy(rng.standard_normal((nq, 16, 128)).astype(np.float32)).cuda().to(torch.bfloat16)
    cand = torch.from_numpy(rng.integers(0, 2000, (nq, 300))).cuda()

    def once():
        a = engine.dense_topk(X, Q, 50)
        b = engine.bm25_topk(dev, qi, qt, mx, 50)
        c = engine.maxsim_scores(T, None, Qt, cand)
        f = engine.fuse_topk(a, b, None, k=50, method="rrf")
        return [t.clone() for t in (*a, *b, c, *f)]

    want = once()
    torch.cuda.synchronize()
    errors = []

    def worker(seed):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(15):
                    got = once()
                    stream.synchronize()
                    for g, w in zip(got, want):
                        if not torch.equal(g, w):
                            errors.append(f"thread {seed}: result differs")
                            return
        except Exception as e:       # noqa: BLE001
            errors.append(f"thread {seed}: {e!r}")

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_colbert_retriever_serves_a_plaid_directory(world, tmp_path):
    """No lrag token store, but a PLAID index directory at the reference's path: it is decompressed at load time and
    searched; scores are the exact MaxSim of the decompressed (4-bit) vectors, as colbert's Searcher would rank them."""
    import copy
    from legal_rag_b200.retrieval import ColBERTRetriever, artifacts, plaid
    from legal_rag_b200.retrieval.colbert_retriever import token_store_path
    cfg = copy.deepcopy(world["cfg"])
    chunks, tok_enc = world["chunks"][:120], world["tok_enc"]
    cfg.retrieval.colbert_index_path = str(tmp_path / "colbert")
    cfg.retrieval.colbert_meta_file = str(tmp_path / "colbert" / "colbert_meta.jsonl")
    d = token_store_path(cfg.retrieval).parent
    docs = [np.asarray(tok_enc.encode_doc(c.text), dtype=np.float32)[:96] for c in chunks]
    plaid.write_plaid_index(d, docs, n_centroids=256, nbits=4, chunk_docs=50)
    artifacts.write_colbert_meta(cfg.retrieval.colbert_meta_file, chunks)
    ColBERTRetriever._instances_by_key.clear()
    r = ColBERTRetriever.from_config(cfg)
    tokens, doclen = plaid.read_plaid_index(d)
    for qn in QUESTIONS[:3]:
        hits = r.search(qn, 10)
        assert len(hits) == 10
        Q = np.asarray(tok_enc.encode_query(qn), dtype=np.float32)[:32][None]
        Qr = torch.from_numpy(Q).to(torch.bfloat16).float().numpy()
        O_s, O_i = omaxsim.rerank_topk(Qr, tokens.float().numpy(), doclen.numpy(), np.arange(len(chunks))[None, :], 20)
        row = {c.id: i for i, c in enumerate(chunks)}
        check_topk_parity(np.array([[s for _, s in hits]]), np.array([[row[c.id] for c, _ in hits]]), O_s, O_i, 10, 1e-3,
                          what="plaid", floor=1.0)
