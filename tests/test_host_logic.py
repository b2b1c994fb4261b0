"""CPU tests of the host side: artifact formats, BM25 index construction, tokenisation, shard maths,
the dedup rule, and the multi-rank gather/merge plumbing on gloo."""
import json
import os
import pickle
import struct

import numpy as np
import pytest

from legal_rag_b200.bm25_index import Bm25HostIndex
from legal_rag_b200.retrieval import artifacts, encoders
from legal_rag_b200.schemas import LawChunk, RetrievalHit
from oracle import bm25 as obm25


def _chunks(n, lang="en"):
    return [LawChunk(id=f"c{i}", law_name="UCC", article_no=str(i), article_id=str(i), text=f"text number {i}", lang=lang)
            for i in range(n)]


# ---------------------------------------------------------------- faiss files
def test_faiss_flat_roundtrip_and_handmade_header(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((37, 24)).astype(np.float32)
    p = tmp_path / "faiss.index"
    artifacts.write_faiss_flat(p, X)
    raw = p.read_bytes()
    assert raw[:4] == b"IxFI" and len(raw) == 45 + 37 * 24 * 4           # 45-byte header
    d, n = struct.unpack_from("<iq", raw, 4)
    assert (d, n) == (24, 37) and struct.unpack_from("<Q", raw, 37)[0] == 37 * 24
    Y, info = artifacts.read_faiss_index(p)
    np.testing.assert_array_equal(X, Y)
    assert info["metric"] == artifacts.METRIC_INNER_PRODUCT and info["hnsw"] is None
    # a byte string assembled by hand, L2 metric
    blob = b"IxF2" + struct.pack("<iqqqBi", 2, 3, 0, 0, 1, 1) + struct.pack("<Q", 6) + np.arange(6, dtype=np.float32).tobytes()
    (tmp_path / "h.index").write_bytes(blob)
    Y, info = artifacts.read_faiss_index(tmp_path / "h.index")
    assert Y.tolist() == [[0, 1], [2, 3], [4, 5]] and info["metric"] == artifacts.METRIC_L2


def test_faiss_hnsw_container_is_unwrapped(tmp_path):
    X = np.random.default_rng(1).standard_normal((11, 8)).astype(np.float32)
    p = tmp_path / "hnsw.index"
    artifacts.write_faiss_hnsw_flat(p, X)
    Y, info = artifacts.read_faiss_index(p)
    np.testing.assert_array_equal(X, Y)
    assert info["outer_fourcc"] == "IHNf" and info["fourcc"] == "IxFI"
    # unknown bytes between the graph and the storage: the reader locates the nested flat index
    raw = p.read_bytes()
    pos = raw.index(b"IxFI", 4)
    (tmp_path / "odd.index").write_bytes(raw[:pos] + b"\x07" * 12 + raw[pos:])
    Y, _ = artifacts.read_faiss_index(tmp_path / "odd.index")
    np.testing.assert_array_equal(X, Y)
    with pytest.raises(ValueError, match="unsupported index type"):
        (tmp_path / "ivf.index").write_bytes(b"IwFl" + b"\0" * 64)
        artifacts.read_faiss_index(tmp_path / "ivf.index")


def test_faiss_hnsw_file_carries_a_walkable_graph(tmp_path):
    """The IndexHNSWFlat writer emits a single-level graph of exact nearest-neighbour links (round-1 finding: the graph was
    empty): levels / offsets / neighbors are consistent with faiss's layout, and a plain level-0 best-first search over the
    links from the entry point finds the true nearest neighbours."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((300, 16)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    M = 8
    p = tmp_path / "hnsw.index"
    artifacts.write_faiss_hnsw_flat(p, X, M=M, ef_construction=40, ef_search=32)
    Y, info = artifacts.read_faiss_index(p, with_graph=True)
    np.testing.assert_array_equal(X, Y)
    g = info["hnsw"]
    n, slots = len(X), 2 * M
    assert g["entry_point"] == 0 and g["max_level"] == 0 and g["efSearch"] == 32 and g["efConstruction"] == 40
    assert (g["levels"] == 1).all() and len(g["levels"]) == n
    np.testing.assert_array_equal(g["offsets"], np.arange(n + 1, dtype=np.uint64) * slots)
    assert g["cum_nneighbor_per_level"][0] == 0 and g["cum_nneighbor_per_level"][1] == slots
    assert abs(sum(g["assign_probas"]) - 1.0) < 1e-6
    nb = g["neighbors"].reshape(n, slots)
    S = X @ X.T
    np.fill_diagonal(S, -np.inf)
    want = np.argsort(-S, axis=1, kind="stable")[:, :slots]
    assert (np.sort(nb, axis=1) == np.sort(want, axis=1)).all()
    # level-0 beam search over the stored links (what a faiss search does once it is on level 0)
    def search(q, ef=32, k=5):
        visited, cand, best = {0}, [(-float(X[0] @ q), 0)], [(float(X[0] @ q), 0)]
        import heapq
        while cand:
            negs, v = heapq.heappop(cand)
            if len(best) >= ef and -negs < min(best)[0]:
                break
            for u in nb[v]:
                if u >= 0 and u not in visited:
                    visited.add(int(u))
                    s_ = float(X[u] @ q)
                    if len(best) < ef or s_ > min(best)[0]:
                        heapq.heappush(cand, (-s_, int(u)))
                        best.append((s_, int(u)))
                        best = sorted(best, reverse=True)[:ef]
        return [i for _, i in sorted(best, reverse=True)[:k]]
    hits = 0
    for _ in range(20):
        q = rng.standard_normal(16).astype(np.float32)
        hits += len(set(search(q)) & set(np.argsort(-(X @ q))[:5].tolist()))
    assert hits >= 90                                    # recall@5 >= 0.9 over 20 random queries


def test_incremental_dense_builder_keeps_fp32_rows_and_the_container_type(tmp_path, monkeypatch):
    """IncrementalDenseBuilder re-writes the index from the fp32 rows of the FILE plus the new fp32 vectors -- not from the
    bf16 device matrix (round-1 finding) -- and keeps an HNSW container an HNSW container.  Host logic only: the device
    index is faked."""
    from types import SimpleNamespace
    from legal_rag_b200.retrieval import builders
    X0 = np.random.default_rng(5).standard_normal((6, 8)).astype(np.float32)
    idx, meta = tmp_path / "faiss.index", tmp_path / "meta.jsonl"
    artifacts.write_faiss_hnsw_flat(idx, X0, M=2)
    chunks = [LawChunk(id=f"c{i}", law_name="L", article_no=str(i), article_id=f"a{i}", text=f"t{i}") for i in range(8)]
    artifacts.write_meta_jsonl(meta, chunks[:6])
    added = []
    new_vecs = np.random.default_rng(6).standard_normal((2, 8)).astype(np.float32)
    store = SimpleNamespace(chunks=list(chunks[:6]), index_path=idx, meta_path=meta, load=lambda: None,
                            _embed=lambda texts, is_query=False: new_vecs, index=SimpleNamespace(add=lambda v: added.append(np.array(v))),
                            _index_mtime=None, _meta_mtime=None)
    monkeypatch.setattr(builders, "_knn_links", lambda X, m: None)
    cfg = SimpleNamespace(retrieval=SimpleNamespace(faiss_index_file=str(idx), hnsw_m=2, hnsw_ef_construction=40, hnsw_ef_search=16))
    b = builders.IncrementalDenseBuilder(cfg, store=store)
    assert b.add_chunks(chunks) == 2 and len(added) == 1
    Y, info = artifacts.read_faiss_index(idx, with_graph=True)
    np.testing.assert_array_equal(Y, np.concatenate([X0, new_vecs]))            # bit-identical fp32 rows, old and new
    assert info["outer_fourcc"] == "IHNf" and len(info["hnsw"]["levels"]) == 8
    assert len(artifacts.read_meta_jsonl(meta)) == 8
    assert b.add_chunks(chunks) == 0                                             # nothing new: no rewrite


def test_meta_jsonl_roundtrip(tmp_path):
    cs = _chunks(5)
    artifacts.write_meta_jsonl(tmp_path / "m.jsonl", cs)
    assert [c.id for c in artifacts.read_meta_jsonl(tmp_path / "m.jsonl")] == [c.id for c in cs]
    artifacts.write_colbert_meta(tmp_path / "c.jsonl", cs)
    back = artifacts.read_colbert_meta(tmp_path / "c.jsonl")
    assert sorted(back) == list(range(5)) and back[3].id == "c3"


# ---------------------------------------------------------------- bm25.pkl
CORPUS = [["a", "b", "b", "c"], ["a", "c"], ["a", "d", "d", "d"], ["e", "b"], ["f"]]


def test_okapi_state_matches_the_oracle_transcription():
    st = artifacts.okapi_state_from_tokens(CORPUS)
    lit = obm25.BM25Okapi(CORPUS)
    assert st.idf == lit.idf and st.avgdl == lit.avgdl and st.doc_len == lit.doc_len
    assert st.doc_freqs == lit.doc_freqs and st.average_idf == lit.average_idf and st.corpus_size == 5


def test_bm25_pickle_is_written_by_reference_and_read_with_the_shim(tmp_path):
    st = artifacts.okapi_state_from_tokens(CORPUS)
    p = tmp_path / "bm25.pkl"
    artifacts.write_bm25_pickle(p, st, _chunks(5))
    raw = p.read_bytes()
    assert b"rank_bm25" in raw and b"BM25Okapi" in raw       # what pickle.load resolves in the reference
    with pytest.raises((ModuleNotFoundError, ImportError, AttributeError)):
        pickle.loads(raw)                                     # the library is not installed here ...
    obj = artifacts.read_bm25_pickle(p)                       # ... the shim reader does not need it
    assert set(obj) == {"bm25", "chunks"} and isinstance(obj["chunks"][0], dict)
    assert obj["bm25"].idf == st.idf and obj["bm25"].doc_freqs == st.doc_freqs and obj["bm25"].k1 == 1.5


def test_host_index_from_okapi_equals_from_tokens_and_oracle():
    st = artifacts.okapi_state_from_tokens(CORPUS)
    a = Bm25HostIndex.from_okapi(st)
    b = Bm25HostIndex.from_tokens(CORPUS)
    lit = obm25.BM25Okapi(CORPUS)
    for w, i in a.term_to_id.items():
        j = b.term_to_id[w]
        assert a.idf[i] == pytest.approx(b.idf[j], rel=1e-12) == pytest.approx(lit.idf[w], rel=1e-12)
        np.testing.assert_array_equal(a.doc_id[a.indptr[i]:a.indptr[i + 1]], b.doc_id[b.indptr[j]:b.indptr[j + 1]])
    # per-posting impacts sum to get_scores
    imp = a.impacts()
    for q in (["a"], ["b", "b", "zzz", "a"], ["d", "f"]):
        sc = np.zeros(5)
        for w in q:
            i = a.term_to_id.get(w)
            if i is not None:
                sl = slice(a.indptr[i], a.indptr[i + 1])
                np.add.at(sc, a.doc_id[sl], imp[sl].astype(np.float64))
        np.testing.assert_allclose(sc, lit.get_scores(q), rtol=1e-6, atol=1e-7)
    qi, qt, mx = a.encode_queries([["a", "zzz", "a"], [], ["f"]])
    assert qi.tolist() == [0, 3, 3, 4] and qt.tolist() == [a.term_to_id["a"], -1, a.term_to_id["a"], a.term_to_id["f"]] and mx == 3


def test_query_tokenizer_keeps_case_and_whitespace_tokens():
    tok = encoders.default_query_tokenizer()
    out = tok("Buyer's remedies, Article 2")
    assert "Buyer's" in out or "Buyer" in out          # jieba (if present) may split the apostrophe
    assert any(t.isspace() for t in out)               # whitespace tokens never match the index
    assert encoders.tokenize_en("Buyer's remedies, Article 2") == ["buyer's", "remedies", "article", "2"]


# ---------------------------------------------------------------- hybrid host logic
def test_dedup_keep_best_matches_reference_golden(golden_dir):
    from legal_rag_b200.retrieval.hybrid_retriever import _dedup_keep_best
    with open(os.path.join(golden_dir, "fuse_golden.json")) as f:
        want = json.load(f)["helpers"]["dedup_E8"]
    c1, c2 = [LawChunk(id=f"d{i}", law_name="L", article_no=str(i), article_id=str(i), text="t") for i in (1, 2)]
    hs = [RetrievalHit(chunk=c1, score=0.4, score_breakdown={"channel": ["dense"], "channel_contrib": {"dense": .4}}),
          RetrievalHit(chunk=c1, score=0.7, source="graph", score_breakdown={"channel": "graph"}),
          RetrievalHit(chunk=c2, score=0.5, score_breakdown={"channel": ["bm25"]})]
    got = _dedup_keep_best(hs)
    assert [(h.chunk.id, h.score, h.rank, h.source) for h in got] == [(w["id"], w["score"], w["rank"], w["source"]) for w in want]
    assert sorted(got[0].score_breakdown["channel"]) == want[0]["channel"]
    assert got[0].score_breakdown["channel_contrib"] == want[0]["channel_contrib"]


def test_hybrid_search_flow_with_fake_channels(caplog):
    """HybridRetriever's host orchestration with every store faked (no GPU): per-channel labelling, oversampling depth,
    min_final_score filter, graph augmentation (seeds + graph hits), reranker splice, dedup, top_k slice, timing log
    (hybrid_retriever.py:172-384 of the reference)."""
    import logging
    from types import SimpleNamespace
    from legal_rag_b200.retrieval.hybrid_retriever import HybridRetriever
    chunks = [LawChunk(id=f"c{i}", law_name="L", article_no=str(i), article_id=f"a{i}", text=f"text {i}") for i in range(8)]
    asked = {}

    class Dense:
        def search(self, q, k):
            asked["dense"] = k
            return [RetrievalHit(chunk=chunks[i], score=s, rank=9, source="graph", semantic_score=s, score_breakdown={"junk": 1})
                    for i, s in ((2, 0.5), (0, 0.9), (1, 0.7))]                     # unsorted on purpose

    class Bm25:
        def search(self, q, k):
            asked["bm25"] = k
            return [(chunks[1], 7.0), (chunks[3], 9.0)]

    class Colbert:
        def search(self, q, k):
            return [(chunks[4], 21.0), RetrievalHit(chunk=chunks[0], score=25.0, score_breakdown={"channel": "colbert", "colbert_raw": 1.0})]

    class Graph:
        def search(self, q, seeds, decision=None, top_k=10):
            asked["graph_seeds"] = [h.chunk.id for h in seeds]
            return [RetrievalHit(chunk=chunks[6], score=0.3, source="graph", score_breakdown={"channel": "graph"}),
                    RetrievalHit(chunk=chunks[0], score=0.2, source="graph", score_breakdown={"channel": "graph"})]

    rcfg = SimpleNamespace(top_k=5, min_final_score=0.25, enable_graph=True, graph_seed_k=3, enable_rerank=True, rerank_top_n=2)
    hr = object.__new__(HybridRetriever)
    hr.cfg = SimpleNamespace(retrieval=rcfg)
    hr.dense, hr.bm25, hr.colbert, hr.graph = Dense(), Bm25(), Colbert(), Graph()

    d = hr.search_dense("q", 3)
    assert [(h.chunk.id, h.rank, h.source) for h in d] == [("c0", 1, "retriever"), ("c1", 2, "retriever"), ("c2", 3, "retriever")]
    assert d[0].score_breakdown == {"channel": ["dense"], "dense_raw": 0.9}            # replaces what the hit carried
    b = hr.search_bm25("q", 0)
    assert asked["bm25"] == 1 and [(h.chunk.id, h.rank) for h in b] == [("c3", 1), ("c1", 2)]
    assert b[0].score_breakdown == {"channel": ["bm25"], "bm25_raw": 9.0}
    c = hr.search_colbert("q", 4)
    assert [h.chunk.id for h in c] == ["c0", "c4"]
    assert c[0].score_breakdown == {"channel": ["colbert"], "colbert_raw": 1.0}       # a carried breakdown is kept, the channel listed
    assert c[1].score_breakdown == {"channel": ["colbert"], "colbert_raw": 21.0}
    hr.colbert = SimpleNamespace(search=lambda q, k: 1 / 0)
    assert hr.search_colbert("q", 4) == []                                             # channel failures are swallowed
    hr.colbert = Colbert()
    hr.search_graph("q", 2)                                                            # no seeds given: channel heads, dense first
    assert asked["graph_seeds"] == ["c0", "c1", "c2", "c3", "c1", "c0", "c4"]

    def fake_fuse(*, dense_hits, bm25_hits, colbert_hits):                             # stands in for the fusion kernel
        sc = {}
        for w, hs in ((0.5, dense_hits), (0.03, bm25_hits), (0.01, colbert_hits)):
            for h in hs:
                sc[h.chunk.id] = sc.get(h.chunk.id, 0.0) + w * float(h.score)
        ranked = sorted(sc.items(), key=lambda kv: -kv[1])
        return [RetrievalHit(chunk=chunks[int(cid[1:])], score=v, rank=r, score_breakdown={"channel": ["dense"]})
                for r, (cid, v) in enumerate(ranked, start=1)]
    hr._fuse = fake_fuse
    hr.reranker = lambda q, hits: [h.model_copy(update={"score": 10.0 - j, "source": "rerank"}) for j, h in enumerate(reversed(hits))]
    with caplog.at_level(logging.INFO, logger="legal_rag_b200.retrieval"):
        out = hr.search("q", None, 4, SimpleNamespace(mode="RoutingMode.GRAPH_AUGMENTED"))
    assert asked["dense"] == 5                                                         # max(top_k, rcfg.top_k) oversampling
    # fused: c0 .45+.25=.70, c1 .35+.21=.56, c3 .27, c2 .25, c4 .21 -> filter >= .25 keeps c0 c1 c3 c2; seeds = first 3
    assert asked["graph_seeds"] == ["c0", "c1", "c3"]
    # seeds + graph hits (c6 .3, c0 .2), reranker reverses the first two (c1 -> 10, c0 -> 9), dedup drops the graph copy of c0
    assert [(h.chunk.id, h.rank) for h in out] == [("c1", 1), ("c0", 2), ("c6", 3), ("c3", 4)]
    assert out[0].source == "rerank" and sorted(out[1].score_breakdown["channel"]) == ["dense", "graph"]
    assert any("[retrieval] dense=" in r.getMessage() and "rerank=" in r.getMessage() for r in caplog.records)
    # without the routing decision there is no graph stage
    asked.pop("graph_seeds")
    hr.search("q", None, 2)
    assert "graph_seeds" not in asked


def test_error_contracts_without_touching_the_gpu(tmp_path):
    from legal_rag_b200.config import AppConfig
    from legal_rag_b200.retrieval import BM25Retriever, VectorStore
    cfg = AppConfig()
    cfg.retrieval.faiss_index_file = str(tmp_path / "nope.index")
    cfg.retrieval.faiss_meta_file = str(tmp_path / "nope.jsonl")
    cfg.retrieval.bm25_index_file = str(tmp_path / "nope.pkl")
    with pytest.raises(FileNotFoundError):                    # tests/test_retrieval.py:114-122 of the reference
        VectorStore(cfg).load()
    with pytest.raises(RuntimeError, match="index not found"):
        BM25Retriever(cfg).load()
    with open(tmp_path / "nope.pkl", "wb") as f:
        pickle.dump({"chunks": []}, f)
    with pytest.raises(RuntimeError, match="missing 'bm25'"):
        BM25Retriever(cfg).load()


def test_shard_ranges_cover_the_corpus():
    from legal_rag_b200.engine import shard_range
    for n, w in [(100, 8), (7, 8), (10_000_000, 3), (0, 2)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_incremental_bm25_builder_dedups_rebuilds_and_replaces_atomically(tmp_path):
    """builders/incremental_bm25_builder.py:43-82: ids already indexed are skipped, the Okapi statistics are
    re-derived over existing + new chunks, the pickle is replaced in one step and stays readable by the shim."""
    from legal_rag_b200.config import AppConfig
    from legal_rag_b200.retrieval import builders
    cfg = AppConfig()
    cfg.retrieval.bm25_index_file = str(tmp_path / "idx" / "bm25.pkl")
    first = _chunks(5)
    builders.build_bm25_index(cfg, first, tokenizer=encoders.tokenize_en)
    inc = builders.IncrementalBM25Builder(cfg, tokenizer=encoders.tokenize_en)
    jl = tmp_path / "new.jsonl"
    more = _chunks(8)                       # c0..c4 are duplicates, c5..c7 are new
    jl.write_text("\n".join(json.dumps(c.model_dump()) for c in more) + "\n\n", encoding="utf-8")
    assert inc.add_jsonl(jl) == 3
    assert inc.add_jsonl(jl) == 0           # nothing new the second time
    payload = artifacts.read_bm25_pickle(cfg.retrieval.bm25_index_file)
    assert [c["id"] for c in payload["chunks"]] == [f"c{i}" for i in range(8)]
    ref = obm25.BM25Okapi([encoders.tokenize_en(c.text) for c in more])
    got = payload["bm25"]
    assert got.corpus_size == 8 and abs(got.avgdl - ref.avgdl) < 1e-12
    assert got.idf == pytest.approx(ref.idf)
    assert not os.path.exists(cfg.retrieval.bm25_index_file + ".tmp")
    with pytest.raises(FileNotFoundError):
        inc.add_jsonl(tmp_path / "missing.jsonl")
    # an unreadable pickle counts as an empty index (incremental_bm25_builder.py:31-41)
    with open(cfg.retrieval.bm25_index_file, "wb") as f:
        f.write(b"not a pickle")
    assert inc.add_chunks(_chunks(2)) == 2


def test_plaid_directory_round_trip(tmp_path):
    """retrieval/plaid.py: compress with the indexer's recipe, read back, compare with an independent decompression and with
    the original vectors (4-bit residuals around 64 centroids keep the cosine above 0.8)."""
    import torch
    from legal_rag_b200.retrieval import plaid
    rng = np.random.default_rng(3)
    docs = []
    for n in (5, 17, 1, 32, 9, 26):
        v = rng.standard_normal((n, 128)).astype(np.float32)
        docs.append(v / np.linalg.norm(v, axis=1, keepdims=True))
    plaid.write_plaid_index(tmp_path / "idx", docs, n_centroids=64, nbits=4, chunk_docs=4)
    assert plaid.is_plaid_dir(tmp_path / "idx") and not plaid.is_plaid_dir(tmp_path)
    tokens, doclen = plaid.read_plaid_index(tmp_path / "idx")
    assert tokens.shape == (6, 32, 128) and doclen.tolist() == [5, 17, 1, 32, 9, 26]
    # independent decompression from the files: explicit bit twiddling instead of the byte table
    cent = torch.load(tmp_path / "idx" / "centroids.pt").float().numpy()
    weights = torch.load(tmp_path / "idx" / "buckets.pt")[1].numpy()
    flat = []
    for c in range(2):
        codes = torch.load(tmp_path / "idx" / f"{c}.codes.pt").numpy()
        packed = torch.load(tmp_path / "idx" / f"{c}.residuals.pt").numpy()
        bits = np.unpackbits(packed, axis=1).reshape(len(codes), 128, 4)            # bit 0 of each index comes first
        idx = (bits * (1 << np.arange(4))).sum(axis=2)
        e = cent[codes] + weights[idx]
        flat.append(e / np.linalg.norm(e, axis=1, keepdims=True))
    flat = np.concatenate(flat)
    got = np.concatenate([tokens[i, :n].float().numpy() for i, n in enumerate(doclen.tolist())])
    np.testing.assert_allclose(got, flat, atol=1e-2)                               # bf16 storage
    orig = np.concatenate(docs)
    assert ((got * orig).sum(axis=1) > 0.8).all()
    assert (tokens[1, 17:] == 0).all()                                              # padding rows are zero
    with pytest.raises(ValueError, match="embeddings"):
        (tmp_path / "idx" / "doclens.0.json").write_text("[1, 2]")
        plaid.read_plaid_index(tmp_path / "idx")


def test_transformers_bge_encoder_matches_flagmodel_semantics(tmp_path):
    """encoders.TransformersBgeEncoder on a tiny random-init BERT saved to disk (no network): [CLS] pooling, unit norm, the
    retrieval instruction on queries only -- the behaviour the reference configures FlagModel for (vector_store.py:66-77)."""
    import torch
    from transformers import BertConfig, BertModel, BertTokenizerFast
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + list("abcdefghijklmnopqrstuvwxyz") + ["buyer", "seller", "goods", "为", "法", "律"]
    (tmp_path / "m").mkdir()
    (tmp_path / "m" / "vocab.txt").write_text("\\n".join(vocab), encoding="utf-8")
    BertTokenizerFast(vocab_file=str(tmp_path / "m" / "vocab.txt")).save_pretrained(tmp_path / "m")
    torch.manual_seed(0)
    BertModel(BertConfig(vocab_size=len(vocab), hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                         max_position_embeddings=64)).save_pretrained(tmp_path / "m")
    enc = encoders.make_dense_encoder(str(tmp_path / "m"), "cpu")
    assert isinstance(enc, encoders.TransformersBgeEncoder) and enc.dim == 32
    texts = ["buyer seller goods", "goods"]
    v = enc.encode(texts)
    assert v.shape == (2, 32) and v.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-5)
    tok = enc.tokenizer(texts, padding=True, truncation=True, max_length=512, return_tensors="pt")
    with torch.no_grad():
        ref = torch.nn.functional.normalize(enc.model(**tok).last_hidden_state[:, 0], dim=-1).numpy()
    np.testing.assert_allclose(v, ref, atol=1e-6)
    q = enc.encode_queries(["goods"])
    np.testing.assert_allclose(q, enc.encode([encoders.BGE_QUERY_INSTRUCTION + "goods"]), atol=1e-6)
    assert not np.allclose(q[0], v[1], atol=1e-3)                       # the instruction changes the query vector
    assert enc.encode("goods").shape == (32,)                            # a bare string gives one vector (graph_retriever.py:177)


def test_transformers_colbert_encoder_restates_the_checkpoint_pipeline(tmp_path):
    """encoders.TransformersColbertEncoder on a tiny random-init ColBERT-style checkpoint saved to disk (no network): ". " prefix,
    [unused0] / [unused1] markers, [MASK] augmentation to query_maxlen with the mask tokens not attended to, punctuation and
    padding dropped from documents, unit-norm rows -- what colbert's Checkpoint does with the directory the reference passes
    to Searcher / Indexer (colbert_retriever.py:135-136, builders/colbert_builder.py:123-132)."""
    import json
    import torch
    from safetensors.torch import load_file, save_file
    from transformers import BertConfig, BertModel, BertTokenizerFast
    vocab = (["[PAD]", "[unused0]", "[unused1]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", ".", ",", "?", "!"]
             + ["buyer", "seller", "goods", "contract", "sale", "of", "the", "a"])
    d = tmp_path / "ckpt"
    d.mkdir()
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, processors
    wp = Tokenizer(models.WordPiece({w: i for i, w in enumerate(vocab)}, unk_token="[UNK]"))
    wp.normalizer = normalizers.BertNormalizer(lowercase=True)
    wp.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    wp.post_processor = processors.BertProcessing(("[SEP]", vocab.index("[SEP]")), ("[CLS]", vocab.index("[CLS]")))
    BertTokenizerFast(tokenizer_object=wp, unk_token="[UNK]", pad_token="[PAD]", cls_token="[CLS]", sep_token="[SEP]",
                      mask_token="[MASK]").save_pretrained(d)
    torch.manual_seed(1)
    bert = BertModel(BertConfig(vocab_size=len(vocab), hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                                max_position_embeddings=64))
    bert.eval()
    bert.save_pretrained(d)
    # an HF_ColBERT checkpoint keeps `linear.weight` next to the bert.* tensors
    sd = {"bert." + k: v.contiguous() for k, v in load_file(d / "model.safetensors").items()}
    W = torch.randn(16, 32)
    sd["linear.weight"] = W
    save_file(sd, d / "model.safetensors", metadata={"format": "pt"})
    (d / "artifact.metadata").write_text(json.dumps({"query_maxlen": 8, "doc_maxlen": 12, "dim": 16}))

    enc = encoders.make_token_encoder(str(d), "cpu")
    assert isinstance(enc, encoders.TransformersColbertEncoder) and enc.dim == 16 and enc.query_maxlen == 8
    tok = enc.tokenizer
    # ---- query: 8 rows whatever the text length, all unit vectors ----
    q = enc.encode_query("sale of goods")
    assert q.shape == (8, 16) and q.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    ids = tok([". sale of goods"], padding="max_length", truncation=True, max_length=8, return_tensors="pt")
    ii, mm = ids["input_ids"].clone(), ids["attention_mask"]
    assert ii[0, 1].item() == tok.convert_tokens_to_ids(".")          # the ". " prefix is what the marker overwrites
    ii[:, 1] = tok.convert_tokens_to_ids("[unused0]")
    ii[ii == tok.pad_token_id] = tok.mask_token_id
    with torch.no_grad():
        ref = torch.nn.functional.normalize(bert(input_ids=ii, attention_mask=mm).last_hidden_state @ W.t(), dim=2)[0].numpy()
    np.testing.assert_allclose(q, ref, atol=1e-5)
    # ---- document: punctuation and padding rows are gone, the rest are unit vectors in order ----
    doc = "the buyer, the seller!"
    dv = enc.encode_doc(doc)
    ids = tok([". " + doc], return_tensors="pt")["input_ids"]
    punct = {tok.convert_tokens_to_ids(x) for x in [".", ",", "?", "!"]}
    ids[:, 1] = tok.convert_tokens_to_ids("[unused1]")
    keep = [i for i, t in enumerate(ids[0].tolist()) if t not in punct]
    assert dv.shape == (len(keep), 16) and len(keep) == ids.shape[1] - 2         # "," and "!" dropped; [CLS] [D] [SEP] kept
    with torch.no_grad():
        ref = torch.nn.functional.normalize(bert(input_ids=ids).last_hidden_state @ W.t(), dim=2)[0].numpy()[keep]
    np.testing.assert_allclose(dv, ref, atol=1e-5)
    # a batch pads to the longest document and reports the lengths
    D, n = enc.encode_docs_device([doc, "goods"])
    assert D.shape[0] == 2 and int(n[0]) == len(keep) and int(n[1]) == 4 and float(D[1, 4:].abs().sum()) == 0.0
    # not a directory, nothing registered: the channel cannot be built, and says why
    with pytest.raises(RuntimeError, match="not a local checkpoint directory"):
        encoders.make_token_encoder("colbert-ir/colbertv2.0", "cpu")


def test_bm25_query_compaction_drops_unknown_tokens_before_the_term_limit():
    """BM25Retriever.search_ids: out-of-vocabulary tokens (-1) leave the query before the kernel's 128-slot term table is
    filled; what is left is cut to the first `cap` tokens.  Checked against a per-query loop, with empty and all-unknown
    queries, on random ragged batches."""
    from legal_rag_b200.retrieval.bm25_retriever import compact_known_terms
    rng = np.random.default_rng(3)
    for trial in range(50):
        nq = int(rng.integers(0, 9))
        lists = []
        for _ in range(nq):
            n = int(rng.integers(0, 40)) if rng.random() < 0.8 else int(rng.integers(100, 400))
            t = rng.integers(0, 1000, n).astype(np.int32)
            t[rng.random(n) < (1.0 if rng.random() < 0.1 else 0.4)] = -1
            lists.append(t)
        qi = np.concatenate([[0], np.cumsum([len(t) for t in lists])]).astype(np.int64)
        qt = np.concatenate(lists).astype(np.int32) if lists else np.zeros(0, np.int32)
        cap = int(rng.choice([4, 16, 128]))
        gi, gt, mx, longest = compact_known_terms(qi, qt, cap)
        want = [t[t >= 0] for t in lists]
        assert longest == max((len(w) for w in want), default=0)
        want = [w[:cap] for w in want]
        assert mx == max((len(w) for w in want), default=0)
        assert gi.dtype == np.int64 and gt.dtype == np.int32 and gi.tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()
        assert gt.tolist() == (np.concatenate(want).tolist() if want else [])
