"""CPU tests: the oracle against the reference-derived golden vectors and hand-computed cases."""
import json
import math
import os

import numpy as np
import pytest

from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import fuse as ofuse
from oracle import maxsim as omaxsim


@pytest.fixture(scope="module")
def fuse_golden(golden_dir):
    with open(os.path.join(golden_dir, "fuse_golden.json")) as f:
        return json.load(f)


def _tied_runs(rows):
    runs, a = [], 0
    while a < len(rows):
        b = a + 1
        while b < len(rows) and rows[b]["score"] == rows[a]["score"]:
            b += 1
        runs.append((a, b))
        a = b
    return runs


def test_fuse_restatement_matches_reference_execution(fuse_golden):
    n = 0
    for name, case in fuse_golden["cases"].items():
        inp = case["inputs"]
        for run in case["runs"]:
            got = ofuse.fuse([tuple(p) for p in inp["dense"]], [tuple(p) for p in inp["bm25"]],
                             [tuple(p) for p in inp["colbert"]], method=run["method"],
                             w_dense=run["w_dense"], w_bm25=run["w_bm25"], w_colbert=run["w_colbert"],
                             rrf_k=run["rrf_k"], alpha=run["alpha"])
            ref = run["out"]
            assert len(got) == len(ref), (name, run["method"])
            by_id = {r["id"]: r for r in got}
            for r in ref:
                g = by_id[r["id"]]
                for key in ("score", "rrf_norm", "weighted_sum", "dense_norm", "bm25_norm", "colbert_norm"):
                    assert g[key] == pytest.approx(r[key], rel=1e-12, abs=1e-15), (name, run["method"], r["id"], key)
                for ch, v in r["channel_contrib"].items():
                    assert g["channel_contrib"][ch] == pytest.approx(v, rel=1e-12, abs=1e-15)
            # order: identical up to exact ties (the reference's tie order depends on PYTHONHASHSEED)
            for a, b in _tied_runs(ref):
                assert {r["id"] for r in ref[a:b]} == {r["id"] for r in got[a:b]}, (name, run["method"], a, b)
            n += 1
    assert n >= 40


def test_fuse_known_answers_from_survey_appendix(fuse_golden):
    a1 = {r["method"]: r["out"] for r in fuse_golden["cases"]["A1"]["runs"]}
    assert a1["rrf_norm_blend"][0]["score"] == pytest.approx(1.175)      # notebook's 1.18
    assert a1["weighted_sum"][0]["score"] == pytest.approx(1.35)
    assert a1["rrf"][1]["score"] == pytest.approx(0.484249, abs=1e-6)
    e1 = {r["method"]: r["out"] for r in fuse_golden["cases"]["E1"]["runs"]}
    assert [r["id"] for r in e1["weighted_sum"][:3]] == [2, 1, 4]
    assert e1["rrf_norm_blend"][0]["score"] == pytest.approx(0.85)
    h = fuse_golden["helpers"]
    assert h["rrf_E6"]["b"] == pytest.approx(0.032522475, abs=1e-9)
    assert h["minmax"][0][1] == [1.0, 0.0, 0.5] and h["minmax"][1][1] == [0.0, 0.0] and h["minmax"][2][1] == []
    assert ofuse.minmax([3, 1, 2]) == [1.0, 0.0, 0.5] and ofuse.minmax([1, 1 + 1e-13]) == [0.0, 0.0]


def test_bm25_hand_computed_three_doc_corpus():
    # doc0: a b b c ; doc1: a c ; doc2: a d d d   -> 'a' is in every doc (negative raw idf)
    corpus = [["a", "b", "b", "c"], ["a", "c"], ["a", "d", "d", "d"]]
    bm = obm25.BM25Okapi(corpus)
    N, avgdl = 3, 10 / 3
    raw = {w: math.log(N - n + 0.5) - math.log(n + 0.5) for w, n in {"a": 3, "b": 1, "c": 2, "d": 1}.items()}
    avg = sum(raw.values()) / 4
    assert raw["a"] < 0 and raw["c"] < 0
    assert bm.idf["a"] == pytest.approx(0.25 * avg) and bm.idf["c"] == pytest.approx(0.25 * avg)
    assert bm.idf["b"] == pytest.approx(raw["b"])

    def term(w, f, dl):
        return bm.idf[w] * f * 2.5 / (f + 1.5 * (1 - 0.75 + 0.75 * dl / avgdl))
    # repeated query term counts twice; OOV 'zzz' contributes 0
    s = bm.get_scores(["b", "b", "zzz", "a"])
    assert s[0] == pytest.approx(2 * term("b", 2, 4) + term("a", 1, 4))
    assert s[1] == pytest.approx(term("a", 1, 2))
    assert s[2] == pytest.approx(term("a", 1, 4))
    # k > matching docs: zero-score docs are returned, ties to the lower index
    sc, idx = obm25.search(bm, ["b"], 3)
    assert idx.tolist() == [0, 1, 2] and sc[1] == 0.0 and sc[2] == 0.0
    sc, idx = obm25.search(bm, ["zzz"], 2)
    assert idx.tolist() == [0, 1]


def test_bm25_csr_matches_literal(golden_dir):
    z = np.load(os.path.join(golden_dir, "ucc_corpus.npz"))
    lens, flat = z["doc_len"], z["tokens"].astype(np.int64)
    assert len(lens) == 591 and len(z["vocab"]) == 3926
    off = np.concatenate([[0], np.cumsum(lens)])
    docs = [flat[off[i]:off[i + 1]] for i in range(len(lens))]
    csr = obm25.CsrBM25.from_token_ids(docs, vocab=3926)
    assert csr.indptr[-1] == 53991                      # nnz of the UCC corpus
    lit = obm25.BM25Okapi([[str(t) for t in d] for d in docs[:]])
    assert lit.avgdl == pytest.approx(csr.avgdl)
    neg = sum(1 for w, v in lit.idf.items() if v == 0.25 * lit.average_idf)
    assert neg == 26 and lit.average_idf == pytest.approx(4.7547, abs=2e-3)
    rng = np.random.default_rng(44)
    for _ in range(8):
        d = rng.integers(0, 591)
        L = rng.integers(3, 9)
        st = rng.integers(0, max(1, lens[d] - L))
        q = docs[d][st:st + L].tolist() + [docs[d][st]]          # one repeated term
        a = lit.get_scores([str(t) for t in q])
        b = csr.get_scores(q)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
        s1, i1 = obm25.search(lit, [str(t) for t in q], 100)
        s2, i2 = csr.search(q, 100)
        assert i1.tolist() == i2.tolist()


def test_dense_oracle_edges():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((50, 16)).astype(np.float32)
    X[7] = X[3]                                          # duplicate vector: tie -> lower id first
    Q = rng.standard_normal((3, 16)).astype(np.float32)
    D, I = odense.flat_ip_topk(Q, X, 60)                 # k > N pads (-FLT_MAX, -1)
    assert (I[:, 50:] == -1).all() and (D[:, 50:] == odense.FLT_LOWEST).all()
    S = Q @ X.T
    for q in range(3):
        assert I[q, :50].tolist() == np.lexsort((np.arange(50), -S[q].astype(np.float64))).tolist()
        p3, p7 = I[q].tolist().index(3), I[q].tolist().index(7)
        assert p7 == p3 + 1
    D2, I2 = odense.flat_ip_topk(Q, X, 5, chunk=8)       # chunked == unchunked
    assert I2.tolist() == I[:, :5].tolist()
    m_s, m_i = odense.merge_topk(np.concatenate([D[:, :5], D[:, 5:10]], 1), np.concatenate([I[:, :5], I[:, 5:10]], 1), 5)
    assert m_i.tolist() == I[:, :5].tolist()


def test_maxsim_oracle_masking_and_skip():
    rng = np.random.default_rng(1)
    D = rng.standard_normal((6, 5, 8)).astype(np.float32)
    Q = rng.standard_normal((2, 3, 8)).astype(np.float32)
    doclen = np.array([5, 2, 5, 1, 5, 3])
    cand = np.array([[0, 1, -1, 3], [5, 4, 2, -1]])
    S = omaxsim.maxsim_scores(Q, D, doclen, cand)
    assert np.isneginf(S[0, 2]) and np.isneginf(S[1, 3])
    ref = sum(max(float(Q[0, i] @ D[1, j]) for j in range(2)) for i in range(3))
    assert S[0, 1] == pytest.approx(ref, rel=1e-6)
    s, i = omaxsim.rerank_topk(Q, D, doclen, cand, 4)
    assert i[0, 3] == -1 and set(i[0, :3].tolist()) == {0, 1, 3}


def test_graph_scoring_restatement_matches_reference_execution(golden_dir):
    """oracle/graph.py against tests/golden/graph_golden.json (the reference's GraphRetriever.search executed)."""
    from oracle import graph as ograph
    g = json.load(open(os.path.join(golden_dir, "graph_golden.json")))
    for dpt, gamma, want in g["functions"]["depth_decay"]:
        assert ograph.depth_decay(dpt, gamma) == pytest.approx(want, rel=1e-12)
    for rels, want in g["functions"]["relation_weight"]:
        assert ograph.relation_weight(rels) == want
    vecs, q = np.array(g["vectors"], dtype=np.float32), np.array(g["qvec"], dtype=np.float32)
    chunks = g["chunks"]
    id2row = {c["article_id"]: i for i, c in enumerate(chunks)}
    for case in g["cases"]:
        got = ograph.graph_hits(q, g["nodes"], id2row, vecs, [c["text"] for c in chunks], [c.get("lang") for c in chunks],
                                lang=case["lang"], top_k=case["top_k"], gamma=case["gamma"])
        assert [chunks[h["row"]]["id"] for h in got] == [h["id"] for h in case["hits"]]
        for a, b in zip(got, case["hits"]):
            assert a["score"] == pytest.approx(b["score"], rel=1e-6, abs=1e-9) and a["rank"] == b["rank"]
            for key in ("semantic", "depth_decay", "relation_weight", "edge_conf", "graph_depth"):
                assert a[key] == pytest.approx(b["breakdown"][key], rel=1e-6, abs=1e-9)


def test_bm25_oracle_reproduces_the_upstream_readme_example():
    """rank_bm25's own README (the library the reference calls at bm25_retriever.py:74, pinned >= 0.2.2) documents
    `bm25.get_scores("windy London".split(" "))` over its three-sentence corpus as array([0., 0.93729472, 0.]) and
    `get_top_n(..., n=1)` as the London sentence: the one published known answer for this boundary."""
    corpus = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
    lit = obm25.BM25Okapi([d.split(" ") for d in corpus])
    scores = lit.get_scores("windy London".split(" "))
    np.testing.assert_allclose(scores, [0.0, 0.93729472, 0.0], atol=5e-9)
    s, i = obm25.search(lit, "windy London".split(" "), 1)
    assert i.tolist() == [1]
