"""On-disk index artifacts of the hot path, read and written without faiss / rank_bm25 installed.

  faiss.index        written by faiss.write_index (builders/faiss_builder.py:96): an IndexHNSWFlat ("IHNf")
                     wrapping an IndexFlat ("IxFI" inner product / "IxF2" L2), or a bare IndexFlat.  The
                     engine is exact, so only the flat storage is used; the HNSW graph is skipped.
  faiss_meta.jsonl   one LawChunk per line, row i <-> line i (builders/faiss_builder.py:101-103).
  bm25.pkl           pickle of {"bm25": rank_bm25.BM25Okapi, "chunks": [dict]} (builders/bm25_builder.py:46-51).
  colbert_meta.jsonl {"pid": int, "chunk": {...}} per line (builders/colbert_builder.py:39-52).

The faiss layouts follow faiss/impl/index_write.cpp / index_read.cpp (faiss-cpu >= 1.7.4; 1.13.2 is what the
reference's notebooks resolved).  They could not be checked against a file written by a real faiss in this
environment (the wheel is not installable here): round-trips of this module's own writer are tested, and the
HNSW reader falls back to locating the nested flat index by its fourcc + header if the graph section differs.
"""
from __future__ import annotations

import io
import json
import os
import pickle
import struct
import sys
import types
from pathlib import Path
from typing import Optional, Any, Dict, List, Tuple

import numpy as np

from ..schemas import LawChunk, chunk_from_obj

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


# --------------------------------------------------------------------------------------------------
# faiss index files
# --------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.p = buf, 0

    def take(self, fmt: str):
        sz = struct.calcsize(fmt)
        if self.p + sz > len(self.b):
            raise ValueError("faiss index: unexpected end of file")
        v = struct.unpack_from(fmt, self.b, self.p)
        self.p += sz
        return v if len(v) > 1 else v[0]

    def vector(self, dtype, scale: int = 1) -> np.ndarray:
        n = self.take("<Q") * scale
        dt = np.dtype(dtype)
        nbytes = n * dt.itemsize
        if self.p + nbytes > len(self.b):
            raise ValueError("faiss index: vector runs past the end of the file")
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.p)
        self.p += nbytes
        return a


def _read_header(r: _Reader) -> Dict[str, Any]:
    d = r.take("<i")
    ntotal = r.take("<q")
    r.take("<q"); r.take("<q")            # two dummies
    is_trained = r.take("<B")
    metric = r.take("<i")
    metric_arg = r.take("<f") if metric > 1 else 0.0
    return {"d": d, "ntotal": ntotal, "is_trained": bool(is_trained), "metric": metric, "metric_arg": metric_arg}


def _read_flat(r: _Reader) -> Tuple[np.ndarray, Dict[str, Any]]:
    fourcc = r.b[r.p:r.p + 4]
    r.p += 4
    if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
        raise ValueError(f"faiss index: nested storage {fourcc!r} is not a flat index")
    h = _read_header(r)
    xb = r.vector(np.float32)               # size field counts floats (READXBVECTOR)
    if xb.size != h["ntotal"] * h["d"]:
        raise ValueError(f"faiss index: {xb.size} floats stored, header says {h['ntotal']} x {h['d']}")
    h["fourcc"] = fourcc.decode()
    return xb.reshape(h["ntotal"], h["d"]), h


def read_faiss_index(path, with_graph: bool = False) -> Tuple[np.ndarray, Dict[str, Any]]:
    """-> (float32 [ntotal, d] vectors in id order, info dict with d / ntotal / metric / fourcc / hnsw).  `with_graph`
    keeps the HNSW arrays (levels, offsets, neighbors, ...) in info["hnsw"] instead of only its scalars."""
    buf = Path(path).read_bytes()
    r = _Reader(buf)
    fourcc = buf[:4]
    if fourcc in (b"IxFI", b"IxF2", b"IxFl"):
        X, h = _read_flat(r)
        h["hnsw"] = None
        return X, h
    if fourcc in (b"IHNf", b"IHN2"):
        r.p = 4
        outer = _read_header(r)
        try:
            hn = {"assign_probas": r.vector(np.float64), "cum_nneighbor_per_level": r.vector(np.int32),
                  "levels": r.vector(np.int32), "offsets": r.vector(np.uint64), "neighbors": r.vector(np.int32)}
            hn["entry_point"], hn["max_level"], hn["efConstruction"], hn["efSearch"] = r.take("<iiii")
            save = r.p
            if buf[r.p:r.p + 2] != b"Ix":       # older writers also stored upper_beam
                r.take("<i")
            if buf[r.p:r.p + 2] != b"Ix":
                r.p = save
                raise ValueError("nested index not where expected")
            X, h = _read_flat(r)
        except ValueError:
            X, h = _locate_nested_flat(buf, outer)
            hn = {}
        h["hnsw"] = {k: (v if (with_graph or np.isscalar(v)) else None) for k, v in hn.items()}
        h["outer_fourcc"] = fourcc.decode()
        return X, h
    raise ValueError(f"faiss index {path}: unsupported index type {fourcc!r} (flat and HNSW-flat files are supported)")


def _locate_nested_flat(buf: bytes, outer: Dict[str, Any]):
    for tag in (b"IxFI", b"IxF2"):
        pos = buf.find(tag, 4)
        while pos >= 0:
            r = _Reader(buf)
            r.p = pos
            try:
                X, h = _read_flat(r)
                if h["d"] == outer["d"] and h["ntotal"] == outer["ntotal"]:
                    return X, h
            except ValueError:
                pass
            pos = buf.find(tag, pos + 1)
    raise ValueError("faiss index: no nested flat storage found")


def write_faiss_flat(path, X: np.ndarray, metric: int = METRIC_INNER_PRODUCT) -> None:
    """IndexFlatIP / IndexFlatL2 file (what faiss.write_index(faiss.IndexFlatIP(d)) produces)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = X.shape
    tmp = str(path) + ".tmp"
    with open(tmp, "wb") as f:
        f.write(b"IxFI" if metric == METRIC_INNER_PRODUCT else b"IxF2")
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric))
        f.write(struct.pack("<Q", n * d))
        f.write(X.tobytes())
    os.replace(tmp, path)


def knn_graph(X: np.ndarray, m: int, block: int = 2048) -> np.ndarray:
    """Exact inner-product m-nearest-neighbour lists [n, m] (self excluded, -1 padded) on the host: the level-0 links of
    the HNSW container below for corpora small enough for numpy (builders.build_faiss_index computes them with the dense
    scan kernel instead when a GPU is there)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n = X.shape[0]
    out = np.full((n, m), -1, dtype=np.int32)
    kk = min(m, max(n - 1, 0))
    if kk == 0:
        return out
    for lo in range(0, n, block):
        S = X[lo:lo + block] @ X.T
        S[np.arange(S.shape[0]), np.arange(lo, lo + S.shape[0])] = -np.inf          # no self links
        idx = np.argpartition(-S, kk - 1, axis=1)[:, :kk]
        order = np.argsort(-np.take_along_axis(S, idx, 1), axis=1, kind="stable")
        out[lo:lo + S.shape[0], :kk] = np.take_along_axis(idx, order, 1)
    return out


def write_faiss_hnsw_flat(path, X: np.ndarray, M: int = 64, ef_construction: int = 400, ef_search: int = 512,
                          neighbors: Optional[np.ndarray] = None) -> None:
    """IndexHNSWFlat file (builders/faiss_builder.py:84-96 writes this type) [upstream faiss write_index: fourcc IHNf, index
    header, HNSW {assign_probas, cum_nneighbor_per_level, levels, offsets, neighbors, entry_point, max_level, efConstruction,
    efSearch, upper_beam}, nested IndexFlat].  The graph is a single-level HNSW: every vector sits on level 0 only
    (levels[i] = 1), owns 2 M neighbour slots there and links to its exact inner-product nearest neighbours (`neighbors`
    [n, <= 2 M] int32, -1 = empty; computed on the host when not given), entry point 0.  That is a graph a faiss search can
    walk (greedy descent is empty, level-0 beam search runs over exact kNN links); this engine itself ignores the graph and
    scans the flat storage exactly.  The layout follows the published format and has not been opened by a real faiss here
    (none is installable): parity unpinned."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = X.shape
    slots = 2 * M
    if neighbors is None:
        neighbors = knn_graph(X, slots) if n <= 50_000 else np.full((n, 0), -1, dtype=np.int32)
    nb = np.full((n, slots), -1, dtype=np.int32)
    w = min(slots, neighbors.shape[1])
    nb[:, :w] = neighbors[:, :w]
    # faiss HNSW::set_default_probas(M, 1 / ln M): level l is drawn with probability e^(-l / mult) (1 - e^(-1 / mult))
    mult = 1.0 / np.log(max(M, 2))
    probas, cum = [], [0]
    level = 0
    while True:
        pr = float(np.exp(-level / mult) * (1 - np.exp(-1 / mult)))
        if pr < 1e-9:
            break
        probas.append(pr)
        cum.append(cum[-1] + (slots if level == 0 else M))
        level += 1
    tmp = str(path) + ".tmp"
    with open(tmp, "wb") as f:
        f.write(b"IHNf")
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, METRIC_INNER_PRODUCT))
        for dtype, vals in ((np.float64, probas), (np.int32, cum), (np.int32, np.ones(n, dtype=np.int32)),
                            (np.uint64, np.arange(n + 1, dtype=np.uint64) * np.uint64(slots)), (np.int32, nb.reshape(-1))):
            a = np.asarray(vals, dtype=dtype)
            f.write(struct.pack("<Q", a.size))
            f.write(a.tobytes())
        f.write(struct.pack("<iiiii", 0 if n else -1, 0 if n else -1, ef_construction, ef_search, 1))
        f.write(b"IxFI")
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, METRIC_INNER_PRODUCT))
        f.write(struct.pack("<Q", n * d))
        f.write(X.tobytes())
    os.replace(tmp, path)


# --------------------------------------------------------------------------------------------------
# meta JSONL
# --------------------------------------------------------------------------------------------------
def read_meta_jsonl(path) -> List[LawChunk]:
    chunks: List[LawChunk] = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                chunks.append(LawChunk.model_validate(json.loads(line)))
    return chunks


def write_meta_jsonl(path, chunks) -> None:
    tmp = str(path) + ".tmp"
    with open(tmp, "w", encoding="utf-8") as f:
        for c in chunks:
            f.write(json.dumps(c.model_dump(), ensure_ascii=False) + "\n")
    os.replace(tmp, path)


def read_colbert_meta(path) -> Dict[int, LawChunk]:
    out: Dict[int, LawChunk] = {}
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if line:
                rec = json.loads(line)
                out[int(rec["pid"])] = LawChunk.model_validate(rec["chunk"])
    return out


def write_colbert_meta(path, chunks) -> None:
    tmp = str(path) + ".tmp"
    with open(tmp, "w", encoding="utf-8") as f:
        for pid, c in enumerate(chunks):
            f.write(json.dumps({"pid": pid, "chunk": c.model_dump()}, ensure_ascii=False) + "\n")
    os.replace(tmp, path)


# --------------------------------------------------------------------------------------------------
# bm25.pkl
# --------------------------------------------------------------------------------------------------
class OkapiState:
    """Stand-in for rank_bm25.BM25Okapi when unpickling without the library: pickle restores the
    instance __dict__ (k1, b, epsilon, corpus_size, avgdl, doc_freqs, idf, doc_len, average_idf,
    tokenizer) onto it."""

    def __init__(self, **state):
        self.__dict__.update(state)


class _ShimUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "rank_bm25":
            return OkapiState
        if module.startswith("legalrag.") and name == "LawChunk":
            return LawChunk
        return super().find_class(module, name)


def read_bm25_pickle(path) -> Dict[str, Any]:
    with open(path, "rb") as f:
        data = f.read()
    try:
        import rank_bm25  # noqa: F401  the real class, when installed
        return pickle.loads(data)
    except ImportError:
        return _ShimUnpickler(io.BytesIO(data)).load()


def write_bm25_pickle(path, okapi_state, chunks) -> None:
    """{"bm25": <BM25Okapi>, "chunks": [dict]} with the BM25 object pickled BY REFERENCE to
    rank_bm25.BM25Okapi, so the reference (with the real library) can load it; written atomically like
    builders/incremental_bm25_builder.py:76-79."""
    try:
        import rank_bm25
        cls = rank_bm25.BM25Okapi
        obj = cls.__new__(cls)
        obj.__dict__.update(okapi_state.__dict__)
        fake = None
    except ImportError:
        fake = types.ModuleType("rank_bm25")
        cls = type("BM25Okapi", (), {"__module__": "rank_bm25"})
        fake.BM25Okapi = cls
        obj = cls.__new__(cls)
        obj.__dict__.update(okapi_state.__dict__)
    payload = {"bm25": obj, "chunks": [c.model_dump() if not isinstance(c, dict) else c for c in chunks]}
    tmp = str(path) + ".tmp"
    try:
        if fake is not None:
            sys.modules["rank_bm25"] = fake
        with open(tmp, "wb") as f:
            pickle.dump(payload, f)
    finally:
        if fake is not None:
            sys.modules.pop("rank_bm25", None)
    os.replace(tmp, path)


def okapi_state_from_tokens(corpus_tokens, k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25) -> OkapiState:
    """The attribute set BM25Okapi(corpus_tokens) ends up with (rank_bm25 0.2.2 __init__ / _calc_idf),
    computed here so that an index can be written without the library."""
    import math
    from collections import Counter
    doc_freqs, doc_len, nd, num = [], [], {}, 0
    for doc in corpus_tokens:
        doc_len.append(len(doc))
        num += len(doc)
        fr = dict(Counter(doc))
        doc_freqs.append(fr)
        for w in fr:
            nd[w] = nd.get(w, 0) + 1
    n = len(doc_len)
    idf, idf_sum, neg = {}, 0.0, []
    for w, f in nd.items():
        v = math.log(n - f + 0.5) - math.log(f + 0.5)
        idf[w] = v
        idf_sum += v
        if v < 0:
            neg.append(w)
    average_idf = idf_sum / len(idf) if idf else 0.0
    for w in neg:
        idf[w] = epsilon * average_idf
    return OkapiState(k1=k1, b=b, epsilon=epsilon, corpus_size=n, avgdl=(num / n if n else 0.0), doc_freqs=doc_freqs, idf=idf,
                      doc_len=doc_len, tokenizer=None, average_idf=average_idf)
