"""VectorStore + the faiss-shaped index object behind it (replaces legalrag/retrieval/vector_store.py:32-181).

The FAISS file is read once into a bf16 matrix resident in HBM; `index.search(x, k)` is the exact
flat inner-product scan of liblrag (the reference's default IndexHNSWFlat is an approximation of it)."""
from __future__ import annotations

import threading
from pathlib import Path
from typing import ClassVar, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import engine
from ..schemas import LawChunk
from . import artifacts, encoders


class GpuFlatIndex:
    """What `faiss.read_index` returns, as far as the reference uses it: `.d`, `.ntotal`,
    `.search(x, k) -> (D float32 [nq, k], I int64 [nq, k])`, `.add(x)`.  Exact, inner product."""

    def __init__(self, d: int, device=None, capacity: int = 0):
        self.d = int(d)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._X = torch.empty((max(capacity, 0), self.d), dtype=torch.bfloat16, device=self.device)
        self.ntotal = 0
        self._lock = threading.Lock()
        self.is_trained = True
        self.metric_type = artifacts.METRIC_INNER_PRODUCT

    @classmethod
    def from_numpy(cls, X: np.ndarray, device=None) -> "GpuFlatIndex":
        idx = cls(X.shape[1], device, capacity=X.shape[0])
        idx.add(X)
        return idx

    def add(self, x) -> None:
        """Append rows (incremental_dense_builder.py:62).  Readers see either the old or the new matrix:
        the grown matrix is built aside and swapped in with one assignment."""
        x = torch.from_numpy(np.array(x, dtype=np.float32, order="C", copy=True))
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"add: expected [n, {self.d}], got {tuple(x.shape)}")
        with self._lock:
            n_new = self.ntotal + x.shape[0]
            if n_new > self._X.shape[0]:
                grown = torch.empty((max(n_new, int(self._X.shape[0] * 1.5) + 1), self.d), dtype=torch.bfloat16, device=self.device)
                grown[: self.ntotal] = self._X[: self.ntotal]
                grown[self.ntotal:n_new] = x.to(self.device).to(torch.bfloat16)
                self._X = grown
            else:
                self._X[self.ntotal:n_new] = x.to(self.device).to(torch.bfloat16)
            self.ntotal = n_new

    @property
    def matrix(self) -> torch.Tensor:
        return self._X[: self.ntotal]

    def search_device(self, Q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        X = self._X[: self.ntotal]
        return engine.dense_topk(X, Q.to(self.device).to(torch.bfloat16).contiguous(), k)

    def search(self, x, k: int) -> Tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"search: expected [nq, {self.d}], got {x.shape}")
        k = int(k)
        D = np.full((x.shape[0], k), engine.PAD_SCORE, dtype=np.float32)
        I = np.full((x.shape[0], k), -1, dtype=np.int64)
        if x.shape[0] == 0 or self.ntotal == 0:
            return D, I
        kk = min(k, engine.LRAG_MAX_K)
        s, i = self.search_device(torch.from_numpy(x), kk)
        D[:, :kk] = s.cpu().numpy()
        I[:, :kk] = i.cpu().numpy()
        return D, I

    def reconstruct_n(self, i0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        return self._X[i0:i0 + n].float().cpu().numpy()


class VectorStore:
    """Same surface as the reference class: from_config singleton, load() with mtime check and
    FileNotFoundError contract, .index / .chunks / ._embed / .search / .index_path / .meta_path."""

    _instances_by_key: ClassVar[Dict[Tuple[str, str, str, str], "VectorStore"]] = {}
    _model_cache: ClassVar[Dict[Tuple[str, str], object]] = {}

    def __init__(self, cfg):
        self.cfg = cfg
        rcfg = cfg.retrieval
        self.index_path = Path(rcfg.faiss_index_file)
        self.meta_path = Path(rcfg.faiss_meta_file)
        self.device = torch.device(getattr(cfg, "device", None) or ("cuda" if torch.cuda.is_available() else "cpu"))
        self._model_name = str(rcfg.embedding_model)
        self._model = None                      # created lazily: loading an index needs no encoder
        self.index: Optional[GpuFlatIndex] = None
        self.chunks: List[LawChunk] = []
        self._index_mtime: Optional[float] = None
        self._meta_mtime: Optional[float] = None
        self._load_lock = threading.Lock()

    @property
    def model(self):
        if self._model is None:
            key = (self._model_name, str(self.device))
            if key not in self._model_cache:
                self._model_cache[key] = encoders.make_dense_encoder(self._model_name, self.device)
            self._model = self._model_cache[key]
        return self._model

    @classmethod
    def from_config(cls, cfg) -> "VectorStore":
        rcfg = cfg.retrieval
        key = (str(rcfg.embedding_model), str(rcfg.faiss_index_file), str(rcfg.faiss_meta_file),
               "cuda" if torch.cuda.is_available() else "cpu")
        inst = cls._instances_by_key.get(key)
        if inst is None:
            inst = cls._instances_by_key[key] = cls(cfg)
        return inst

    def load(self) -> None:
        if not self.index_path.exists() or not self.meta_path.exists():
            raise FileNotFoundError("FAISS index or metadata not found; run scripts.build_index first.")
        index_mtime = self.index_path.stat().st_mtime
        meta_mtime = self.meta_path.stat().st_mtime
        if self.index is not None and self.chunks and self._index_mtime == index_mtime and self._meta_mtime == meta_mtime:
            return
        with self._load_lock:
            # threads that arrived together (Starlette's pool on a cold start) load once and share the snapshot
            if self.index is not None and self.chunks and self._index_mtime == index_mtime and self._meta_mtime == meta_mtime:
                return
            X, info = artifacts.read_faiss_index(self.index_path)
            if info.get("metric", artifacts.METRIC_INNER_PRODUCT) != artifacts.METRIC_INNER_PRODUCT:
                # the reference builds inner-product indexes only (builders/faiss_builder.py:84); scanning an L2 index as
                # inner product would silently rank by the wrong measure
                raise ValueError(f"{self.index_path}: FAISS metric_type {info.get('metric')} is not METRIC_INNER_PRODUCT; "
                                 "this engine scans inner-product indexes only")
            index = GpuFlatIndex.from_numpy(X, self.device if self.device.type == "cuda" else None)
            chunks = artifacts.read_meta_jsonl(self.meta_path)
            # one snapshot, swapped in after it is complete (readers never see a half-loaded store)
            self.index, self.chunks = index, chunks
            self._index_mtime, self._meta_mtime = index_mtime, meta_mtime

    def _embed(self, texts: List[str], is_query: bool = False) -> np.ndarray:
        if not texts:
            dim = self.index.d if self.index is not None else getattr(self.model, "dim", 768)
            return np.zeros((0, dim), dtype="float32")
        if is_query:
            embs = self.model.encode_queries(texts, batch_size=64, max_length=512)
        else:
            embs = self.model.encode(texts, batch_size=64, max_length=512)
        return np.asarray(embs).astype("float32")

    def search(self, query: str, top_k: int) -> List[Tuple[LawChunk, float]]:
        self.load()
        if hasattr(self.model, "encode_queries_device") and self.index.ntotal:
            # encoder on the index's device: the query vector never visits the host
            k = min(int(top_k), engine.LRAG_MAX_K)
            s, i = self.index.search_device(self.model.encode_queries_device([query]), k)
            scores, idxs = s.cpu().numpy(), i.cpu().numpy()
        else:
            q_vec = self._embed([query], is_query=True)
            scores, idxs = self.index.search(q_vec, top_k)
        hits: List[Tuple[LawChunk, float]] = []
        for score, idx in zip(scores[0], idxs[0]):
            if idx == -1:
                continue
            hits.append((self.chunks[idx], float(score)))
        return hits
