"""Text -> vectors / tokens.  The encoders (BGE for the dense channel, a ColBERT checkpoint for late
interaction, jieba for BM25 queries) are NOT part of the hot path this package accelerates (SURVEY 8f ranks
them "next"); they are pluggable here.  The library-backed defaults are used when the reference's
dependencies are installed; tests and benchmarks plug in deterministic stand-ins."""
from __future__ import annotations

import hashlib
import re
from typing import Callable, List, Optional, Sequence

import numpy as np

_warned_no_jieba = False
_TOKEN_RE = re.compile(r"[A-Za-z0-9]+(?:'[A-Za-z0-9]+)?|[一-鿿]|\s+|[^\sA-Za-z0-9一-鿿]")


def tokenize_en(text: str) -> List[str]:
    """The reference's English INDEX tokenizer (builders/bm25_builder.py:18-19)."""
    return re.findall(r"[A-Za-z0-9]+(?:'[A-Za-z0-9]+)?", text.lower())


def default_query_tokenizer() -> Callable[[str], List[str]]:
    """bm25_retriever.py:73 tokenises QUERIES with jieba.cut for both languages.  Without jieba a regex
    splitter keeps its observable behaviour on ASCII text (case preserved; words, whitespace runs and
    punctuation become separate tokens) and falls back to single CJK characters."""
    try:
        import jieba
        return lambda q: list(jieba.cut(q))
    except ImportError:
        global _warned_no_jieba
        if not _warned_no_jieba:
            _warned_no_jieba = True
            import logging
            logging.getLogger(__name__).warning(
                "jieba is not installed: BM25 queries are split by a regex (words / whitespace runs / punctuation / single CJK "
                "characters); Chinese queries tokenise differently from the reference (bm25_retriever.py:73)")
        return lambda q: _TOKEN_RE.findall(q)


class HashingDenseEncoder:
    """Deterministic stand-in for BGE (tests / offline use): signed feature hashing of word tokens into
    `dim` buckets, L2-normalised float32 like FlagModel output (vector_store.py:131-155)."""

    def __init__(self, dim: int = 768):
        self.dim = dim

    def _one(self, text: str) -> np.ndarray:
        v = np.zeros(self.dim, dtype=np.float32)
        for tok in tokenize_en(text) or [text]:
            h = int.from_bytes(hashlib.blake2b(tok.encode("utf-8"), digest_size=8).digest(), "little")
            v[h % self.dim] += 1.0 if (h >> 40) & 1 else -1.0
        n = float(np.linalg.norm(v))
        return v / n if n > 0 else v

    def encode(self, texts: Sequence[str], **_) -> np.ndarray:
        return np.stack([self._one(t) for t in texts]) if len(texts) else np.zeros((0, self.dim), np.float32)

    encode_queries = encode


class HashingTokenEncoder:
    """Deterministic stand-in for a ColBERT checkpoint: one unit 128-d vector per word token."""

    def __init__(self, dim: int = 128, query_maxlen: int = 32, doc_maxlen: int = 220):
        self.dim, self.query_maxlen, self.doc_maxlen = dim, query_maxlen, doc_maxlen

    def _tok(self, tok: str) -> np.ndarray:
        seed = int.from_bytes(hashlib.blake2b(tok.encode("utf-8"), digest_size=8).digest(), "little")
        v = np.random.default_rng(seed).standard_normal(self.dim).astype(np.float32)
        return v / np.linalg.norm(v)

    def encode_query(self, text: str) -> np.ndarray:
        toks = tokenize_en(text)[: self.query_maxlen] or ["[empty]"]
        return np.stack([self._tok(t) for t in toks])

    def encode_doc(self, text: str) -> np.ndarray:
        toks = tokenize_en(text)[: self.doc_maxlen] or ["[empty]"]
        return np.stack([self._tok(t) for t in toks])


BGE_QUERY_INSTRUCTION = "为这个法律问题生成表示以用于检索相关法律条文："        # vector_store.py:72 (used for both languages)


class TransformersBgeEncoder:
    """BGE-style dense encoder on the same GPU as the index, for installs that have the checkpoint on
    disk and `transformers` but not FlagEmbedding.  Same observable behaviour as FlagModel as the reference configures it
    (vector_store.py:66-77, 131-155): [CLS] pooling, L2-normalised float32 vectors, the retrieval instruction prepended
    to queries only, max_length 512, fp16 weights on a GPU.  `encode_device` / `encode_queries_device` return the vectors
    as a device tensor, so a query goes from text to the scan kernel without a host round trip.  The transformer forward
    itself is library code (HF transformers); nothing here is on the accelerated path."""

    def __init__(self, model_path: str, device="cpu", query_instruction: str = BGE_QUERY_INSTRUCTION, use_fp16: Optional[bool] = None):
        import torch
        from transformers import AutoModel, AutoTokenizer
        self.device = torch.device(device)
        self.query_instruction = query_instruction
        self.tokenizer = AutoTokenizer.from_pretrained(model_path, local_files_only=True)
        model = AutoModel.from_pretrained(model_path, local_files_only=True)
        if use_fp16 is None:
            use_fp16 = self.device.type == "cuda"
        self.model = (model.half() if use_fp16 else model.float()).to(self.device).eval()
        self.dim = int(self.model.config.hidden_size)

    def encode_device(self, texts: Sequence[str], batch_size: int = 64, max_length: int = 512, **_):
        import torch
        out = []
        with torch.no_grad():
            for i in range(0, len(texts), batch_size):
                batch = self.tokenizer(list(texts[i:i + batch_size]), padding=True, truncation=True, max_length=max_length,
                                       return_tensors="pt").to(self.device)
                cls = self.model(**batch).last_hidden_state[:, 0]
                out.append(torch.nn.functional.normalize(cls.float(), dim=-1))
        return torch.cat(out) if out else torch.zeros((0, self.dim), device=self.device)

    def encode_queries_device(self, queries: Sequence[str], **kw):
        return self.encode_device([self.query_instruction + q for q in queries], **kw)

    def encode(self, texts, **kw) -> np.ndarray:
        single = isinstance(texts, str)
        v = self.encode_device([texts] if single else list(texts), **kw).cpu().numpy()
        return v[0] if single else v

    def encode_queries(self, queries, **kw) -> np.ndarray:
        single = isinstance(queries, str)
        v = self.encode_queries_device([queries] if single else list(queries), **kw).cpu().numpy()
        return v[0] if single else v


class TransformersColbertEncoder:
    """ColBERT checkpoint directory -> unit 128-d token vectors on the index's device, for installs that
    have the checkpoint on disk and `transformers` but not colbert-ai.  It restates what the reference's Searcher / Indexer do
    with the checkpoint (colbert_retriever.py:119-152, builders/colbert_builder.py:109-134) [upstream colbert-ai 0.2.x,
    colbert/modeling/{hf_colbert,checkpoint,tokenization}]: BERT encoder + a bias-free linear layer hidden -> dim; every
    text is prefixed with ". " and the token after [CLS] is overwritten by the marker [unused0] (queries) / [unused1]
    (documents); queries are padded to `query_maxlen` with [MASK] tokens that are not attended to but whose output vectors
    are kept (query augmentation); document tokens that are padding or punctuation are dropped (mask_punctuation); every
    kept vector is L2-normalised.  The transformer forward is library code (HF transformers); the vectors go straight to
    the MaxSim kernels as device tensors (`encode_queries_device`)."""

    def __init__(self, model_path: str, device="cpu", query_maxlen: Optional[int] = None, doc_maxlen: Optional[int] = None,
                 use_fp16: Optional[bool] = None):
        import json
        import os
        import string
        import torch
        from transformers import AutoModel, AutoTokenizer
        self.device = torch.device(device)
        meta = {}
        meta_path = os.path.join(model_path, "artifact.metadata")           # written by colbert's checkpoint saver
        if os.path.exists(meta_path):
            with open(meta_path) as f:
                meta = json.load(f)
        self.query_maxlen = int(query_maxlen or meta.get("query_maxlen", 32))
        self.doc_maxlen = int(doc_maxlen or meta.get("doc_maxlen", 220))
        self.attend_to_mask_tokens = bool(meta.get("attend_to_mask_tokens", False))
        self.mask_punctuation = bool(meta.get("mask_punctuation", True))
        self.tokenizer = AutoTokenizer.from_pretrained(model_path, local_files_only=True)
        bert = AutoModel.from_pretrained(model_path, local_files_only=True)
        weight = self._linear_weight(model_path)
        if weight.shape[1] != bert.config.hidden_size:
            raise RuntimeError(f"ColBERT checkpoint {model_path}: linear.weight {tuple(weight.shape)} does not match hidden size "
                               f"{bert.config.hidden_size}")
        if use_fp16 is None:
            use_fp16 = self.device.type == "cuda"
        self.model = (bert.half() if use_fp16 else bert.float()).to(self.device).eval()
        self.linear = weight.to(self.device).to(torch.float16 if use_fp16 else torch.float32)     # [dim, hidden], no bias
        self.dim = int(weight.shape[0])
        tok = self.tokenizer
        self.q_marker = tok.convert_tokens_to_ids("[unused0]")
        self.d_marker = tok.convert_tokens_to_ids("[unused1]")
        self.skip_ids = set()
        if self.mask_punctuation:
            for sym in string.punctuation:
                ids = tok.encode(sym, add_special_tokens=False)
                if ids:
                    self.skip_ids.add(int(ids[0]))

    @staticmethod
    def _linear_weight(model_path: str):
        """`linear.weight` of the checkpoint (HF_ColBERT keeps it next to the `bert.*` tensors)."""
        import os
        import torch
        st = os.path.join(model_path, "model.safetensors")
        if os.path.exists(st):
            from safetensors import safe_open
            with safe_open(st, framework="pt") as f:
                for name in ("linear.weight", "colbert.linear.weight"):
                    if name in f.keys():
                        return f.get_tensor(name).float()
        pt = os.path.join(model_path, "pytorch_model.bin")
        if os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu", weights_only=True)
            for name in ("linear.weight", "colbert.linear.weight"):
                if name in sd:
                    return sd[name].float()
        raise RuntimeError(f"ColBERT checkpoint {model_path}: no linear.weight tensor (not a ColBERT checkpoint directory?)")

    def _forward(self, ids, mask):
        import torch
        with torch.no_grad():
            h = self.model(input_ids=ids, attention_mask=mask).last_hidden_state
            return h @ self.linear.t()

    def encode_queries_device(self, queries: Sequence[str]):
        """-> float32 [nq, query_maxlen, dim] on the device; every row is a unit vector ([MASK] rows included)."""
        import torch
        tok = self.tokenizer
        obj = tok([". " + q for q in queries], padding="max_length", truncation=True, max_length=self.query_maxlen, return_tensors="pt")
        ids, mask = obj["input_ids"], obj["attention_mask"]
        ids[:, 1] = self.q_marker
        ids[ids == tok.pad_token_id] = tok.mask_token_id
        if self.attend_to_mask_tokens:
            mask[ids == tok.mask_token_id] = 1
        Q = self._forward(ids.to(self.device), mask.to(self.device)).float()
        return torch.nn.functional.normalize(Q, p=2, dim=2)

    def encode_docs_device(self, docs: Sequence[str], batch_size: int = 32):
        """-> (float32 [n, Lmax, dim] on the device, zero rows past each doc's length; int32 [n] lengths).  Kept tokens are
        packed to the front in their original order, like the Indexer's `D[mask]`."""
        import torch
        tok = self.tokenizer
        mats, lens = [], []
        for i in range(0, len(docs), batch_size):
            obj = tok([". " + d for d in docs[i:i + batch_size]], padding="longest", truncation="longest_first", max_length=self.doc_maxlen,
                      return_tensors="pt")
            ids, mask = obj["input_ids"], obj["attention_mask"]
            ids[:, 1] = self.d_marker
            keep = torch.tensor([[(int(x) not in self.skip_ids) and (int(x) != tok.pad_token_id) for x in row] for row in ids.tolist()])
            D = self._forward(ids.to(self.device), mask.to(self.device)).float()
            D = torch.nn.functional.normalize(D * keep.to(self.device).unsqueeze(2), p=2, dim=2)
            for r in range(D.shape[0]):
                rows = D[r][keep[r].to(self.device)]
                mats.append(rows)
                lens.append(rows.shape[0])
        Lmax = max(lens) if lens else 0
        out = torch.zeros((len(mats), Lmax, self.dim), dtype=torch.float32, device=self.device)
        for i, m in enumerate(mats):
            out[i, : m.shape[0]] = m
        return out, torch.tensor(lens, dtype=torch.int32, device=self.device)

    # the numpy interface the token-store builder and ColBERTRetriever use
    def encode_query(self, text: str) -> np.ndarray:
        return self.encode_queries_device([text])[0].cpu().numpy()

    def encode_doc(self, text: str) -> np.ndarray:
        D, n = self.encode_docs_device([text])
        return D[0, : int(n[0])].cpu().numpy()


_dense_encoder_factory: Optional[Callable] = None
_token_encoder_factory: Optional[Callable] = None


def register_dense_encoder(factory: Optional[Callable]) -> None:
    """factory(model_name, device) -> object with encode(texts) / encode_queries(texts) -> float32 [n, d]."""
    global _dense_encoder_factory
    _dense_encoder_factory = factory


def register_token_encoder(factory: Optional[Callable]) -> None:
    """factory(model_name, device) -> object with encode_query(text) -> [Lq, 128], encode_doc(text) -> [Ld, 128]."""
    global _token_encoder_factory
    _token_encoder_factory = factory


def make_dense_encoder(model_name: str, device):
    if _dense_encoder_factory is not None:
        return _dense_encoder_factory(model_name, device)
    try:
        from FlagEmbedding import FlagModel
    except ImportError as e:
        import os
        if os.path.isdir(str(model_name)):
            # a checkpoint directory on disk and no FlagEmbedding: HF transformers on the index's own device
            return TransformersBgeEncoder(str(model_name), device)
        raise RuntimeError(
            "no dense encoder: FlagEmbedding is not installed, the embedding model is not a local checkpoint directory and "
            "no encoder was registered (legal_rag_b200.retrieval.encoders.register_dense_encoder)") from e
    import torch
    return FlagModel(model_name, query_instruction_for_retrieval=BGE_QUERY_INSTRUCTION,
                     use_fp16=torch.cuda.is_available(), device=device)      # vector_store.py:70-75


def make_token_encoder(model_name: str, device, query_maxlen: Optional[int] = None, doc_maxlen: Optional[int] = None):
    """A registered factory wins; else a ColBERT checkpoint DIRECTORY is loaded through HF transformers on the index's
    device (the reference hands the same `colbert_model_name` to colbert's Searcher, colbert_retriever.py:135-136)."""
    if _token_encoder_factory is not None:
        return _token_encoder_factory(model_name, device)
    import os
    if os.path.isdir(str(model_name)):
        return TransformersColbertEncoder(str(model_name), device, query_maxlen=query_maxlen, doc_maxlen=doc_maxlen)
    raise RuntimeError(
        f"no ColBERT token encoder: {model_name!r} is not a local checkpoint directory (there is no network access to a model "
        "hub here) and no encoder was registered (legal_rag_b200.retrieval.encoders.register_token_encoder)")
