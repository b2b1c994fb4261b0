"""ctypes binding of liblrag.so (C ABI declared in include/lrag.h).

There is no CPU fallback: if the library is missing or the device is not sm_100 every entry
point raises.  torch is imported first so that the CUDA runtime the library links against is the
one torch already loaded.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import torch  # noqa: F401  (loads libcudart before liblrag)

# LRAG_LIB_PATH: an alternative build of the same library (A/B runs of kernel variants); never a different implementation
LIB_PATH = Path(os.environ.get("LRAG_LIB_PATH") or Path(__file__).resolve().parent / "liblrag.so")

LRAG_MAX_K = 1024
LRAG_BM25_MAX_QUERY_TERMS = 128
PAD_SCORE = -3.4028234663852886e38

_c_i64, _c_int, _c_sz, _c_p, _c_f64 = C.c_int64, C.c_int, C.c_size_t, C.c_void_p, C.c_double

# name -> (restype, argtypes); mirrors include/lrag.h one to one
SIGNATURES = {
    "lrag_version": (_c_int, []),
    "lrag_init": (_c_int, [_c_int]),
    "lrag_last_error": (C.c_char_p, []),
    "lrag_sm_count": (_c_int, []),
    "lrag_prof_enable": (_c_int, [_c_int]),
    "lrag_prof_collect": (_c_int, [_c_p, _c_p, _c_int]),
    "lrag_launch_count": (C.c_longlong, []),
    "lrag_dense_topk_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_int, _c_int]),
    "lrag_dense_topk_bf16": (_c_int, [_c_p, _c_i64, _c_int, _c_p, _c_int, _c_int, _c_i64, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_dense_topk_workspace_bytes_part": (_c_sz, [_c_i64, _c_int, _c_int, _c_int, _c_int]),
    "lrag_dense_topk_bf16_part": (_c_int, [_c_p, _c_i64, _c_int, _c_p, _c_int, _c_int, _c_i64, _c_int, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_dense_topk_ref_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_int, _c_int]),
    "lrag_dense_topk_bf16_ref": (_c_int, [_c_p, _c_i64, _c_int, _c_p, _c_int, _c_int, _c_i64, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_dense_gather_scores_bf16": (_c_int, [_c_p, _c_i64, _c_int, _c_p, _c_int, _c_p, _c_int, _c_p, _c_p]),
    "lrag_topk_select_workspace_bytes": (_c_sz, [_c_int, _c_i64, _c_int]),
    "lrag_topk_select_f32": (_c_int, [_c_p, _c_i64, _c_int, _c_i64, _c_int, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_topk_merge": (_c_int, [_c_p, _c_p, _c_int, _c_int, _c_int, _c_p, _c_p, _c_p]),
    "lrag_topk_merge_shards": (_c_int, [_c_p, _c_p, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_p, _c_p, _c_p]),
    "lrag_bm25_topk_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_int, _c_i64]),
    "lrag_bm25_set_item_slabs": (_c_int, [_c_int]),
    "lrag_bm25_topk": (_c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, _c_p, _c_int, _c_i64, _c_i64, _c_int, _c_i64, _c_int,
                                C.c_float, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_bm25_topk_dense": (_c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, _c_p, _c_int, _c_i64, _c_p, _c_p, _c_int, _c_i64, _c_i64, _c_int,
                                      _c_i64, _c_int, C.c_float, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_bm25_grid": (_c_int, [_c_i64, _c_int, _c_int, _c_i64, _c_int]),
    "lrag_bm25_topk_part": (_c_int, [_c_p, _c_p, _c_p, _c_i64, _c_i64, _c_p, _c_p, _c_int, _c_i64, _c_p, _c_p, _c_int, _c_i64, _c_i64, _c_int,
                                     _c_i64, _c_int, C.c_float, _c_int, _c_p, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_sm_reserve": (_c_int, [_c_int, _c_p, C.c_ulonglong, _c_int, _c_p]),
    "lrag_maxsim_rerank_workspace_bytes": (_c_sz, [_c_int, _c_int, _c_int]),
    "lrag_maxsim_rerank_bf16": (_c_int, [_c_p, _c_p, _c_i64, _c_int, _c_int, _c_p, _c_int, _c_int, _c_p, _c_int, _c_int,
                                         _c_i64, _c_p, _c_p, _c_p, _c_sz, _c_p]),
    "lrag_maxsim_scores_bf16": (_c_int, [_c_p, _c_p, _c_i64, _c_int, _c_int, _c_p, _c_int, _c_int, _c_p, _c_int, _c_p, _c_p]),
    "lrag_maxsim_scan_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_int]),
    "lrag_maxsim_scan_scores_bf16": (_c_int, [_c_p, _c_p, _c_i64, _c_int, _c_int, _c_p, _c_int, _c_int, _c_p, _c_i64, _c_p, _c_sz, _c_p]),
    "lrag_maxsim_scan_topk_bf16": (_c_int, [_c_p, _c_p, _c_i64, _c_int, _c_int, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_p, _c_p,
                                            _c_p, _c_sz, _c_p]),
    "lrag_fuse_topk": (_c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_f64, _c_f64, _c_f64,
                                _c_int, _c_f64, _c_f64, _c_p, _c_p, _c_p, _c_p]),
}

_lib = None
_lock = threading.Lock()
_inited_device = None


class LragError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen liblrag.so and attach the prototypes (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise LragError(
                    f"{LIB_PATH} is missing: build it with `python -m legal_rag_b200.build` "
                    "(or __graft_entry__.build()); this engine has no CPU or PyTorch fallback")
            lib = C.CDLL(os.fspath(LIB_PATH))
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def last_error() -> str:
    return load().lrag_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise LragError(f"{what} failed (code {rc}): {last_error()}")


def init(device: int | None = None) -> C.CDLL:
    """Bind the library to a CUDA device (once per process: one process per GPU)."""
    global _inited_device
    lib = load()
    if not torch.cuda.is_available():
        raise LragError("no CUDA device: liblrag runs on sm_100a only and has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    if _inited_device != dev:
        with _lock:
            if _inited_device != dev:
                check(lib.lrag_init(dev), "lrag_init")
                _inited_device = dev
    return lib
