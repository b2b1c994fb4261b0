"""Seeded synthetic inputs shaped like BASELINE.json's configs, generated on the device so that the
10^7..10^8-row corpora never cross PCIe.  (The reference's own seed convention is 42,
scripts/generate_synthetic_data.py:36; per-config seeds follow SURVEY.md section 8d.)
"""
from __future__ import annotations

import torch


def unit_rows_bf16(n: int, d: int, seed: int, device, chunk: int = 1 << 20) -> torch.Tensor:
    """[n, d] bf16, rows L2-normalised in fp32 before rounding (BGE embeddings are unit norm,
    legalrag/retrieval/vector_store.py:154)."""
    out = torch.empty((n, d), dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, d), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        out[s:e] = x.to(torch.bfloat16)
    return out


def unit_tokens_bf16(n: int, L: int, d: int, seed: int, device, chunk: int = 1 << 14) -> torch.Tensor:
    out = torch.empty((n, L, d), dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, L, d), generator=g, device=device, dtype=torch.float32)
        out[s:e] = torch.nn.functional.normalize(x, dim=2).to(torch.bfloat16)
    return out
