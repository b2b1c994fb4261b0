/* A plain-C caller of liblrag's C ABI (include/lrag.h): no Python, no torch -- device memory from the CUDA runtime,
 * device pointers and sizes into the library, results checked against a scalar loop.  It is what a non-Python host
 * (or a ctypes / cffi stub on the reference side) does.
 *
 *   gcc tests/c_abi/smoke.c -Iinclude -I/usr/local/cuda/include -Llegal_rag_b200 -llrag -L/usr/local/cuda/lib64 -lcudart -lm
 *
 * Exercises: lrag_init, lrag_dense_topk_bf16 (+ workspace sizing), lrag_bm25_topk (+ workspace sizing), lrag_fuse_topk.
 * Exit code 0 = every check passed.  Compiled without a GPU by tests/test_abi.py (link check); run by the GPU tests. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lrag.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "cuda: %s (%s)\n", cudaGetErrorString(e_), #x); return 2; } } while (0)
#define LK(x) do { int r_ = (x); if (r_ != LRAG_OK) { fprintf(stderr, "lrag: %d %s (%s)\n", r_, lrag_last_error(), #x); return 3; } } while (0)

static uint16_t to_bf16(float f) {               /* round to nearest even */
  uint32_t u; memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float from_bf16(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint32_t rng_state = 12345u;
static float rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return (float)((rng_state >> 8) & 0xffff) / 65536.0f - 0.5f; }

int main(void) {
  enum { N = 5000, D = 64, NQ = 4, K = 5, V = 50 };
  LK(lrag_init(0));
  if (lrag_sm_count() <= 0) { fprintf(stderr, "no SMs?\n"); return 4; }

  /* ---------------- dense: top-K of Q . X^T ---------------- */
  uint16_t* X = malloc(sizeof(uint16_t) * N * D); uint16_t* Q = malloc(sizeof(uint16_t) * NQ * D);
  for (int i = 0; i < N * D; ++i) X[i] = to_bf16(rnd());
  for (int i = 0; i < NQ * D; ++i) Q[i] = to_bf16(rnd());
  void *dX, *dQ, *dws; float* ds; int64_t* di;
  CK(cudaMalloc(&dX, sizeof(uint16_t) * N * D)); CK(cudaMalloc(&dQ, sizeof(uint16_t) * NQ * D));
  CK(cudaMalloc((void**)&ds, sizeof(float) * NQ * K)); CK(cudaMalloc((void**)&di, sizeof(int64_t) * NQ * K));
  CK(cudaMemcpy(dX, X, sizeof(uint16_t) * N * D, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dQ, Q, sizeof(uint16_t) * NQ * D, cudaMemcpyHostToDevice));
  size_t wsb = lrag_dense_topk_workspace_bytes(N, D, NQ, K);
  CK(cudaMalloc(&dws, wsb));
  LK(lrag_dense_topk_bf16(dX, N, D, dQ, NQ, K, /*id_base=*/1000, ds, di, dws, wsb, /*stream=*/NULL));
  CK(cudaDeviceSynchronize());
  float hs[NQ * K]; int64_t hi[NQ * K];
  CK(cudaMemcpy(hs, ds, sizeof(hs), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hi, di, sizeof(hi), cudaMemcpyDeviceToHost));
  for (int q = 0; q < NQ; ++q) {
    /* scalar reference: best score and its row */
    double best = -1e30; int arg = -1;
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int j = 0; j < D; ++j) acc += (double)from_bf16(Q[q * D + j]) * from_bf16(X[n * D + j]);
      if (acc > best) { best = acc; arg = n; }
    }
    if (hi[q * K] != 1000 + arg || fabs(hs[q * K] - best) > 1e-3 * fabs(best) + 1e-5) {
      fprintf(stderr, "dense q=%d: got (%lld, %g), want (%d, %g)\n", q, (long long)hi[q * K], hs[q * K], 1000 + arg, best); return 5;
    }
    for (int r = 1; r < K; ++r) if (hs[q * K + r] > hs[q * K + r - 1]) { fprintf(stderr, "dense q=%d not sorted\n", q); return 6; }
  }

  /* ---------------- BM25: term-major CSR postings with precomputed impacts ---------------- */
  /* doc n holds term (n % V) with impact 1 + (n % 7) and term ((n / V) % V) with impact 0.5 (merged when equal) */
  int64_t* indptr = calloc(V + 1, sizeof(int64_t));
  int* cntv = calloc(V, sizeof(int));
  for (int n = 0; n < N; ++n) { int a = n % V, b = (n / V) % V; cntv[a]++; if (b != a) cntv[b]++; }
  for (int t = 0; t < V; ++t) indptr[t + 1] = indptr[t] + cntv[t];
  int64_t nnz = indptr[V];
  int32_t* pdoc = malloc(sizeof(int32_t) * nnz); float* pimp = malloc(sizeof(float) * nnz);
  memset(cntv, 0, sizeof(int) * V);
  for (int n = 0; n < N; ++n) {              /* ascending doc ids inside every term */
    int a = n % V, b = (n / V) % V;
    float ia = 1.0f + (float)(n % 7) + (a == b ? 0.5f : 0.0f);
    pdoc[indptr[a] + cntv[a]] = n; pimp[indptr[a] + cntv[a]++] = ia;
    if (b != a) { pdoc[indptr[b] + cntv[b]] = n; pimp[indptr[b] + cntv[b]++] = 0.5f; }
  }
  int64_t q_indptr[NQ + 1]; int32_t q_term[NQ * 2];
  for (int q = 0; q < NQ; ++q) { q_indptr[q] = 2 * q; q_term[2 * q] = 3 + q; q_term[2 * q + 1] = 7 * q % V; }
  q_indptr[NQ] = 2 * NQ;
  void *dip, *dpd, *dpi, *dqi, *dqt, *dws2; float* bs; int64_t* bi;
  CK(cudaMalloc(&dip, sizeof(int64_t) * (V + 1))); CK(cudaMalloc(&dpd, sizeof(int32_t) * nnz)); CK(cudaMalloc(&dpi, sizeof(float) * nnz));
  CK(cudaMalloc(&dqi, sizeof(q_indptr))); CK(cudaMalloc(&dqt, sizeof(q_term)));
  CK(cudaMalloc((void**)&bs, sizeof(float) * NQ * K)); CK(cudaMalloc((void**)&bi, sizeof(int64_t) * NQ * K));
  CK(cudaMemcpy(dip, indptr, sizeof(int64_t) * (V + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpd, pdoc, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dpi, pimp, sizeof(float) * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dqi, q_indptr, sizeof(q_indptr), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dqt, q_term, sizeof(q_term), cudaMemcpyHostToDevice));
  size_t wsb2 = lrag_bm25_topk_workspace_bytes(N, NQ, K, 2);
  CK(cudaMalloc(&dws2, wsb2));
  LK(lrag_bm25_topk(dip, dpd, dpi, V, nnz, dqi, dqt, NQ, 2, N, K, /*id_base=*/1000, /*nonneg=*/1, /*impact_bound=*/8.0f, bs, bi, dws2, wsb2, NULL));
  CK(cudaDeviceSynchronize());
  float hbs[NQ * K]; int64_t hbi[NQ * K];
  CK(cudaMemcpy(hbs, bs, sizeof(hbs), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hbi, bi, sizeof(hbi), cudaMemcpyDeviceToHost));
  for (int q = 0; q < NQ; ++q) {
    float best = -1.0f; int arg = -1;
    for (int n = 0; n < N; ++n) {
      float sc = 0.0f;
      for (int j = 0; j < 2; ++j) {
        int t = q_term[2 * q + j];
        for (int64_t p = indptr[t]; p < indptr[t + 1]; ++p) if (pdoc[p] == n) sc += pimp[p];
      }
      if (sc > best) { best = sc; arg = n; }                               /* ties: lower id wins */
    }
    if (hbi[q * K] != 1000 + arg || fabsf(hbs[q * K] - best) > 1e-4f * best) {
      fprintf(stderr, "bm25 q=%d: got (%lld, %g), want (%d, %g)\n", q, (long long)hbi[q * K], hbs[q * K], 1000 + arg, best); return 7;
    }
  }

  /* ---------------- fusion of the two lists (weighted_sum 0.6 / 0.4): smoke = runs, sorted, ids from the inputs ------ */
  float* fs; int64_t* fi;
  CK(cudaMalloc((void**)&fs, sizeof(float) * NQ * K)); CK(cudaMalloc((void**)&fi, sizeof(int64_t) * NQ * K));
  LK(lrag_fuse_topk(ds, di, bs, bi, NULL, NULL, NQ, K, K, /*method=*/0, 0.6, 0.4, 0.0, 60, 0.5, 0.0, fs, fi, NULL, NULL));
  CK(cudaDeviceSynchronize());
  float hfs[NQ * K]; int64_t hfi[NQ * K];
  CK(cudaMemcpy(hfs, fs, sizeof(hfs), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hfi, fi, sizeof(hfi), cudaMemcpyDeviceToHost));
  for (int q = 0; q < NQ; ++q)
    for (int r = 0; r < K; ++r) {
      if (r && hfs[q * K + r] > hfs[q * K + r - 1]) { fprintf(stderr, "fuse q=%d not sorted\n", q); return 8; }
      int found = 0;
      for (int j = 0; j < K; ++j) found |= (hfi[q * K + r] == hi[q * K + j]) || (hfi[q * K + r] == hbi[q * K + j]);
      if (!found) { fprintf(stderr, "fuse q=%d rank %d: id %lld is in neither list\n", q, r, (long long)hfi[q * K + r]); return 9; }
    }
  printf("c abi smoke ok: dense, bm25 and fusion through plain C (%lld kernel launches)\n", lrag_launch_count());
  return 0;
}
