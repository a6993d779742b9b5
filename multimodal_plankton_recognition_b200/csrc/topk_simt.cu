// fp32 CUDA-core path of the k-nearest candidate search (parity mode).
// Ranking key = |g|^2 - 2 q.g  (== squared euclidean distance minus the per-query constant |q|^2),
// ascending; ties by gallery index.  Gallery is split into chunks across blockIdx.y for
// parallelism; per-chunk lists are merged by select_candidates().
#include <math_constants.h>
#include "common.cuh"

namespace plk {

constexpr int TQ = 64, TG = 64, TK = 16, TNT = 256;
constexpr int kMaxKc = 64;

static int topk_f32_chunks(int64_t nq, int64_t ng) {
  // aim for >= ~4 CTAs per SM worth of (query tile, gallery chunk) pairs, chunks of >= 4096 rows
  int64_t qtiles = ceil_div(nq, TQ);
  int64_t want = ceil_div(148 * 4, qtiles);
  int64_t maxc = ceil_div(ng, 4096);
  int64_t c = want < 1 ? 1 : want;
  if (c > maxc) c = maxc;
  if (c > 64) c = 64;
  return (int)(c < 1 ? 1 : c);
}

size_t topk_ws_f32(int64_t nq, int64_t ng, int64_t d, int kc) {
  int c = topk_f32_chunks(nq, ng);
  if (c == 1) return 0;
  return (size_t)nq * c * kc * (sizeof(int32_t) + sizeof(float));
}

__global__ void __launch_bounds__(TNT) topk_simt_kernel(
    const float* __restrict__ q, const float* __restrict__ g, int64_t ld,
    const float* __restrict__ g_sqn, int64_t nq, int64_t ng, int64_t d, int kc, int64_t goff,
    int64_t chunk_rows, int nchunks, int32_t* __restrict__ out_idx, float* __restrict__ out_key) {
  __shared__ float Qs[TK][TQ + 1];
  __shared__ float Gs[TK][TG + 1];
  __shared__ float Ks[TQ][TG + 1];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int64_t q0 = (int64_t)blockIdx.x * TQ;
  const int chunk = blockIdx.y;
  const int64_t g_begin = (int64_t)chunk * chunk_rows;
  int64_t g_end = g_begin + chunk_rows;
  if (g_end > ng) g_end = ng;

  float bk[kMaxKc];
  int32_t bi[kMaxKc];
  if (t < TQ)
    for (int e = 0; e < kc; ++e) { bk[e] = CUDART_INF_F; bi[e] = 0x7fffffff; }

  const int lr = t >> 2, lk = (t & 3) * 4;
  for (int64_t j0 = g_begin; j0 < g_end; j0 += TG) {
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int64_t k0 = 0; k0 < d; k0 += TK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t k = k0 + lk + e;
        Qs[lk + e][lr] = (q0 + lr < nq && k < d) ? q[(q0 + lr) * ld + k] : 0.f;
        Gs[lk + e][lr] = (j0 + lr < g_end && k < d) ? g[(j0 + lr) * ld + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = Qs[kk][ty * 4 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = Gs[kk][tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t j = j0 + tx + 16 * c;
        Ks[ty * 4 + r][tx + 16 * c] = (j < g_end) ? fmaf(-2.0f, acc[r][c], g_sqn[j]) : CUDART_INF_F;
      }
    __syncthreads();
    if (t < TQ) {
      for (int c = 0; c < TG; ++c) {
        const float key = Ks[t][c];
        const int32_t id = (int32_t)(goff + j0 + c);
        // gallery is scanned in increasing index order: strict '<' keeps the lowest index on ties
        if (key < bk[kc - 1]) {
          int p = kc - 1;
          while (p > 0 && key < bk[p - 1]) { bk[p] = bk[p - 1]; bi[p] = bi[p - 1]; --p; }
          bk[p] = key;
          bi[p] = id;
        }
      }
    }
    __syncthreads();
  }
  if (t < TQ && q0 + t < nq) {
    const int64_t base = ((q0 + t) * nchunks + chunk) * kc;
    for (int e = 0; e < kc; ++e) {
      out_idx[base + e] = bi[e] == 0x7fffffff ? -1 : bi[e];
      out_key[base + e] = bk[e];
    }
  }
}

int topk_candidates_f32(const float* q, const float* g, int64_t ld, const float* g_sqn, int64_t nq,
                        int64_t ng, int64_t d, int kc, int64_t goff, int32_t* cand_idx,
                        float* cand_key, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int nchunks = topk_f32_chunks(nq, ng);
  const int64_t chunk_rows = ceil_div(ceil_div(ng, nchunks), TG) * TG;
  int32_t* o_idx = cand_idx;
  float* o_key = cand_key;
  if (nchunks > 1) {
    o_idx = (int32_t*)ws;
    o_key = (float*)((char*)ws + (size_t)nq * nchunks * kc * sizeof(int32_t));
  }
  dim3 grid((unsigned)ceil_div(nq, TQ), (unsigned)nchunks);
  topk_simt_kernel<<<grid, TNT, 0, st>>>(q, g, ld, g_sqn, nq, ng, d, kc, goff, chunk_rows, nchunks, o_idx, o_key);
  PLK_LAUNCHED(1);
  if (nchunks > 1) return select_candidates(o_idx, o_key, nq, nchunks * kc, kc, cand_idx, cand_key, st);
  return PLK_OK;
}

}  // namespace plk
