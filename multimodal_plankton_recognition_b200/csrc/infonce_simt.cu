// fp32 CUDA-core path of the fused similarity + InfoNCE forward / recompute backward
// ("parity mode": <=1e-5 relative vs the reference's fp32 PyTorch run).  Same algorithm and
// outputs as the tcgen05 path in infonce_tc.cu -- shared-memory tiled FFMA GEMM with the
// temperature / exp / row+column sum / diagonal epilogue fused, logits never leave the SM.
#include "common.cuh"

namespace plk {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// 64x64 tile of A.B^T (K = d) accumulated in a 4x4 register micro-tile per thread.
// Thread (ty, tx) = (t/16, t%16) owns rows ty*4+r and columns tx + 16*c (bank-conflict free).
__device__ __forceinline__ void tile_gemm_64x64(const float* __restrict__ A, int64_t a_rows,
                                                int64_t i0, const float* __restrict__ Bm,
                                                int64_t b_rows, int64_t j0, int64_t d, int64_t ld,
                                                float (*As)[BM + 1], float (*Bs)[BN + 1],
                                                float acc[4][4]) {
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const int lr = t >> 2;        // 0..63 : tile row loaded by this thread
  const int lk = (t & 3) * 4;   // 0,4,8,12
  for (int64_t k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t k = k0 + lk + e;
      const int64_t ia = i0 + lr, jb = j0 + lr;
      As[lk + e][lr] = (ia < a_rows && k < d) ? A[ia * ld + k] : 0.f;
      Bs[lk + e][lr] = (jb < b_rows && k < d) ? Bm[jb * ld + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
}

// Column range a row tile needs: union of the buckets of its rows.
__device__ __forceinline__ void tile_col_range(int64_t i0, int64_t n_rows, int64_t row_offset,
                                               int64_t bs, int64_t n_cols, int64_t& jlo,
                                               int64_t& jhi) {
  int64_t last = i0 + BM - 1;
  if (last >= n_rows) last = n_rows - 1;
  int64_t lo, hi, lo2, hi2;
  bucket_range(row_offset + i0, bs, n_cols, lo, hi);
  bucket_range(row_offset + last, bs, n_cols, lo2, hi2);
  jlo = lo;
  jhi = hi2;
}

// SigLIP forward (reference src/coordination.py:85-93) on the same tiles: z = S + bias, the loss term is
// softplus(z) off the diagonal and softplus(-z) on it; no row / column statistics.  Accumulates
// sums[0] += loss terms, sums[1] += G_ii S_ii, sums[2] += G_ii with G_ii = -sigmoid(-z_ii).
__global__ void __launch_bounds__(NT) siglip_fwd_simt(
    const float* __restrict__ u, const float* __restrict__ v, int64_t ld, int64_t n_rows,
    int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* __restrict__ ls,
    const float* __restrict__ bias, float* __restrict__ diag, double* __restrict__ sums) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ float red[3][NT / 32];
  const int64_t i0 = (int64_t)blockIdx.y * BM;
  const int64_t j0 = (int64_t)blockIdx.x * BN;
  int64_t jlo, jhi;
  tile_col_range(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  if (j0 + BN <= jlo || j0 >= jhi) return;
  float acc[4][4];
  tile_gemm_64x64(u, n_rows, i0, v, n_cols, j0, d, ld, As, Bs, acc);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float s = expf(*ls), b0 = *bias;
  float part[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + ty * 4 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    if (i < n_rows) bucket_range(gi, bs, n_cols, lo, hi);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t j = j0 + tx + 16 * c;
      if (!(i < n_rows && j >= lo && j < hi)) continue;
      const float S = s * acc[r][c];
      const float z = S + b0;
      if (j == gi) {
        diag[i] = S;
        part[0] += softplus_f(-z);
        const float g = -sigmoid_f(-z);
        part[1] = fmaf(g, S, part[1]);
        part[2] += g;
      } else {
        part[0] += softplus_f(z);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
    if ((t & 31) == 0) red[k][t >> 5] = part[k];
  }
  __syncthreads();
  if (t < 3) {
    double tot = 0.0;
    for (int w = 0; w < NT / 32; ++w) tot += (double)red[t][w];
    if (tot != 0.0) atomicAdd(sums + t, tot);
  }
}

__global__ void __launch_bounds__(NT) infonce_fwd_simt(
    const float* __restrict__ u, const float* __restrict__ v, int64_t ld, int64_t n_rows,
    int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* __restrict__ ls,
    float* __restrict__ row_sumexp, float* __restrict__ col_sumexp, float* __restrict__ diag) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ float colsum[BN];
  const int64_t i0 = (int64_t)blockIdx.y * BM;
  const int64_t j0 = (int64_t)blockIdx.x * BN;
  int64_t jlo, jhi;
  tile_col_range(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  if (j0 + BN <= jlo || j0 >= jhi) return;  // tile outside every bucket of these rows

  float acc[4][4];
  tile_gemm_64x64(u, n_rows, i0, v, n_cols, j0, d, ld, As, Bs, acc);

  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  if (t < BN) colsum[t] = 0.f;
  __syncthreads();
  const float s = expf(*ls);
  float cpart[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + ty * 4 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    if (i < n_rows) bucket_range(gi, bs, n_cols, lo, hi);
    float rpart = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t j = j0 + tx + 16 * c;
      const bool valid = i < n_rows && j >= lo && j < hi;
      const float S = s * acc[r][c];
      const float E = valid ? expf(S - s + kShiftK) : 0.f;
      if (valid && j == gi) diag[i] = S;
      rpart += E;
      cpart[c] += E;
    }
    // reduce over the 16 lanes (tx) that share this row
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rpart += __shfl_xor_sync(0xffffffffu, rpart, o);
    if (tx == 0 && i < n_rows) atomicAdd(row_sumexp + i, rpart);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) atomicAdd(&colsum[tx + 16 * c], cpart[c]);
  __syncthreads();
  if (t < BN && j0 + t < n_cols && colsum[t] != 0.f) atomicAdd(col_sumexp + j0 + t, colsum[t]);
}

// One direction of the recompute backward.  CTA = (64 owned rows) x (128 output columns of d);
// loops over the column tiles of the rows' buckets: S tile -> G tile in shared memory -> G.B.
// SIG: SigLIP weights G = sigmoid(S + bias) off the diagonal (rs / cs unused); gs_out is then
// float[2] = (sum G*S, sum G).
constexpr int DC = 128;
template <bool SIG>
__global__ void __launch_bounds__(NT) infonce_grad_simt(
    const float* __restrict__ a, const float* __restrict__ b, int64_t ld, int64_t n_rows,
    int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* __restrict__ ls,
    const float* __restrict__ rs, const float* __restrict__ cs, float* __restrict__ acc_out,
    float* __restrict__ gs_out, const float* __restrict__ bias) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ float Gs[BM][BN + 1];
  __shared__ float Ys[BK][DC];
  __shared__ float rcs[BN];
  __shared__ float red[NT / 32];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int64_t i0 = (int64_t)blockIdx.y * BM;
  const int64_t dc0 = (int64_t)blockIdx.x * DC;
  int64_t jlo, jhi;
  tile_col_range(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const float s = expf(*ls);

  float rrs[4];
  int64_t lo[4], hi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + ty * 4 + r;
    lo[r] = hi[r] = 0;
    rrs[r] = 0.f;
    if (i < n_rows) {
      bucket_range(row_offset + i, bs, n_cols, lo[r], hi[r]);
      if constexpr (!SIG) rrs[r] = 1.0f / rs[i];
    }
  }
  float b0 = 0.f;
  if constexpr (SIG) b0 = *bias;
  float gsum_local = 0.f;
  float out[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) out[r][c] = 0.f;
  float gs_local = 0.f;

  for (int64_t j0 = jlo; j0 < jhi; j0 += BN) {
    float acc[4][4];
    tile_gemm_64x64(a, n_rows, i0, b, n_cols, j0, d, ld, As, Bs, acc);
    if constexpr (!SIG) {
      if (t < BN) rcs[t] = (j0 + t < n_cols) ? 1.0f / cs[j0 + t] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t j = j0 + tx + 16 * c;
        const bool valid = j >= lo[r] && j < hi[r];
        const float S = s * acc[r][c];
        const bool on_diag = j == row_offset + i0 + ty * 4 + r;
        float G;
        if constexpr (SIG) {
          G = (valid && !on_diag) ? sigmoid_f(S + b0) : 0.f;
          gsum_local += G;
        } else {
          G = valid ? expf(S - s + kShiftK) * (rrs[r] + rcs[tx + 16 * c]) : 0.f;
        }
        gs_local = fmaf(G, S, gs_local);
        // the j == i term is added (in fp32, together with -2*delta) by plk_infonce_grad_finish
        Gs[ty * 4 + r][tx + 16 * c] = on_diag ? 0.f : G;
      }
    __syncthreads();
    // out[64 x 128] += Gs[64 x 64] . b[j0:j0+64, dc0:dc0+128]
    for (int jj0 = 0; jj0 < BN; jj0 += BK) {
      for (int e = t; e < BK * DC; e += NT) {
        const int jr = e / DC, dcol = e % DC;
        const int64_t j = j0 + jj0 + jr, k = dc0 + dcol;
        Ys[jr][dcol] = (j < n_cols && k < d) ? b[j * ld + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float g[4], y[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r] = Gs[ty * 4 + r][jj0 + kk];
#pragma unroll
        for (int c = 0; c < 8; ++c) y[c] = Ys[kk][tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) out[r][c] = fmaf(g[r], y[c], out[r][c]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = i0 + ty * 4 + r;
    if (i >= n_rows) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int64_t k = dc0 + tx + 16 * c;
      if (k < d) acc_out[i * d + k] = out[r][c];
    }
  }
  if (gs_out != nullptr && blockIdx.x == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
    if ((t & 31) == 0) red[t >> 5] = gs_local;
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      for (int w = 0; w < NT / 32; ++w) tot += red[w];
      atomicAdd(gs_out, tot);
    }
    if constexpr (SIG) {
      __syncthreads();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      if ((t & 31) == 0) red[t >> 5] = gsum_local;
      __syncthreads();
      if (t == 0) {
        float tot = 0.f;
        for (int w = 0; w < NT / 32; ++w) tot += red[w];
        atomicAdd(gs_out + 1, tot);
      }
    }
  }
}

int infonce_fwd_f32(const float* u, const float* v, int64_t ld, int64_t n_rows, int64_t row_offset,
                    int64_t n_cols, int64_t d, int64_t bs, const float* ls, float* row_sumexp,
                    float* col_sumexp, float* diag, int sums_zeroed, cudaStream_t st) {
  int rc;
  // sums are accumulated with atomics -> zero first; diag needs no init (every owned row has its
  // diagonal column inside its bucket, so it is always written)
  if (!sums_zeroed && (rc = zero2(row_sumexp, n_rows, col_sumexp, n_cols, st))) return rc;
  dim3 grid((unsigned)ceil_div(n_cols, BN), (unsigned)ceil_div(n_rows, BM));
  infonce_fwd_simt<<<grid, NT, 0, st>>>(u, v, ld, n_rows, row_offset, n_cols, d, bs, ls, row_sumexp,
                                        col_sumexp, diag);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int infonce_grad_f32(const float* a, const float* b, int64_t ld, int64_t n_rows, int64_t row_offset,
                     int64_t n_cols, int64_t d, int64_t bs, const float* ls, const float* rs,
                     const float* cs, float* acc, float* gs, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(d, DC), (unsigned)ceil_div(n_rows, BM));
  infonce_grad_simt<false><<<grid, NT, 0, st>>>(a, b, ld, n_rows, row_offset, n_cols, d, bs, ls, rs, cs, acc, gs,
                                                nullptr);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int siglip_fwd_f32(const float* u, const float* v, int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                   int64_t d, int64_t bs, const float* ls, const float* bias, float* diag, double* sums,
                   cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(n_cols, BN), (unsigned)ceil_div(n_rows, BM));
  siglip_fwd_simt<<<grid, NT, 0, st>>>(u, v, ld, n_rows, row_offset, n_cols, d, bs, ls, bias, diag, sums);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int siglip_grad_f32(const float* a, const float* b, int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                    int64_t d, int64_t bs, const float* ls, const float* bias, float* acc, float* gs2,
                    cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(d, DC), (unsigned)ceil_div(n_rows, BM));
  infonce_grad_simt<true><<<grid, NT, 0, st>>>(a, b, ld, n_rows, row_offset, n_cols, d, bs, ls, nullptr, nullptr, acc,
                                               gs2, bias);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

}  // namespace plk
