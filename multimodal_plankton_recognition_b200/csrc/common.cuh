// Shared host-side helpers for libplk.so (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/plk.h"

namespace plk {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

constexpr float kNormEps = 1e-12f;   // F.normalize eps, reference src/coordination.py:33-34
constexpr float kLog2e = 1.4426950408889634f;
// Range-centred fixed shift of the InfoNCE exponentials: E_ij = exp(S_ij - s + kShiftK), s = exp(logit_scale).
// |u.v| <= 1 bounds every logit by s, so E <= e^64 and a sum over up to e^24 terms stays finite in fp32,
// while a row / column keeps a non-zero sum as long as its best match satisfies s (1 - cos_max) < 64 + 87
// (s <= 75 for ANY data, s <= 151 when every row has a non-negative best cosine; the plain shift by s
// stopped at s = 43).  One exp per logit still serves the row AND the column sums; the constant cancels in
// G = E (1/R + 1/C) and enters the loss as 2 (s - kShiftK).
constexpr float kShiftK = 64.0f;

#define PLK_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::plk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                        \
      return PLK_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define PLK_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::plk::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// ---- accurate fp32 logistic helpers (SigLIP epilogues of the fp32 path and the gradient tails) ----
#ifdef __CUDACC__
__device__ __forceinline__ float sigmoid_f(float z) {
  const float e = expf(-fabsf(z));
  return z >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
}
__device__ __forceinline__ float softplus_f(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
// Scale and diagonal term of the gradient tail (plk_infonce_grad_finish*):
//   InfoNCE (bias == nullptr): coef = g s / (2B), dterm = E_ii (1/rs_i + 1/cs_i) - 2
//   SigLIP  (bias != nullptr): coef = g s / B,    dterm = -sigmoid(-(S_ii + bias))
__device__ __forceinline__ void tail_terms(const float* bias, float diag_i, const float* rs, const float* cs,
                                           int64_t row, float s, float go, int64_t batch, float& coef,
                                           float& dterm) {
  if (bias != nullptr) {
    dterm = -sigmoid_f(-(diag_i + *bias));
    coef = go * s / (float)batch;
  } else {
    dterm = expf(diag_i - s + kShiftK) * (1.0f / rs[row] + 1.0f / cs[row]) - 2.0f;
    coef = go * s / (2.0f * (float)batch);
  }
}
// One row of the gradient tail, the row (width NV*128 fp32) held in the registers of ONE warp:
//   acc_i = sum of the `parts` partial slabs;  dU = coef (acc_i + dterm p / den_p);
//   dx = (dU - u (u . dU)) / den   (no projection term below the eps clamp of F.normalize)
// Shared by the stand-alone tail kernels (elementwise.cu) and the tail fused into infonce_grad_tc4.
// The row's own and partner raw rows (inputs of the step, never written by it): loaded first -- in the stand-alone
// tail even before griddepcontrol.wait, under the last thread blocks of the backward.
template <int NV>
__device__ __forceinline__ void finish_row_load_inputs(const float* __restrict__ x_row, const float* __restrict__ p_row,
                                                       float4 (&xv)[NV], float4 (&pv)[NV], int lane) {
  const float4* xr = reinterpret_cast<const float4*>(x_row);
  const float4* pr = reinterpret_cast<const float4*>(p_row);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    xv[i] = xr[lane + 32 * i];
    pv[i] = pr[lane + 32 * i];
  }
}
template <int NV>
__device__ __forceinline__ void finish_row_vec_loaded(const float* __restrict__ acc_row, int parts, int64_t slab4,
                                                      float4 (&xv)[NV], float4 (&pv)[NV], float coef, float dterm,
                                                      float idx_, float idp, bool clamped, float* __restrict__ dx_row,
                                                      int lane) {
  const float4* ar = reinterpret_cast<const float4*>(acc_row);
  // every load of the row's partial slabs is issued before the first use
  float4 acc[NV], acc1[NV];
  const bool two = parts > 1;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = ar[lane + 32 * i];
    acc1[i] = two ? ar[lane + 32 * i + slab4] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i].x += acc1[i].x; acc[i].y += acc1[i].y; acc[i].z += acc1[i].z; acc[i].w += acc1[i].w;
  }
  for (int p = 2; p < parts; ++p) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 t = ar[lane + 32 * i + p * slab4];
      acc[i].x += t.x; acc[i].y += t.y; acc[i].z += t.z; acc[i].w += t.w;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i].x = coef * fmaf(dterm, pv[i].x * idp, acc[i].x);
    acc[i].y = coef * fmaf(dterm, pv[i].y * idp, acc[i].y);
    acc[i].z = coef * fmaf(dterm, pv[i].z * idp, acc[i].z);
    acc[i].w = coef * fmaf(dterm, pv[i].w * idp, acc[i].w);
    xv[i].x *= idx_; xv[i].y *= idx_; xv[i].z *= idx_; xv[i].w *= idx_;
    dot = fmaf(xv[i].x, acc[i].x, fmaf(xv[i].y, acc[i].y, fmaf(xv[i].z, acc[i].z, fmaf(xv[i].w, acc[i].w, dot))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (clamped) dot = 0.f;
  float4* dr = reinterpret_cast<float4*>(dx_row);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    dr[lane + 32 * i] = make_float4((acc[i].x - xv[i].x * dot) * idx_, (acc[i].y - xv[i].y * dot) * idx_,
                                    (acc[i].z - xv[i].z * dot) * idx_, (acc[i].w - xv[i].w * dot) * idx_);
}
template <int NV>
__device__ __forceinline__ void finish_row_vec(const float* __restrict__ acc_row, int parts, int64_t slab4,
                                               const float* __restrict__ x_row, const float* __restrict__ p_row,
                                               float coef, float dterm, float idx_, float idp, bool clamped,
                                               float* __restrict__ dx_row, int lane) {
  float4 xv[NV], pv[NV];
  finish_row_load_inputs<NV>(x_row, p_row, xv, pv, lane);
  finish_row_vec_loaded<NV>(acc_row, parts, slab4, xv, pv, coef, dterm, idx_, idp, clamped, dx_row, lane);
}
#endif

// ---- programmatic dependent launch (sm_90+) ---------------------------------------------------
// A grid launched with launch_overlapped() may become resident while the previous kernel in the
// stream is still running: once every thread block of that kernel has executed pdl_trigger() (or
// exited).  It must execute pdl_wait() -- the previous kernel has completed and its writes are
// visible -- before touching anything that kernel produces or still reads.  Both are no-ops for
// normally launched grids / grids without dependents.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
cudaError_t launch_overlapped(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define PLK_LAUNCHED(n)                     \
  do {                                      \
    ::plk::count_launch(n);                 \
    PLK_CUDA(cudaGetLastError());           \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- entry points implemented per translation unit (dispatched from api.cu) ----
int infonce_fwd_f32(const float* u, const float* v, int64_t ld, int64_t n_rows, int64_t row_offset,
                    int64_t n_cols, int64_t d, int64_t bs, const float* ls, float* row_sumexp,
                    float* col_sumexp, float* diag, int sums_zeroed, cudaStream_t st);
int infonce_grad_f32(const float* a, const float* b, int64_t ld, int64_t n_rows, int64_t row_offset,
                     int64_t n_cols, int64_t d, int64_t bs, const float* ls, const float* rs,
                     const float* cs, float* acc, float* gs, cudaStream_t st);
int topk_candidates_f32(const float* q, const float* g, int64_t ld, const float* g_sqn, int64_t nq,
                        int64_t ng, int64_t d, int kc, int64_t goff, int32_t* cand_idx,
                        float* cand_key, void* ws, size_t ws_bytes, cudaStream_t st);
size_t topk_ws_f32(int64_t nq, int64_t ng, int64_t d, int kc);

// 16-bit tensor-core path: f16 = 0 -> bf16 operands, 1 -> fp16 operands
int infonce_fwd_tc16(const void* u, const void* v, int f16, int64_t ld, int64_t n_rows,
                     int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* ls,
                     float* row_sumexp, float* col_sumexp, float* diag, int sums_zeroed, cudaStream_t st);
int infonce_grad_tc16(const void* a, const void* b, int f16, int64_t ld, int64_t n_rows,
                      int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* ls,
                      const float* rs, const float* cs, float* acc, float* gs, cudaStream_t st);
int grad_parts_tc16(int64_t n_rows, int64_t n_cols, int64_t d, int64_t bs, int ndir);
int finish_pair_scaled(const float* acc_x, const float* acc_y, int parts, const float* x, const float* y, int64_t n,
                       int64_t d, int64_t ldx, const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                       const float* nrm_y, const float* diag, const float* rs, const float* cs,
                       const float* logit_scale, const float* grad_out_emb, float emb_scale, const float* grad_out,
                       int64_t batch_global, float* gs, const float* diag_sum, float* dx, float* dy, float* dls_out,
                       const float* loss_partial, void* const* peer_bufs, int rank, int world, unsigned* epoch,
                       float* out2, void* stream);
int infonce_grad_pair_tc16(const void* a0, const void* b0, const void* a1, const void* b1, int f16,
                           int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d,
                           int64_t bs, const float* ls, const float* rs0, const float* cs0,
                           const float* rs1, const float* cs1, float* acc0, float* acc1, float* gs,
                           cudaStream_t st, int overlap_prev = 0, const float* siglip_bias = nullptr);
// gradient tail fused into the d <= 256 recompute backward (infonce_tc.cu)
struct GradTailHost {
  const float *x, *y;                       // raw fp32 rows [B, ldx]
  int64_t ldx, batch;
  const float *inv_den_x, *nrm_x, *inv_den_y, *nrm_y, *diag;
  const float *grad_out_emb, *grad_out;
  float emb_scale;
  const float* diag_sum;
  float *dx, *dy, *dls_out;
  int* counters;                            // 2 * ceil(B/128) + 1 ints, zero on entry, left zero
};
bool grad_tail_fusable(int64_t n_rows, int64_t n_cols, int64_t d, int64_t bs, int64_t ldx, const void* x, const void* y,
                       const void* dx, const void* dy, const void* acc);
int infonce_grad_pair_tc16_tail(const void* a0, const void* b0, const void* a1, const void* b1, int f16, int64_t ld,
                                int64_t n_rows, int64_t d, const float* ls, const float* rs0, const float* cs0,
                                const float* rs1, const float* cs1, float* acc0, float* acc1, float* gs, cudaStream_t st,
                                int overlap_prev, const GradTailHost& th);
// SigLIP variants (reference src/coordination.py:67-95): fp32 CUDA-core path / tensor-core forward
int siglip_fwd_f32(const float* u, const float* v, int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                   int64_t d, int64_t bs, const float* ls, const float* bias, float* diag, double* sums,
                   cudaStream_t st);
int siglip_grad_f32(const float* a, const float* b, int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                    int64_t d, int64_t bs, const float* ls, const float* bias, float* acc, float* gs2,
                    cudaStream_t st);
int siglip_fwd_tc16(const void* u, const void* v, int f16, int64_t ld, int64_t n_rows, int64_t row_offset,
                    int64_t n_cols, int64_t d, int64_t bs, const float* ls, const float* bias, float* diag,
                    double* sums, cudaStream_t st);
int topk_candidates_tc16(const void* q, const void* g, int f16, int64_t ld, const float* g_sqn,
                         int64_t nq, int64_t ng, int64_t d, int kc, int64_t goff, int32_t* cand_idx,
                         float* cand_key, void* ws, size_t ws_bytes, cudaStream_t st);
size_t topk_ws_tc16(int64_t nq, int64_t ng, int64_t d, int kc);
int zero2(float* a, int64_t na, float* b, int64_t nb, cudaStream_t st);
int select_candidates(const int32_t* in_idx, const float* in_key, int64_t nq, int m, int kc,
                      int32_t* out_idx, float* out_key, cudaStream_t st);

// Bucket column range of a global row (block-diagonal InfoNCE, reference src/coordination.py:29-37).
__host__ __device__ inline void bucket_range(int64_t gi, int64_t bs, int64_t n_cols, int64_t& lo,
                                             int64_t& hi) {
  int64_t b = gi / bs;
  lo = b * bs;
  hi = lo + bs;
  if (hi > n_cols) hi = n_cols;
}

}  // namespace plk
