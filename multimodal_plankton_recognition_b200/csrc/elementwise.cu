// HBM-bound row-wise kernels around the similarity GEMMs: L2 normalisation, loss reduction,
// gradient tail (diag term + normalisation backward), candidate re-score / merge, k-NN vote.
// One warp per row, coalesced 32-lane sweeps along d; grids are sized from the row count.
#include <math_constants.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace plk {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__device__ __forceinline__ void st_from_float(T* p, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void st_from_float<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename T> constexpr bool kIsHalf = false;
template <> constexpr bool kIsHalf<__half> = true;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// a2: u = x / max(||x||, eps)                                    reference src/coordination.py:33-34
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) l2norm_kernel(const TI* __restrict__ x, int64_t n, int64_t d,
                                                     int64_t ldx, TO* __restrict__ u, int64_t ldu,
                                                     float* __restrict__ inv_den,
                                                     float* __restrict__ nrm_out,
                                                     float* __restrict__ sqn_out, int normalise) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const TI* xr = x + row * ldx;
  float ss = 0.f;
  for (int64_t k = lane; k < d; k += 32) {
    float v = ld_as_float(xr + k);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float den = fmaxf(nrm, kNormEps);
  const float scale = normalise ? 1.0f / den : 1.0f;
  TO* ur = u + row * ldu;
  float sq = 0.f;
  for (int64_t k = lane; k < ldu; k += 32) {
    float v = 0.f;
    if (k < d) v = normalise ? ld_as_float(xr + k) / den : ld_as_float(xr + k);
    st_from_float(ur + k, v);
    float w = ld_as_float(ur + k);  // value as stored (after rounding)
    sq = fmaf(w, w, sq);
  }
  sq = warp_sum(sq);
  if (lane == 0) {
    if (inv_den) inv_den[row] = scale;
    if (nrm_out) nrm_out[row] = nrm;
    if (sqn_out) sqn_out[row] = sq;
  }
}

// Fast path: fp32 rows with d % 128 == 0 (d <= 1024) held entirely in registers, one warp per row,
// every global access a 16-byte vector with all loads of a row in flight at once.
template <int NV, typename TO>
__global__ void __launch_bounds__(256) l2norm_vec_kernel(const float* __restrict__ x, int64_t n, int64_t ldx,
                                                         TO* __restrict__ u, int64_t ldu,
                                                         float* __restrict__ inv_den,
                                                         float* __restrict__ nrm_out,
                                                         float* __restrict__ sqn_out, int normalise) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) ss = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, ss))));
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float den = fmaxf(nrm, kNormEps);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 w = v[i];
    if (normalise) { w.x /= den; w.y /= den; w.z /= den; w.w /= den; }
    if constexpr (sizeof(TO) == 4) {
      reinterpret_cast<float4*>(u + row * ldu)[lane + 32 * i] = w;
    } else if constexpr (sizeof(TO) == 2 && !kIsHalf<TO>) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      reinterpret_cast<uint2*>(u + row * ldu)[lane + 32 * i] = pk;
      w.x = __bfloat162float(lo.x); w.y = __bfloat162float(lo.y);
      w.z = __bfloat162float(hi.x); w.w = __bfloat162float(hi.y);
    } else {
      __half2 lo = __floats2half2_rn(w.x, w.y), hi = __floats2half2_rn(w.z, w.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      reinterpret_cast<uint2*>(u + row * ldu)[lane + 32 * i] = pk;
      w.x = __half2float(lo.x); w.y = __half2float(lo.y);
      w.z = __half2float(hi.x); w.w = __half2float(hi.y);
    }
    sq = fmaf(w.x, w.x, fmaf(w.y, w.y, fmaf(w.z, w.z, fmaf(w.w, w.w, sq))));
  }
  if (sqn_out) sq = warp_sum(sq);
  if (lane == 0) {
    if (inv_den) inv_den[row] = normalise ? 1.0f / den : 1.0f;
    if (nrm_out) nrm_out[row] = nrm;
    if (sqn_out) sqn_out[row] = sq;
  }
}

template <typename TO>
static bool l2norm_try_vec(const float* x, int64_t n, int64_t d, int64_t ldx, TO* u, int64_t ldu,
                           float* inv_den, float* nrm, float* sqn, int normalise, cudaStream_t st) {
  if (d % 128 != 0 || d > 1024 || ldu != d || (ldx & 3) || ((uintptr_t)x & 15) || ((uintptr_t)u & 15)) return false;
  dim3 block(256), grid((unsigned)ceil_div(n, 8));
  switch (d / 128) {
#define PLK_CASE(NV) case NV: l2norm_vec_kernel<NV, TO><<<grid, block, 0, st>>>(x, n, ldx, u, ldu, inv_den, nrm, sqn, normalise); return true;
    PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8)
#undef PLK_CASE
  }
  return false;
}

// Both modalities in one launch (blockIdx.y selects), optionally zero-filling the two sum-exp
// accumulators the forward kernel adds into: 3 launches -> 1 on the launch-bound small-batch path.
struct NormPairArgs {
  const float* x[2];
  void* u[2];
  float* inv_den[2];
  float* nrm[2];
  float* zero[2];
  int64_t nzero[2];
};
// Two rows per warp: all four 16-byte loads of a lane are issued before the first use, which doubles the
// bytes in flight per SM -- this is the first kernel of a step, its inputs come from HBM.
template <int NV, typename TO>
__global__ void __launch_bounds__(256) l2norm_pair_vec_kernel(NormPairArgs a, int64_t n, int64_t ldx, int64_t ldu) {
  constexpr int RPW = 2;
  const int m = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  pdl_trigger();   // the forward kernel may set itself up; it waits for this grid before reading
  {  // zero fill: thread-linear over this modality's accumulator
    const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a.zero[m] != nullptr && z < a.nzero[m]) a.zero[m][z] = 0.f;
  }
  if (row0 >= n) return;
  TO* u = reinterpret_cast<TO*>(a.u[m]);
  float4 v[RPW][NV];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r < n ? row0 + r : row0;   // odd tail: re-read the first row, write nothing
    const float4* xr = reinterpret_cast<const float4*>(a.x[m] + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[r][i] = xr[lane + 32 * i];
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r;
    if (row >= n) continue;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      ss = fmaf(v[r][i].x, v[r][i].x, fmaf(v[r][i].y, v[r][i].y, fmaf(v[r][i].z, v[r][i].z, fmaf(v[r][i].w, v[r][i].w, ss))));
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, kNormEps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 w = v[r][i];
      w.x /= den; w.y /= den; w.z /= den; w.w /= den;
      if constexpr (sizeof(TO) == 4) {
        reinterpret_cast<float4*>(u + row * ldu)[lane + 32 * i] = w;
      } else if constexpr (!kIsHalf<TO>) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
        reinterpret_cast<uint2*>(u + row * ldu)[lane + 32 * i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      } else {
        __half2 lo = __floats2half2_rn(w.x, w.y), hi = __floats2half2_rn(w.z, w.w);
        reinterpret_cast<uint2*>(u + row * ldu)[lane + 32 * i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
    if (lane == 0) {
      a.inv_den[m][row] = 1.0f / den;
      a.nrm[m][row] = nrm;
    }
  }
}

template <typename TO>
static bool l2norm_pair_try(const NormPairArgs& a, int64_t n, int64_t d, int64_t ldx, int64_t ldu, cudaStream_t st) {
  int64_t nz = a.nzero[0] > a.nzero[1] ? a.nzero[0] : a.nzero[1];
  int64_t blocks = ceil_div(n, 16);   // 8 warps x 2 rows
  if (ceil_div(nz, 256) > blocks) blocks = ceil_div(nz, 256);
  dim3 block(256), grid((unsigned)blocks, 2);
  switch (d / 128) {
#define PLK_CASE(NV) case NV: l2norm_pair_vec_kernel<NV, TO><<<grid, block, 0, st>>>(a, n, ldx, ldu); return true;
    PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8)
#undef PLK_CASE
  }
  return false;
}

template <typename TI>
static int l2norm_dispatch_out(const TI* x, int64_t n, int64_t d, int64_t ldx, void* u, int u_dtype,
                               int64_t ldu, float* inv_den, float* nrm, float* sqn, int normalise,
                               cudaStream_t st) {
  dim3 block(256), grid((unsigned)ceil_div(n, 8));
  if constexpr (sizeof(TI) == 4) {
    const bool done = u_dtype == PLK_F32
                          ? l2norm_try_vec((const float*)x, n, d, ldx, (float*)u, ldu, inv_den, nrm, sqn, normalise, st)
                          : u_dtype == PLK_BF16
                                ? l2norm_try_vec((const float*)x, n, d, ldx, (__nv_bfloat16*)u, ldu, inv_den, nrm, sqn, normalise, st)
                                : l2norm_try_vec((const float*)x, n, d, ldx, (__half*)u, ldu, inv_den, nrm, sqn, normalise, st);
    if (done) {
      PLK_LAUNCHED(1);
      return PLK_OK;
    }
  }
  if (u_dtype == PLK_F32)
    l2norm_kernel<TI, float><<<grid, block, 0, st>>>(x, n, d, ldx, (float*)u, ldu, inv_den, nrm, sqn, normalise);
  else if (u_dtype == PLK_BF16)
    l2norm_kernel<TI, __nv_bfloat16><<<grid, block, 0, st>>>(x, n, d, ldx, (__nv_bfloat16*)u, ldu, inv_den, nrm, sqn, normalise);
  else
    l2norm_kernel<TI, __half><<<grid, block, 0, st>>>(x, n, d, ldx, (__half*)u, ldu, inv_den, nrm, sqn, normalise);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

// Cross-GPU sum of the two per-rank scalars of a sharded step (loss partial, d logit_scale
// partial), fused into the gradient tail: the ranks exchange them through peer-mapped symmetric
// memory over NVLink -- a release/acquire flag per (parity, rank) and one 8-byte slot per parity --
// instead of a separate NCCL all-reduce launch (which costs ~25 us of a ~90 us step).
//   peer[r] -> rank r's buffer: float data[2 parities][2], then at +64 bytes uint32 flags[2][8].
// Every rank sums the slots in rank order, so the result is bitwise identical everywhere.
struct XGpuArgs {
  void* const* peer;          // device array [world] of peer-mapped base pointers (nullptr: single GPU)
  int rank, world;
  unsigned* epoch;            // two local device counters, incremented once per launch (CUDA-graph safe)
  const float* loss_partial;  // this rank's loss partial
  float* out2;                // OUT: (global loss, global d logit_scale)
  long long timeout;          // spin budget in SM clock cycles before the exchange is declared dead
};

// Peers are other processes (a rank can sit in a checkpoint write, a validation pass or a data-loader stall
// while the others wait): the default budget matches NCCL's watchdog, PLK_XGPU_TIMEOUT_S overrides it.
static long long xgpu_timeout_cycles() {
  static const long long cycles = [] {
    const char* e = getenv("PLK_XGPU_TIMEOUT_S");
    double sec = e != nullptr ? atof(e) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    return (long long)(sec * 2.0e9);
  }();
  return cycles;
}

// Publish (first thread block of the kernel) and collect (LAST thread block, which is scheduled near the
// end of the kernel) are separate, so the time the ranks are out of step with each other is
// absorbed by the kernel's own row work instead of stalling it.  epoch[0] / epoch[1] count the
// launches seen by the publisher / the collector.
__device__ __forceinline__ void xgpu_publish(const XGpuArgs& xg, float loss_part, float dls_part) {
  // executed by warp 0 of block (0,0); lane r signals rank r
  const int lane = threadIdx.x;
  unsigned e = 0;
  if (lane == 0) { e = xg.epoch[0] + 1; xg.epoch[0] = e; }
  e = __shfl_sync(0xffffffffu, e, 0);
  const int p = e & 1;
  if (lane == 0) {
    volatile float* mine = reinterpret_cast<volatile float*>(xg.peer[xg.rank]) + 2 * p;
    mine[0] = loss_part;
    mine[1] = dls_part;
    __threadfence_system();
  }
  __syncwarp();
  if (lane < xg.world) {
    unsigned* their_flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(xg.peer[lane]) + 64) + p * 8;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(their_flags + xg.rank), "r"(e) : "memory");
  }
}
__device__ __forceinline__ void xgpu_collect(const XGpuArgs& xg) {
  // executed by warp 0 of the last block; lane r waits for rank r (its own rank included)
  const int lane = threadIdx.x;
  unsigned e = 0;
  if (lane == 0) { e = xg.epoch[1] + 1; xg.epoch[1] = e; }
  e = __shfl_sync(0xffffffffu, e, 0);
  const int p = e & 1;
  float sl = 0.f, sd = 0.f;
  if (lane < xg.world) {
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(xg.peer[xg.rank]) + 64) + p * 8;
    unsigned seen = 0;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_flags + lane) : "memory");
      // a protocol error (or a dead peer) ends in a trap after xg.timeout cycles, not in a hang
      if (seen != e && clock64() - t0 > xg.timeout) {
        printf("plk: cross-GPU scalar exchange timed out (rank %d waiting for rank %d, epoch %u, saw %u)\n",
               xg.rank, lane, e, seen);
        __trap();
      }
    } while (seen != e);
    const volatile float* theirs = reinterpret_cast<const volatile float*>(xg.peer[lane]) + 2 * p;
    sl = theirs[0];
    sd = theirs[1];
  }
  // fixed-order sum (lane 0 adds ranks 0..world-1 in order): identical bits on every rank
  float tl = 0.f, td = 0.f;
  for (int r = 0; r < xg.world; ++r) {
    tl += __shfl_sync(0xffffffffu, sl, r);
    td += __shfl_sync(0xffffffffu, sd, r);
  }
  if (lane == 0) { xg.out2[0] = tl; xg.out2[1] = td; }
}

// ---------------------------------------------------------------------------------------------
// a8: loss partial over the owned rows                              reference src/coordination.py:45
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) loss_kernel(const float* __restrict__ rs,
                                                    const float* __restrict__ cs,
                                                    const float* __restrict__ diag,
                                                    const float* __restrict__ ls, int64_t n,
                                                    int64_t batch, float* __restrict__ loss_out,
                                                    float* __restrict__ diag_sum_out,
                                                    float* __restrict__ gs_zero, XGpuArgs xg,
                                                    float* __restrict__ partial_out) {
  // xg.peer != nullptr: *loss_out is the sum of the partials of all ranks (exchanged through peer
  // memory by warp 0, see xgpu_publish / xgpu_collect); this rank's partial goes to *partial_out.
  __shared__ double sh[2][32];
  pdl_wait();      // the forward kernel is complete
  // A kernel queued behind this one with programmatic serialization (the recompute backward) may
  // start now: the forward's results are final and it consumes nothing of ours before its own wait.
  pdl_trigger();
  if (threadIdx.x == 0 && gs_zero != nullptr) *gs_zero = 0.f;
  const double s = (double)expf(*ls);
  const double shift2 = 2.0 * (s - (double)kShiftK);   // both sum-exps carry exp(-(s - kShiftK)), see common.cuh
  double a = 0.0, dsum = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double dg = (double)diag[i];
    const float r = rs[i], c = cs[i];
    // A sum-exp that underflowed to zero (temperature beyond the range stated in common.cuh) or overflowed
    // must not come back as a finite-looking loss: poison it, the gradients follow (1/0 in the backward).
    const bool ok = r > 0.f && c > 0.f && r <= 3.0e38f && c <= 3.0e38f;
    a += ok ? shift2 + (double)logf(r) + (double)logf(c) - 2.0 * dg : (double)CUDART_NAN_F;
    dsum += dg;
  }
  a = warp_sum(a);
  dsum = warp_sum(dsum);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = dsum; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.0;
    dsum = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.0;
    a = warp_sum(a);
    dsum = warp_sum(dsum);
    const float part = (float)(a / (2.0 * (double)batch));
    if (threadIdx.x == 0) {
      if (diag_sum_out) *diag_sum_out = (float)dsum;
      if (xg.peer == nullptr) *loss_out = part;
      else *partial_out = part;
    }
    if (xg.peer != nullptr) {
      xgpu_publish(xg, threadIdx.x == 0 ? part : 0.f, 0.f);
      xgpu_collect(xg);
      __syncwarp();
      if (threadIdx.x == 0) *loss_out = xg.out2[0];
    }
  }
}

__global__ void dls_kernel(const float* gs, const float* diag_sum, const float* grad_out,
                           int64_t batch, float* out) {
  *out = (float)((double)(*grad_out) / (2.0 * (double)batch) * ((double)(*gs) - 2.0 * (double)(*diag_sum)));
}

// SigLIP scalars.  loss = sums[0] / B;  d logit_scale = g/B (gs2[0] + sums[1]);  d bias = g/B (gs2[1] + sums[2])
__global__ void siglip_loss_kernel(const double* __restrict__ sums, int64_t batch, float* __restrict__ loss_out) {
  pdl_wait();
  pdl_trigger();
  *loss_out = (float)(sums[0] / (double)batch);
}
__global__ void siglip_scalars_kernel(float* gs2, const double* __restrict__ sums, const float* __restrict__ grad_out,
                                      int64_t batch, float* dls_out, float* dbias_out) {
  const double g_b = (double)(*grad_out) / (double)batch;
  *dls_out = (float)(g_b * ((double)gs2[0] + sums[1]));
  *dbias_out = (float)(g_b * ((double)gs2[1] + sums[2]));
  gs2[0] = 0.f;
  gs2[1] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// a9 tail: -2*delta term, g*s/(2B) scaling, normalisation backward.
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) grad_finish_kernel(
    const float* __restrict__ acc, int parts, const TI* __restrict__ x,
    const TI* __restrict__ partner, int64_t n, int64_t d, int64_t ldx, const float* __restrict__ inv_den_x,
    const float* __restrict__ nrm_x, const float* __restrict__ inv_den_p,
    const float* __restrict__ diag, const float* __restrict__ rs, const float* __restrict__ cs,
    const float* __restrict__ ls, const float* __restrict__ grad_out, int64_t batch,
    TO* __restrict__ dx, const float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float s = expf(*ls);
  float coef, dterm;   // fp32 diagonal term (InfoNCE: G_ii - 2; SigLIP: G_ii)
  tail_terms(bias, diag[row], rs, cs, row, s, *grad_out, batch, coef, dterm);
  const float idx_ = inv_den_x[row], idp = inv_den_p[row];
  const bool clamped = !(nrm_x[row] > kNormEps);
  const float* ar = acc + row * d;
  const TI* xr = x + row * ldx;
  const TI* pr = partner + row * ldx;
  const int64_t slab = n * d;
  float dot = 0.f;
  for (int64_t k = lane; k < d; k += 32) {
    float a = ar[k];
    for (int p = 1; p < parts; ++p) a += ar[k + p * slab];
    float dU = coef * fmaf(dterm, ld_as_float(pr + k) * idp, a);
    dot = fmaf(ld_as_float(xr + k) * idx_, dU, dot);
  }
  dot = warp_sum(dot);
  if (clamped) dot = 0.f;  // below the eps clamp the denominator is constant: dx = dU / eps
  TO* dr = dx + row * d;
  for (int64_t k = lane; k < d; k += 32) {
    float a = ar[k];
    for (int p = 1; p < parts; ++p) a += ar[k + p * slab];
    float dU = coef * fmaf(dterm, ld_as_float(pr + k) * idp, a);
    float uk = ld_as_float(xr + k) * idx_;
    st_from_float(dr + k, (dU - uk * dot) * idx_);
  }
}

// Fast path (fp32 in / fp32 out, d % 128 == 0, d <= 1024): the row lives in registers, single pass.
template <int NV>
__global__ void __launch_bounds__(256) grad_finish_vec_kernel(
    const float* __restrict__ acc, int parts, const float* __restrict__ x,
    const float* __restrict__ partner, int64_t n, int64_t ldx, const float* __restrict__ inv_den_x,
    const float* __restrict__ nrm_x, const float* __restrict__ inv_den_p,
    const float* __restrict__ diag, const float* __restrict__ rs, const float* __restrict__ cs,
    const float* __restrict__ ls, const float* __restrict__ grad_out, int64_t batch,
    float* __restrict__ dx, const float* __restrict__ bias) {
  constexpr int64_t d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float4* ar = reinterpret_cast<const float4*>(acc + row * d);
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const float4* pr = reinterpret_cast<const float4*>(partner + row * ldx);
  const int64_t slab4 = n * d / 4;
  float4 a[NV], xv[NV], pv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    a[i] = ar[lane + 32 * i];
    xv[i] = xr[lane + 32 * i];
    pv[i] = pr[lane + 32 * i];
  }
  for (int p = 1; p < parts; ++p) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 t = ar[lane + 32 * i + p * slab4];
      a[i].x += t.x; a[i].y += t.y; a[i].z += t.z; a[i].w += t.w;
    }
  }
  const float s = expf(*ls);
  float coef, dterm;
  tail_terms(bias, diag[row], rs, cs, row, s, *grad_out, batch, coef, dterm);
  const float idx_ = inv_den_x[row], idp = inv_den_p[row];
  const bool clamped = !(nrm_x[row] > kNormEps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    a[i].x = coef * fmaf(dterm, pv[i].x * idp, a[i].x);
    a[i].y = coef * fmaf(dterm, pv[i].y * idp, a[i].y);
    a[i].z = coef * fmaf(dterm, pv[i].z * idp, a[i].z);
    a[i].w = coef * fmaf(dterm, pv[i].w * idp, a[i].w);
    xv[i].x *= idx_; xv[i].y *= idx_; xv[i].z *= idx_; xv[i].w *= idx_;
    dot = fmaf(xv[i].x, a[i].x, fmaf(xv[i].y, a[i].y, fmaf(xv[i].z, a[i].z, fmaf(xv[i].w, a[i].w, dot))));
  }
  dot = warp_sum(dot);
  if (clamped) dot = 0.f;
  float4* dr = reinterpret_cast<float4*>(dx + row * d);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    dr[lane + 32 * i] = make_float4((a[i].x - xv[i].x * dot) * idx_, (a[i].y - xv[i].y * dot) * idx_,
                                    (a[i].z - xv[i].z * dot) * idx_, (a[i].w - xv[i].w * dot) * idx_);
}

// Both modalities + d logit_scale in one launch (blockIdx.y selects the modality).
struct FinishPairArgs {
  const float* acc[2];
  const float* x[2];        // raw rows of this modality (partner = x[1 - m])
  const float* inv_den[2];
  const float* nrm[2];
  float* dx[2];
  float emb_scale;          // extra factor on the embedding gradients (world size under DDP averaging)
  // SigLIP mode (bias != nullptr): gs -> float[2] (sum G*S, sum G over the off-diagonal), sig_sums ->
  // double[3] from the forward (sum of loss terms, sum_i G_ii S_ii, sum_i G_ii), dbias_out = d loss / d bias
  const float* bias;
  const double* sig_sums;
  float* dbias_out;
};
// BOTH = true (d <= 512): one warp finishes row i of BOTH modalities -- the two raw rows are each other's partner, so
// they are loaded once instead of twice (32 MB instead of 40 MB through L2 per step at B = 4096, d = 256; the tail is
// bound by exactly that traffic).  BOTH = false: blockIdx.y selects the modality (wide rows: register budget).
template <int NV, bool BOTH>
__global__ void __launch_bounds__(256) grad_finish_pair_vec_kernel(
    FinishPairArgs a, int parts, int64_t n, int64_t ldx, const float* __restrict__ diag,
    const float* __restrict__ rs, const float* __restrict__ cs, const float* __restrict__ ls,
    const float* __restrict__ grad_out, const float* __restrict__ grad_out_dls, int64_t batch,
    float* __restrict__ gs, const float* __restrict__ diag_sum, float* __restrict__ dls_out, XGpuArgs xg) {
  constexpr int64_t d = NV * 128;
  const int m = BOTH ? 0 : blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // Launched under the tail of the recompute backward.  Everything that kernel does not write -- the raw rows, the
  // statistics of the normalisation and of the forward (both complete before the backward could start), the
  // upstream gradient -- is fetched BEFORE griddepcontrol.wait; only the partial slabs and sum G*S come after it.
  float4 xv[NV], pv[NV];
  float coef = 0.f, dterm = 0.f, idx_ = 0.f, idp = 0.f;
  bool clamped = false, clamped_p = false;
  if (row < n) {
    finish_row_load_inputs<NV>(a.x[m] + row * ldx, a.x[1 - m] + row * ldx, xv, pv, lane);
    const float s = expf(*ls);
    tail_terms(a.bias, diag[row], rs, cs, row, s, (*grad_out) * a.emb_scale, batch, coef, dterm);
    idx_ = a.inv_den[m][row];
    idp = a.inv_den[1 - m][row];
    clamped = !(a.nrm[m][row] > kNormEps);
    if constexpr (BOTH) clamped_p = !(a.nrm[1 - m][row] > kNormEps);
  }
  pdl_wait();
  if (blockIdx.x == 0 && m == 0 && threadIdx.x < 32 && dls_out != nullptr) {
    float dls_local = 0.f;
    if (threadIdx.x == 0) {
      if (a.bias != nullptr) {
        const double g_b = (double)(*grad_out_dls) / (double)batch;
        dls_local = (float)(g_b * ((double)gs[0] + a.sig_sums[1]));
        *a.dbias_out = (float)(g_b * ((double)gs[1] + a.sig_sums[2]));
        gs[1] = 0.f;
      } else {
        dls_local = (float)((double)(*grad_out_dls) / (2.0 * (double)batch) * ((double)(*gs) - 2.0 * (double)(*diag_sum)));
      }
      *dls_out = dls_local;
      *gs = 0.f;   // consumed: the accumulator is back to its zero-initialised state
    }
    if (xg.peer != nullptr) xgpu_publish(xg, threadIdx.x == 0 ? *xg.loss_partial : 0.f, dls_local);
  }
  if (xg.peer != nullptr && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && threadIdx.x < 32 &&
      dls_out != nullptr)
    xgpu_collect(xg);
  if (row >= n) return;
  if constexpr (BOTH) {
    float4 xo[NV], po[NV];   // finish_row_vec_loaded scales its own row in place: the partner's tail needs the raw rows
#pragma unroll
    for (int i = 0; i < NV; ++i) { xo[i] = pv[i]; po[i] = xv[i]; }
    finish_row_vec_loaded<NV>(a.acc[0] + row * d, parts, n * d / 4, xv, pv, coef, dterm, idx_, idp, clamped,
                              a.dx[0] + row * d, lane);
    finish_row_vec_loaded<NV>(a.acc[1] + row * d, parts, n * d / 4, xo, po, coef, dterm, idp, idx_, clamped_p,
                              a.dx[1] + row * d, lane);
  } else {
    finish_row_vec_loaded<NV>(a.acc[m] + row * d, parts, n * d / 4, xv, pv, coef, dterm, idx_, idp, clamped,
                              a.dx[m] + row * d, lane);
  }
}

template <typename TI>
static int grad_finish_dispatch(const float* acc, int parts, const TI* x, const TI* p, int64_t n, int64_t d,
                                int64_t ldx, const float* idx_, const float* nrm, const float* idp,
                                const float* diag, const float* rs, const float* cs, const float* ls, const float* go, int64_t batch, void* dx,
                                int dx_dtype, cudaStream_t st, const float* bias = nullptr) {
  dim3 block(256), grid((unsigned)ceil_div(n, 8));
  if constexpr (sizeof(TI) == 4) {
    if (dx_dtype == PLK_F32 && d % 128 == 0 && d <= 1024 && (ldx & 3) == 0 && (((uintptr_t)x | (uintptr_t)p | (uintptr_t)acc | (uintptr_t)dx) & 15) == 0) {
      switch (d / 128) {
#define PLK_CASE(NV) case NV: grad_finish_vec_kernel<NV><<<grid, block, 0, st>>>(acc, parts, (const float*)x, (const float*)p, n, ldx, idx_, nrm, idp, diag, rs, cs, ls, go, batch, (float*)dx, bias); break;
        PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8)
#undef PLK_CASE
      }
      PLK_LAUNCHED(1);
      return PLK_OK;
    }
  }
  if (dx_dtype == PLK_F32)
    grad_finish_kernel<TI, float><<<grid, block, 0, st>>>(acc, parts, x, p, n, d, ldx, idx_, nrm, idp, diag, rs, cs, ls, go, batch, (float*)dx, bias);
  else if (dx_dtype == PLK_BF16)
    grad_finish_kernel<TI, __nv_bfloat16><<<grid, block, 0, st>>>(acc, parts, x, p, n, d, ldx, idx_, nrm, idp, diag, rs, cs, ls, go, batch, (__nv_bfloat16*)dx, bias);
  else
    grad_finish_kernel<TI, __half><<<grid, block, 0, st>>>(acc, parts, x, p, n, d, ldx, idx_, nrm, idp, diag, rs, cs, ls, go, batch, (__half*)dx, bias);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

// ---------------------------------------------------------------------------------------------
// Retrieval tails.  Lists are tiny (<= a few hundred entries per query): one thread per query
// does an insertion selection ordered by (key, index).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxK = 64;

__device__ __forceinline__ bool pair_less(float ka, int32_t ia, float kb, int32_t ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// keep the k smallest (key, idx) pairs of a stream; best[] sorted ascending.
__device__ __forceinline__ void topk_insert(float* bk, int32_t* bi, int k, float key, int32_t idx) {
  if (!pair_less(key, idx, bk[k - 1], bi[k - 1])) return;
  int p = k - 1;
  while (p > 0 && pair_less(key, idx, bk[p - 1], bi[p - 1])) {
    bk[p] = bk[p - 1];
    bi[p] = bi[p - 1];
    --p;
  }
  bk[p] = key;
  bi[p] = idx;
}

__global__ void __launch_bounds__(128) select_kernel(const int32_t* __restrict__ in_idx,
                                                     const float* __restrict__ in_key, int64_t nq,
                                                     int m, int k, int32_t* __restrict__ out_idx,
                                                     float* __restrict__ out_key) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float bk[kMaxK];
  int32_t bi[kMaxK];
  for (int t = 0; t < k; ++t) { bk[t] = CUDART_INF_F; bi[t] = 0x7fffffff; }
  for (int t = 0; t < m; ++t) {
    int32_t id = in_idx[q * m + t];
    if (id < 0) continue;
    topk_insert(bk, bi, k, in_key[q * m + t], id);
  }
  for (int t = 0; t < k; ++t) {
    const bool empty = bi[t] == 0x7fffffff;
    out_idx[q * k + t] = empty ? -1 : bi[t];
    out_key[q * k + t] = bk[t];
  }
}

// One small kernel instead of several memset nodes (a memset node costs more than a kernel node
// inside a CUDA graph and this path is launch-bound at the reference's batch sizes).
__global__ void __launch_bounds__(256) zero2_kernel(float* __restrict__ a, int64_t na, float* __restrict__ b, int64_t nb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) a[i] = 0.f;
  if (i < nb) b[i] = 0.f;
}
int zero2(float* a, int64_t na, float* b, int64_t nb, cudaStream_t st) {
  const int64_t n = na > nb ? na : nb;
  zero2_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(a, na, b, nb);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int select_candidates(const int32_t* in_idx, const float* in_key, int64_t nq, int m, int kc,
                      int32_t* out_idx, float* out_key, cudaStream_t st) {
  select_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(in_idx, in_key, nq, m, kc, out_idx, out_key);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

// exact distance of each candidate: one warp per (query, candidate), fp64 accumulation.
__global__ void __launch_bounds__(256) rescore_dist_kernel(const float* __restrict__ q32,
                                                           const float* __restrict__ g32, int64_t nq,
                                                           int64_t ng, int64_t d,
                                                           const int32_t* __restrict__ cand, int m,
                                                           int64_t goff, float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nq * m) return;
  const int64_t qi = w / m;
  const int32_t id = cand[w];
  const int64_t lr = (int64_t)id - goff;
  if (id < 0 || lr < 0 || lr >= ng) {
    if (lane == 0) dist[w] = CUDART_INF_F;
    return;
  }
  const float* qr = q32 + qi * d;
  const float* gr = g32 + lr * d;
  double a = 0.0;
  for (int64_t k = lane; k < d; k += 32) {
    double t = (double)qr[k] - (double)gr[k];
    a = fma(t, t, a);
  }
  a = warp_sum(a);
  if (lane == 0) dist[w] = (float)sqrt(a);
}

// ---------------------------------------------------------------------------------------------
// a13/a14: inverse-distance weighted vote                            reference src/ann.py:19-34
// ---------------------------------------------------------------------------------------------
constexpr int kMaxVote = 256;
__global__ void __launch_bounds__(128) vote_kernel(const int32_t* __restrict__ idx,
                                                   const float* __restrict__ dist, int64_t nq, int m,
                                                   const int64_t* __restrict__ labels, int64_t ng,
                                                   int64_t* __restrict__ pred) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int32_t* ir = idx + q * m;
  const float* dr = dist + q * m;
  bool any_zero = false;
  for (int t = 0; t < m; ++t)
    if (ir[t] >= 0 && dr[t] == 0.0f) any_zero = true;
  double best_w = 0.0;
  int64_t best_c = 0;
  bool have = false;
  for (int t = 0; t < m; ++t) {
    if (ir[t] < 0) continue;
    const int64_t c = labels[ir[t]];
    bool first = true;  // evaluate each distinct class once (at its first occurrence)
    for (int r = 0; r < t; ++r)
      if (ir[r] >= 0 && labels[ir[r]] == c) { first = false; break; }
    if (!first) continue;
    double tot = 0.0;
    for (int r = t; r < m; ++r) {
      if (ir[r] < 0 || labels[ir[r]] != c) continue;
      float w = any_zero ? (dr[r] == 0.0f ? 1.0f : 0.0f) : 1.0f / dr[r];
      tot += (double)w;
    }
    // weighted_mode visits classes in ascending order with a strict '>' starting from a zero
    // count, i.e. the winner is the LOWEST class id among those with the largest positive sum.
    if (tot > best_w || (have && tot == best_w && c < best_c)) {
      best_w = tot;
      best_c = c;
      have = true;
    }
  }
  pred[q] = have ? best_c : 0;
}

}  // namespace plk

using namespace plk;

extern "C" {

int plk_l2norm_fwd(const void* x, int x_dtype, int64_t n, int64_t d, int64_t ldx, void* u,
                   int u_dtype, int64_t ldu, float* inv_den, float* nrm, float* sqn, int normalise,
                   void* stream) {
  PLK_REQUIRE(x && u, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n > 0 && d > 0 && ldx >= d && ldu >= d, PLK_ERR_INVALID, "bad shape n=%lld d=%lld ldx=%lld ldu=%lld",
              (long long)n, (long long)d, (long long)ldx, (long long)ldu);
  PLK_REQUIRE(u_dtype == PLK_F32 || u_dtype == PLK_BF16 || u_dtype == PLK_F16, PLK_ERR_INVALID, "u_dtype must be F32, BF16 or F16");
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case PLK_F32: return l2norm_dispatch_out((const float*)x, n, d, ldx, u, u_dtype, ldu, inv_den, nrm, sqn, normalise, st);
    case PLK_BF16: return l2norm_dispatch_out((const __nv_bfloat16*)x, n, d, ldx, u, u_dtype, ldu, inv_den, nrm, sqn, normalise, st);
    case PLK_F16: return l2norm_dispatch_out((const __half*)x, n, d, ldx, u, u_dtype, ldu, inv_den, nrm, sqn, normalise, st);
  }
  set_error("bad x_dtype %d", x_dtype);
  return PLK_ERR_INVALID;
}

int plk_l2norm_pair_fwd(const float* x, const float* y, int64_t n, int64_t d, int64_t ldx, void* u, void* v,
                        int u_dtype, int64_t ldu, float* inv_den_x, float* nrm_x, float* inv_den_y,
                        float* nrm_y, float* zero_a, int64_t n_zero_a, float* zero_b, int64_t n_zero_b,
                        void* stream) {
  PLK_REQUIRE(x && y && u && v && inv_den_x && nrm_x && inv_den_y && nrm_y, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n > 0 && d > 0 && ldx >= d && ldu >= d, PLK_ERR_INVALID, "bad shape");
  PLK_REQUIRE(u_dtype >= PLK_F32 && u_dtype <= PLK_F16, PLK_ERR_INVALID, "bad u_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = d % 128 == 0 && d <= 1024 && ldu == d && (ldx & 3) == 0 &&
                      ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)u | (uintptr_t)v) & 15) == 0);
  if (vec_ok) {
    NormPairArgs a;
    a.x[0] = x; a.x[1] = y; a.u[0] = u; a.u[1] = v;
    a.inv_den[0] = inv_den_x; a.inv_den[1] = inv_den_y; a.nrm[0] = nrm_x; a.nrm[1] = nrm_y;
    a.zero[0] = zero_a; a.zero[1] = zero_b; a.nzero[0] = zero_a ? n_zero_a : 0; a.nzero[1] = zero_b ? n_zero_b : 0;
    const bool ok = u_dtype == PLK_F32 ? l2norm_pair_try<float>(a, n, d, ldx, ldu, st)
                    : u_dtype == PLK_BF16 ? l2norm_pair_try<__nv_bfloat16>(a, n, d, ldx, ldu, st)
                                          : l2norm_pair_try<__half>(a, n, d, ldx, ldu, st);
    if (ok) {
      PLK_LAUNCHED(1);
      return PLK_OK;
    }
  }
  int rc = plk_l2norm_fwd(x, PLK_F32, n, d, ldx, u, u_dtype, ldu, inv_den_x, nrm_x, nullptr, 1, stream);
  if (rc) return rc;
  rc = plk_l2norm_fwd(y, PLK_F32, n, d, ldx, v, u_dtype, ldu, inv_den_y, nrm_y, nullptr, 1, stream);
  if (rc) return rc;
  if (zero_a || zero_b) return zero2(zero_a, zero_a ? n_zero_a : 0, zero_b, zero_b ? n_zero_b : 0, st);
  return PLK_OK;
}

int plk_infonce_loss(const float* row_sumexp, const float* col_sumexp_own, const float* diag,
                     const float* logit_scale, int64_t n_rows, int64_t batch_global, float* loss_out,
                     float* diag_sum_out, float* gs_zero, void* stream) {
  PLK_REQUIRE(row_sumexp && col_sumexp_own && diag && logit_scale && loss_out, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n_rows > 0 && batch_global >= n_rows, PLK_ERR_INVALID, "bad sizes");
  XGpuArgs xg = {};
  PLK_CUDA(launch_overlapped(loss_kernel, dim3(1), dim3(1024), (cudaStream_t)stream, row_sumexp, col_sumexp_own, diag,
                             logit_scale, n_rows, batch_global, loss_out, diag_sum_out, gs_zero, xg, (float*)nullptr));
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int plk_infonce_loss_xgpu(const float* row_sumexp, const float* col_sumexp_own, const float* diag,
                          const float* logit_scale, int64_t n_rows, int64_t batch_global, float* loss_out,
                          float* diag_sum_out, float* gs_zero, float* partial_out, void* const* peer_bufs,
                          int rank, int world, unsigned* epoch, float* out2, void* stream) {
  PLK_REQUIRE(row_sumexp && col_sumexp_own && diag && logit_scale && loss_out && partial_out && peer_bufs && epoch &&
                  out2, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n_rows > 0 && batch_global >= n_rows, PLK_ERR_INVALID, "bad sizes");
  PLK_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, PLK_ERR_INVALID,
              "world must be in [2, 8] (got rank %d of %d)", rank, world);
  XGpuArgs xg;
  xg.peer = peer_bufs; xg.rank = rank; xg.world = world; xg.epoch = epoch; xg.loss_partial = nullptr; xg.out2 = out2;
  xg.timeout = xgpu_timeout_cycles();
  PLK_CUDA(launch_overlapped(loss_kernel, dim3(1), dim3(1024), (cudaStream_t)stream, row_sumexp, col_sumexp_own, diag,
                             logit_scale, n_rows, batch_global, loss_out, diag_sum_out, gs_zero, xg, partial_out));
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int plk_infonce_dls(const float* gs, const float* diag_sum, const float* grad_out,
                    int64_t batch_global, float* dls_out, void* stream) {
  PLK_REQUIRE(gs && diag_sum && grad_out && dls_out && batch_global > 0, PLK_ERR_INVALID, "bad args");
  dls_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(gs, diag_sum, grad_out, batch_global, dls_out);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int plk_infonce_grad_finish(const float* acc, int parts, const void* x, const void* partner, int x_dtype,
                            int64_t n, int64_t d, int64_t ldx, const float* inv_den_x,
                            const float* nrm_x, const float* inv_den_p, const float* diag,
                            const float* rs, const float* cs, const float* logit_scale,
                            const float* grad_out, int64_t batch_global, void* dx, int dx_dtype,
                            void* stream) {
  PLK_REQUIRE(acc && x && partner && inv_den_x && nrm_x && inv_den_p && diag && rs && cs && logit_scale && grad_out && dx,
              PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n > 0 && d > 0 && ldx >= d && batch_global >= n && parts >= 1, PLK_ERR_INVALID, "bad sizes");
  PLK_REQUIRE(dx_dtype >= PLK_F32 && dx_dtype <= PLK_F16, PLK_ERR_INVALID, "bad dx_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case PLK_F32: return grad_finish_dispatch(acc, parts, (const float*)x, (const float*)partner, n, d, ldx, inv_den_x, nrm_x, inv_den_p, diag, rs, cs, logit_scale, grad_out, batch_global, dx, dx_dtype, st);
    default: break;
  }
  set_error("grad_finish: raw embeddings must be fp32 (got dtype %d); cast on the host side", x_dtype);
  return PLK_ERR_UNSUPPORTED;
}

static int finish_pair_impl(const float* acc_x, const float* acc_y, int parts, const float* x,
                                 const float* y, int64_t n, int64_t d, int64_t ldx,
                                 const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                                 const float* nrm_y, const float* diag, const float* rs, const float* cs,
                                 const float* logit_scale, const float* grad_out_emb,
                                 const float* grad_out, int64_t batch_global, float* gs,
                                 const float* diag_sum, float* dx, float* dy, float* dls_out,
                                 const XGpuArgs& xg, void* stream, float emb_scale = 1.0f,
                                 const float* bias = nullptr, const double* sig_sums = nullptr,
                                 float* dbias_out = nullptr) {
  PLK_REQUIRE(acc_x && acc_y && x && y && inv_den_x && nrm_x && inv_den_y && nrm_y && diag &&
                  logit_scale && grad_out_emb && grad_out && gs && dx && dy && dls_out,
              PLK_ERR_INVALID, "null pointer");
  if (bias != nullptr) PLK_REQUIRE(sig_sums && dbias_out, PLK_ERR_INVALID, "null pointer");
  else PLK_REQUIRE(rs && cs && diag_sum, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(n > 0 && d > 0 && ldx >= d && batch_global >= n && parts >= 1, PLK_ERR_INVALID, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = d % 128 == 0 && d <= 1024 && (ldx & 3) == 0 &&
                      ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)acc_x | (uintptr_t)acc_y | (uintptr_t)dx | (uintptr_t)dy) & 15) == 0);
  if (vec_ok) {
    FinishPairArgs a;
    a.acc[0] = acc_x; a.acc[1] = acc_y; a.x[0] = x; a.x[1] = y;
    a.inv_den[0] = inv_den_x; a.inv_den[1] = inv_den_y; a.nrm[0] = nrm_x; a.nrm[1] = nrm_y;
    a.dx[0] = dx; a.dx[1] = dy;
    a.emb_scale = emb_scale;
    a.bias = bias; a.sig_sums = sig_sums; a.dbias_out = dbias_out;
    const bool both = d <= 512;   // one warp per row PAIR (see grad_finish_pair_vec_kernel)
    dim3 block(256), grid((unsigned)ceil_div(n, 8), both ? 1 : 2);
    switch (d / 128) {
#define PLK_CASE(NV, BOTH) case NV: PLK_CUDA(launch_overlapped(grad_finish_pair_vec_kernel<NV, BOTH>, grid, block, st, a, parts, n, ldx, diag, rs, cs, logit_scale, grad_out_emb, grad_out, batch_global, gs, diag_sum, dls_out, xg)); break;
      PLK_CASE(1, true) PLK_CASE(2, true) PLK_CASE(3, true) PLK_CASE(4, true)
      PLK_CASE(5, false) PLK_CASE(6, false) PLK_CASE(7, false) PLK_CASE(8, false)
#undef PLK_CASE
    }
    PLK_LAUNCHED(1);
    return PLK_OK;
  }
  PLK_REQUIRE(xg.peer == nullptr && emb_scale == 1.0f, PLK_ERR_UNSUPPORTED,
              "the fused cross-GPU scalar exchange / emb_scale need d % 128 == 0 (d <= 1024) and 16-byte aligned rows");
  if (bias != nullptr) {   // SigLIP, ragged width: one tail per modality + the scalars
    int rc = grad_finish_dispatch(acc_x, parts, x, y, n, d, ldx, inv_den_x, nrm_x, inv_den_y, diag, rs, cs, logit_scale,
                                  grad_out_emb, batch_global, dx, PLK_F32, st, bias);
    if (rc) return rc;
    rc = grad_finish_dispatch(acc_y, parts, y, x, n, d, ldx, inv_den_y, nrm_y, inv_den_x, diag, rs, cs, logit_scale,
                              grad_out_emb, batch_global, dy, PLK_F32, st, bias);
    if (rc) return rc;
    siglip_scalars_kernel<<<1, 1, 0, st>>>(gs, sig_sums, grad_out, batch_global, dls_out, dbias_out);
    PLK_LAUNCHED(1);
    return PLK_OK;
  }
  int rc = plk_infonce_grad_finish(acc_x, parts, x, y, PLK_F32, n, d, ldx, inv_den_x, nrm_x, inv_den_y, diag, rs, cs,
                                   logit_scale, grad_out_emb, batch_global, dx, PLK_F32, stream);
  if (rc) return rc;
  rc = plk_infonce_grad_finish(acc_y, parts, y, x, PLK_F32, n, d, ldx, inv_den_y, nrm_y, inv_den_x, diag, rs, cs,
                               logit_scale, grad_out_emb, batch_global, dy, PLK_F32, stream);
  if (rc) return rc;
  rc = plk_infonce_dls(gs, diag_sum, grad_out, batch_global, dls_out, stream);
  if (rc) return rc;
  PLK_CUDA(cudaMemsetAsync(gs, 0, sizeof(float), st));
  return PLK_OK;
}

int plk_infonce_grad_finish_pair(const float* acc_x, const float* acc_y, int parts, const float* x,
                                 const float* y, int64_t n, int64_t d, int64_t ldx,
                                 const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                                 const float* nrm_y, const float* diag, const float* rs, const float* cs,
                                 const float* logit_scale, const float* grad_out_emb,
                                 const float* grad_out, int64_t batch_global, float* gs,
                                 const float* diag_sum, float* dx, float* dy, float* dls_out,
                                 void* stream) {
  XGpuArgs xg = {};
  return finish_pair_impl(acc_x, acc_y, parts, x, y, n, d, ldx, inv_den_x, nrm_x, inv_den_y, nrm_y, diag, rs, cs,
                          logit_scale, grad_out_emb, grad_out, batch_global, gs, diag_sum, dx, dy, dls_out, xg,
                          stream);
}

int plk_infonce_grad_finish_pair_xgpu(const float* acc_x, const float* acc_y, int parts, const float* x,
                                      const float* y, int64_t n, int64_t d, int64_t ldx,
                                      const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                                      const float* nrm_y, const float* diag, const float* rs, const float* cs,
                                      const float* logit_scale, const float* grad_out_emb,
                                      const float* grad_out, int64_t batch_global, float* gs,
                                      const float* diag_sum, float* dx, float* dy, float* dls_out,
                                      const float* loss_partial, void* const* peer_bufs, int rank, int world,
                                      unsigned* epoch, float* out2, void* stream) {
  PLK_REQUIRE(loss_partial && peer_bufs && epoch && out2, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, PLK_ERR_INVALID,
              "world must be in [2, 8] (got rank %d of %d)", rank, world);
  XGpuArgs xg;
  xg.peer = peer_bufs; xg.rank = rank; xg.world = world; xg.epoch = epoch;
  xg.loss_partial = loss_partial; xg.out2 = out2;
  xg.timeout = xgpu_timeout_cycles();
  return finish_pair_impl(acc_x, acc_y, parts, x, y, n, d, ldx, inv_den_x, nrm_x, inv_den_y, nrm_y, diag, rs, cs,
                          logit_scale, grad_out_emb, grad_out, batch_global, gs, diag_sum, dx, dy, dls_out, xg,
                          stream);
}

int plk_siglip_loss(const double* sums, int64_t batch, float* loss_out, void* stream) {
  PLK_REQUIRE(sums && loss_out && batch > 0, PLK_ERR_INVALID, "bad args");
  PLK_CUDA(launch_overlapped(siglip_loss_kernel, dim3(1), dim3(1), (cudaStream_t)stream, sums, batch, loss_out));
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int plk_siglip_grad_finish_pair(const float* acc_x, const float* acc_y, int parts, const float* x, const float* y,
                                int64_t n, int64_t d, int64_t ldx, const float* inv_den_x, const float* nrm_x,
                                const float* inv_den_y, const float* nrm_y, const float* diag,
                                const float* logit_scale, const float* bias, const float* grad_out,
                                int64_t batch, float* gs2, const double* sums, float* dx, float* dy,
                                float* dls_out, float* dbias_out, void* stream) {
  PLK_REQUIRE(bias != nullptr, PLK_ERR_INVALID, "null pointer");
  XGpuArgs xg = {};
  return finish_pair_impl(acc_x, acc_y, parts, x, y, n, d, ldx, inv_den_x, nrm_x, inv_den_y, nrm_y, diag, nullptr,
                          nullptr, logit_scale, grad_out, grad_out, batch, gs2, nullptr, dx, dy, dls_out, xg, stream,
                          1.0f, bias, sums, dbias_out);
}

}  // extern "C"

namespace plk {
// plk_infonce_grad_finish_pair(_xgpu) with an extra host-side factor on the embedding gradients
// (peer_bufs == nullptr: no cross-GPU exchange).  Used by the composite backward (api.cu).
int finish_pair_scaled(const float* acc_x, const float* acc_y, int parts, const float* x, const float* y, int64_t n,
                       int64_t d, int64_t ldx, const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                       const float* nrm_y, const float* diag, const float* rs, const float* cs,
                       const float* logit_scale, const float* grad_out_emb, float emb_scale, const float* grad_out,
                       int64_t batch_global, float* gs, const float* diag_sum, float* dx, float* dy, float* dls_out,
                       const float* loss_partial, void* const* peer_bufs, int rank, int world, unsigned* epoch,
                       float* out2, void* stream) {
  XGpuArgs xg = {};
  if (peer_bufs != nullptr) {
    PLK_REQUIRE(loss_partial && epoch && out2, PLK_ERR_INVALID, "null pointer");
    PLK_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, PLK_ERR_INVALID,
                "world must be in [2, 8] (got rank %d of %d)", rank, world);
    xg.peer = peer_bufs; xg.rank = rank; xg.world = world; xg.epoch = epoch;
    xg.loss_partial = loss_partial; xg.out2 = out2;
    xg.timeout = xgpu_timeout_cycles();
  }
  return finish_pair_impl(acc_x, acc_y, parts, x, y, n, d, ldx, inv_den_x, nrm_x, inv_den_y, nrm_y, diag, rs, cs,
                          logit_scale, grad_out_emb, grad_out, batch_global, gs, diag_sum, dx, dy, dls_out, xg, stream,
                          emb_scale);
}
}  // namespace plk

extern "C" {

int plk_topk_rescore(const float* q32, const float* g32, int64_t nq, int64_t ng, int64_t d,
                     const int32_t* cand_idx, int m, int64_t gallery_offset, int k, float* scratch,
                     int32_t* out_idx, float* out_dist, void* stream) {
  PLK_REQUIRE(q32 && g32 && cand_idx && scratch && out_idx && out_dist, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(nq > 0 && ng > 0 && d > 0 && m >= 1 && k >= 1 && k <= kMaxK, PLK_ERR_INVALID, "bad sizes (k<=64)");
  cudaStream_t st = (cudaStream_t)stream;
  rescore_dist_kernel<<<(unsigned)ceil_div(nq * m, 8), 256, 0, st>>>(q32, g32, nq, ng, d, cand_idx, m, gallery_offset, scratch);
  PLK_LAUNCHED(1);
  return select_candidates(cand_idx, scratch, nq, m, k, out_idx, out_dist, st);
}

int plk_topk_merge(const int32_t* cand_idx, const float* cand_dist, int64_t nq, int m, int k,
                   int32_t* out_idx, float* out_dist, void* stream) {
  PLK_REQUIRE(cand_idx && cand_dist && out_idx && out_dist, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(nq > 0 && m >= 1 && k >= 1 && k <= kMaxK, PLK_ERR_INVALID, "bad sizes (k<=64)");
  return select_candidates(cand_idx, cand_dist, nq, m, k, out_idx, out_dist, (cudaStream_t)stream);
}

int plk_knn_vote(const int32_t* idx, const float* dist, int64_t nq, int m, const int64_t* labels,
                 int64_t ng, int64_t* pred, void* stream) {
  PLK_REQUIRE(idx && dist && labels && pred, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(nq > 0 && m >= 1 && m <= kMaxVote && ng > 0, PLK_ERR_INVALID, "bad sizes (m<=256)");
  vote_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, (cudaStream_t)stream>>>(idx, dist, nq, m, labels, ng, pred);
  PLK_LAUNCHED(1);
  return PLK_OK;
}

}  // extern "C"
