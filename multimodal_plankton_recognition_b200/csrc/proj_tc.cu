// N1: the bias-free projection Linear of one modality fused with the L2 normalisation
// that opens the loss (reference src/model.py:29-30,:38-39,:80-83 followed by src/coordination.py:33-34).
//
//   emb  = feat . W^T            [n, f] x [d, f]^T, 16-bit operands, fp32 accumulation in TMEM
//   u    = emb / max(||emb||, eps)   written as the 16-bit, zero-padded operand of the similarity kernels
// in ONE kernel: a CTA owns 128 rows and the whole output width (d <= 512 = all 512 TMEM columns), streams
// [128 x 64] chunks of the features and [d x 64] chunks of the weight through a TMA ring (SS-mode
// tcgen05.mma, N = d up to 256 per instruction), and its epilogue -- thread = one output row -- takes the
// squared norm straight from the accumulator, then writes the raw fp32 embedding (the gradient tail and
// CLIPPlus' MSE term read it), the normalised operand and the (1/den, ||.||) statistics.  The [n, d]
// embedding is never re-read to be normalised and the separate normalisation launch disappears.
#include <cstdlib>

#include "tc_common.cuh"

namespace plk {
using namespace tc;

constexpr int kPjThreads = 576;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int kPjEpi = 512;
constexpr int kPjAux = 4096;      // barriers + tmem pointer (first 512 B), [4][128] partial squared norms

template <int ND>
struct ProjCfg {
  static constexpr int kN = ND * 64;                            // padded output width
  static constexpr int kStageBytes = kChunkBytes + kN * 128;    // feature chunk + weight chunk
  static constexpr int kStagesMax = (kMaxSmem - 1024 - kPjAux) / kStageBytes;
  static constexpr int kStages = kStagesMax > 4 ? 4 : kStagesMax;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kPjAux;
  static_assert(kStages >= 2, "not enough shared memory for the ring");
};

template <int ND>
__global__ void __launch_bounds__(kPjThreads, 1) proj_norm_tc(
    const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
    const __grid_constant__ CUtensorMap tmap_w2, int64_t n, int64_t d, int kchunks, int f16, void* __restrict__ u16, int64_t ldu, float* __restrict__ emb, float* __restrict__ inv_den,
    float* __restrict__ nrm_out) {
  using Cfg = ProjCfg<ND>;
  constexpr int NST = Cfg::kStages;
  constexpr int N = Cfg::kN;
  constexpr int N1 = N > 256 ? 256 : N, N2 = N - N1;       // one or two MMAs per K step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_ring = smem;
  uint8_t* aux = sm_ring + NST * Cfg::kStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);   // [NST]
  uint64_t* bar_empty = bar_full + NST;                    // [NST]
  uint64_t* bar_acc = bar_empty + NST;                     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* ss_s = reinterpret_cast<float*>(aux + 512);       // [4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * kTileRows;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(bar_empty + st, ph ^ 1);
        mbar_expect_tx(bar_full + st, Cfg::kStageBytes);
        uint8_t* slot = sm_ring + st * Cfg::kStageBytes;
        tma_load_2d(slot, &tmap_x, bar_full + st, kc * kChunkK, (int)i0);                       // [128 x 64] features
        tma_load_2d(slot + kChunkBytes, &tmap_w, bar_full + st, kc * kChunkK, 0);              // weight rows 0..N1-1
        if constexpr (N2 > 0)
          tma_load_2d(slot + kChunkBytes + N1 * 128, &tmap_w2, bar_full + st, kc * kChunkK, N1);  // rows N1..N-1 (box N2)
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    const uint32_t idesc1 = umma_idesc_16(128, N1, 0, 0, f16);
    const uint32_t idesc2 = umma_idesc_16(128, N2 > 0 ? N2 : 16, 0, 0, f16);
    const uint32_t r_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);
    int st = 0; uint32_t ph = 0;
    for (int kc = 0; kc < kchunks; ++kc) {
      mbar_wait(bar_full + st, ph);
      tc_fence_after();
      const uint32_t a_lo = r_lo0 + st * (Cfg::kStageBytes >> 4);
      const uint32_t b_lo = a_lo + (kChunkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kChunkK / kUmmaK; ++k) {
          umma_bf16_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc1, (kc | k) != 0);
          if constexpr (N2 > 0)
            umma_bf16_lo(tmem_base + N1, a_lo + 2 * k, b_lo + ((N1 * 128) >> 4) + 2 * k, idesc2, (kc | k) != 0);
        }
        umma_commit(bar_empty + st);
        if (kc == kchunks - 1) umma_commit(bar_acc);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int e = (warp - 2) >> 2;           // the four warps of a quadrant take alternate 32-column chunks
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    float ss = 0.f;
#pragma unroll 1
    for (int ch = e; ch < 2 * ND; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + ch * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) ss = fmaf(__uint_as_float(raw[x]), __uint_as_float(raw[x]), ss);
    }
    ss_s[e * 128 + r] = ss;
    named_barrier_sync(1, kPjEpi);
    const float tot = (ss_s[r] + ss_s[128 + r]) + (ss_s[256 + r] + ss_s[384 + r]);
    const float nrm = sqrtf(tot);
    const float den = fmaxf(nrm, kNormEps);
    if (e == 0 && i < n) {
      inv_den[i] = 1.0f / den;
      nrm_out[i] = nrm;
    }
#pragma unroll 1
    for (int ch = e; ch < 2 * ND; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + ch * 32, raw);
      tmem_ld_wait();
      if (i >= n) continue;
      const int64_t col0 = (int64_t)ch * 32;
      float* dst = emb + i * d + col0;
      if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
        for (int x = 0; x < 32; x += 4)
          *reinterpret_cast<float4*>(dst + x) = make_float4(__uint_as_float(raw[x]), __uint_as_float(raw[x + 1]),
                                                            __uint_as_float(raw[x + 2]), __uint_as_float(raw[x + 3]));
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x)
          if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
      }
      // normalised 16-bit operand, zero padded to ldu (weight rows past d are zero-filled by the TMA unit)
      uint32_t pk[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float lo = __uint_as_float(raw[2 * x]) / den, hi = __uint_as_float(raw[2 * x + 1]) / den;
        pk[x] = f16 ? pack_16x2<true>(lo, hi) : pack_16x2<false>(lo, hi);
      }
      uint4* ud = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(u16) + i * ldu + col0);
#pragma unroll
      for (int x = 0; x < 4; ++x) ud[x] = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Output width split over a cluster of two CTAs (padded d = 128 / 256 / 512): CTA `rank` of the pair owns the
// same 128 rows and the output columns [rank * d/2, (rank + 1) * d/2).  Twice as many CTAs (a batch of 4096 rows
// is only 32 row blocks on 148 SMs), 8 KiB instead of 12 per K step through the shared-memory operand port,
// half the epilogue per CTA.  The squared norm of a row is the sum of the two halves: each CTA stores its 128
// partial sums into the PEER's shared memory (st.shared::cluster) and both meet at one cluster barrier.
// ---------------------------------------------------------------------------------------------
template <int NH>
struct Proj2Cfg {
  static constexpr int kN = NH * 64;                            // output columns per CTA
  static constexpr int kStageBytes = kChunkBytes + kN * 128;
  static constexpr int kStagesMax = (kMaxSmem - 1024 - kPjAux) / kStageBytes;
  static constexpr int kStages = kStagesMax > 6 ? 6 : kStagesMax;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kPjAux;
};

template <int NH>
__global__ void __launch_bounds__(kPjThreads, 1) proj_norm_tc2(
    const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, int64_t n, int64_t d,
    int kchunks, int f16, void* __restrict__ u16, int64_t ldu, float* __restrict__ emb, float* __restrict__ inv_den,
    float* __restrict__ nrm_out) {
  using Cfg = Proj2Cfg<NH>;
  constexpr int NST = Cfg::kStages;
  constexpr int N = Cfg::kN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_ring = smem;
  uint8_t* aux = sm_ring + NST * Cfg::kStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);   // [NST]
  uint64_t* bar_empty = bar_full + NST;                    // [NST]
  uint64_t* bar_acc = bar_empty + NST;                     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);
  float* ss_s = reinterpret_cast<float*>(aux + 512);       // [4][128] partial sums of this CTA's warps
  float* ss_peer = ss_s + 512;                             // [128] written by the peer CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t i0 = (int64_t)(blockIdx.x >> 1) * kTileRows;
  const int col_base = (int)rank * N;                      // first output column of this CTA
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(bar_empty + st, ph ^ 1);
        mbar_expect_tx(bar_full + st, Cfg::kStageBytes);
        uint8_t* slot = sm_ring + st * Cfg::kStageBytes;
        tma_load_2d(slot, &tmap_x, bar_full + st, kc * kChunkK, (int)i0);                 // [128 x 64] features
        tma_load_2d(slot + kChunkBytes, &tmap_w, bar_full + st, kc * kChunkK, col_base);  // this CTA's weight rows
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_16(128, N, 0, 0, f16);
    const uint32_t r_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);
    int st = 0; uint32_t ph = 0;
    for (int kc = 0; kc < kchunks; ++kc) {
      mbar_wait(bar_full + st, ph);
      tc_fence_after();
      const uint32_t a_lo = r_lo0 + st * (Cfg::kStageBytes >> 4);
      const uint32_t b_lo = a_lo + (kChunkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kChunkK / kUmmaK; ++k)
          umma_bf16_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc, (kc | k) != 0);
        umma_commit(bar_empty + st);
        if (kc == kchunks - 1) umma_commit(bar_acc);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  }
  const int q = warp & 3;
  const int e = (warp - 2) >> 2;
  const int r = q * 32 + lane;
  const int64_t i = i0 + r;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  if (warp >= 2) {
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    float ss = 0.f;
#pragma unroll 1
    for (int ch = e; ch < 2 * NH; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + ch * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) ss = fmaf(__uint_as_float(raw[x]), __uint_as_float(raw[x]), ss);
    }
    ss_s[e * 128 + r] = ss;
    named_barrier_sync(1, kPjEpi);
    if (e == 0) {   // this CTA's half of the squared norm -> the peer
      const float half = (ss_s[r] + ss_s[128 + r]) + (ss_s[256 + r] + ss_s[384 + r]);
      ss_s[r] = half;
      asm volatile(
          "{\n\t.reg .b32 ra;\n\t"
          "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
          "st.shared::cluster.f32 [ra], %2;\n\t}"
          ::"r"(smem_u32(ss_peer + r)), "r"(rank ^ 1u), "f"(half)
          : "memory");
    }
  }
  cluster_sync_all();   // release / acquire: the peer's partial sums are visible (every thread of both CTAs)
  if (warp >= 2) {
    const float tot = ss_s[r] + ss_peer[r];
    const float nrm = sqrtf(tot);
    const float den = fmaxf(nrm, kNormEps);
    if (e == 0 && rank == 0 && i < n) {
      inv_den[i] = 1.0f / den;
      nrm_out[i] = nrm;
    }
#pragma unroll 1
    for (int ch = e; ch < 2 * NH; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + ch * 32, raw);
      tmem_ld_wait();
      if (i >= n) continue;
      const int64_t col0 = (int64_t)col_base + ch * 32;
      float* dst = emb + i * d + col0;
      if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
        for (int x = 0; x < 32; x += 4)
          *reinterpret_cast<float4*>(dst + x) = make_float4(__uint_as_float(raw[x]), __uint_as_float(raw[x + 1]),
                                                            __uint_as_float(raw[x + 2]), __uint_as_float(raw[x + 3]));
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x)
          if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
      }
      uint32_t pk[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float lo = __uint_as_float(raw[2 * x]) / den, hi = __uint_as_float(raw[2 * x + 1]) / den;
        pk[x] = f16 ? pack_16x2<true>(lo, hi) : pack_16x2<false>(lo, hi);
      }
      uint4* ud = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(u16) + i * ldu + col0);
#pragma unroll
      for (int x = 0; x < 4; ++x) ud[x] = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int NH>
static int launch_proj2(const CUtensorMap& tx, const CUtensorMap& tw, int64_t n, int64_t d, int kchunks, int f16,
                        void* u16, int64_t ldu, float* emb, float* inv_den, float* nrm, cudaStream_t st) {
  auto kern = proj_norm_tc2<NH>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Proj2Cfg<NH>::kSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * ceil_div(n, kTileRows)));
  cfg.blockDim = dim3(kPjThreads);
  cfg.dynamicSmemBytes = Proj2Cfg<NH>::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PLK_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tw, n, d, kchunks, f16, u16, ldu, emb, inv_den, nrm));
  PLK_LAUNCHED(1);
  return PLK_OK;
}

template <int ND>
static int launch_proj(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tw2, int64_t n, int64_t d,
                       int kchunks, int f16,
                       void* u16, int64_t ldu, float* emb, float* inv_den, float* nrm, cudaStream_t st) {
  auto kern = proj_norm_tc<ND>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ProjCfg<ND>::kSmem));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(n, kTileRows));
  int rc = launch_kernel(kern, grid, dim3(kPjThreads), ProjCfg<ND>::kSmem, st, 1, tx, tw, tw2, n, d, kchunks, f16, u16,
                         ldu, emb, inv_den, nrm);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}

}  // namespace plk

using namespace plk;

extern "C" int plk_project_normalise(const void* feat16, int64_t ldf, const void* w16, int64_t ldw, int op_dtype,
                                     int64_t n, int64_t f, int64_t d, void* u16, int64_t ldu, float* emb,
                                     float* inv_den, float* nrm, void* stream) {
  PLK_REQUIRE(feat16 && w16 && u16 && emb && inv_den && nrm, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype == PLK_BF16 || op_dtype == PLK_F16, PLK_ERR_INVALID,
              "the fused projection runs on 16-bit operands (got op_dtype %d); project in fp32 on the host side", op_dtype);
  PLK_REQUIRE(n > 0 && f > 0 && d > 0, PLK_ERR_INVALID, "bad shape n=%lld f=%lld d=%lld", (long long)n, (long long)f,
              (long long)d);
  PLK_REQUIRE(d <= 512, PLK_ERR_UNSUPPORTED, "the fused projection supports d <= 512 (got %lld)", (long long)d);
  const int64_t fpad = ceil_div(f, kChunkK) * kChunkK, dpad = ceil_div(d, kChunkK) * kChunkK;
  PLK_REQUIRE(ldf >= fpad && ldw >= fpad, PLK_ERR_INVALID,
              "features and weight must be zero-padded to ld >= ceil(f/64)*64 = %lld (ldf=%lld ldw=%lld)", (long long)fpad,
              (long long)ldf, (long long)ldw);
  PLK_REQUIRE(ldu == dpad, PLK_ERR_INVALID, "operand output must have ld = ceil(d/64)*64 = %lld (got %lld)",
              (long long)dpad, (long long)ldu);
  PLK_REQUIRE((((uintptr_t)u16 | (uintptr_t)emb) & 15) == 0, PLK_ERR_INVALID, "outputs must be 16-byte aligned");
  PLK_REQUIRE(plk_device_supports_tc(), PLK_ERR_ARCH, "the tensor-core path needs an sm_100 device");
  const int nd = (int)(dpad / kChunkK);
  const int n1 = dpad > 256 ? 256 : (int)dpad;
  CUtensorMap tx, tw, tw2;
  int rc;
  if ((rc = make_tmap_bf16(&tx, feat16, n, fpad, ldf, kTileRows))) return rc;
  // PLK_PROJ_SPLIT=0: one CTA per row block for every width (the first version of this kernel)
  static const bool split = getenv("PLK_PROJ_SPLIT") == nullptr || getenv("PLK_PROJ_SPLIT")[0] != '0';
  if (split && (nd == 2 || nd == 4 || nd == 8)) {   // output width halved over a cluster of two CTAs
    if ((rc = make_tmap_bf16(&tw, w16, d, fpad, ldw, (int)(dpad / 2)))) return rc;
    cudaStream_t st2 = (cudaStream_t)stream;
    const int kch = (int)(fpad / kChunkK), h16 = op_dtype == PLK_F16;
    switch (nd) {
      case 2: return launch_proj2<1>(tx, tw, n, d, kch, h16, u16, ldu, emb, inv_den, nrm, st2);
      case 4: return launch_proj2<2>(tx, tw, n, d, kch, h16, u16, ldu, emb, inv_den, nrm, st2);
      default: return launch_proj2<4>(tx, tw, n, d, kch, h16, u16, ldu, emb, inv_den, nrm, st2);
    }
  }
  if ((rc = make_tmap_bf16(&tw, w16, d, fpad, ldw, n1))) return rc;
  // widths above 256: the weight rows 256.. arrive through a second box of dpad - 256 rows
  if ((rc = make_tmap_bf16(&tw2, w16, d, fpad, ldw, dpad > 256 ? (int)(dpad - 256) : n1))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int kchunks = (int)(fpad / kChunkK), f16 = op_dtype == PLK_F16;
  switch (nd) {
#define PLK_CASE(ND) \
  case ND: return launch_proj<ND>(tx, tw, tw2, n, d, kchunks, f16, u16, ldu, emb, inv_den, nrm, st);
    PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8)
#undef PLK_CASE
  }
  set_error("unsupported output width %lld", (long long)d);
  return PLK_ERR_UNSUPPORTED;
}
