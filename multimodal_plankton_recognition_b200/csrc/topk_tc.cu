// tcgen05 (sm_100a) path of the k-nearest candidate search.
// The 128 query rows of a CTA live in TENSOR MEMORY (packed 16-bit pairs, written once by the
// epilogue warps with tcgen05.st) and are the A operand of a TS-mode tcgen05.mma: an SS-mode
// 128x128x16 MMA reads 8 KiB of operands from shared memory per 64 cycles, i.e. all of the
// shared-memory bandwidth, and the TMA refill of the ring on top of that capped the mainloop at
// ~65 % of tensor peak; with A in TMEM only the streamed gallery chunks touch shared memory.
// Gallery streamed as [128 x 64] TMA chunks through a deep ring (cluster-multicast), double-
// buffered 128x128 fp32 score tiles in TMEM; the epilogue keeps, per
// thread (= per query row x 64 of the tile's 128 columns), a sorted list of the KC smallest keys
// |g|^2 - 2 q.g.  The common case is one FFMA + one FMNMX per score and a single compare of the
// 32-column minimum against the current KC-th best; the insertion code exists once (not inlined,
// lists in thread-local memory) so the hot loop stays small enough for the instruction cache.
#include <math_constants.h>
#include "tc_common.cuh"

namespace plk {
using namespace tc;

constexpr int kTkThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per sub-partition)
constexpr int kTkEpi = 256;
constexpr int kTkAux = 4096;

template <int KD>
struct TopkCfg {
  static constexpr int kCPS = (KD % 2 == 0) ? 2 : 1;             // 64-wide K chunks per ring stage
  static constexpr int kStageBytes = kCPS * kChunkBytes;         // one barrier round trip per 8 (or 4) MMAs
  static constexpr int kStagesMax = (kMaxSmem - 1024 - kTkAux) / kStageBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kTkAux;
  static constexpr int kACols = KD * 32;   // packed query rows: 64 elements = 32 TMEM columns per chunk
  static_assert(kStages >= 2, "not enough shared memory for the ring");
};

// Insert (key, id) into the ascending register list (fully unrolled, static indices only).
// The new entry starts in the last slot and bubbles forward with strict '<', so among equal
// keys the one inserted first (lower gallery index) stays in front.
template <int KC>
__device__ __forceinline__ void list_insert(float (&bk)[KC], int32_t (&bi)[KC], float key, int32_t id) {
  bk[KC - 1] = key;
  bi[KC - 1] = id;
#pragma unroll
  for (int p = KC - 1; p > 0; --p) {
    const bool sw = bk[p] < bk[p - 1];
    const float k0 = sw ? bk[p] : bk[p - 1], k1 = sw ? bk[p - 1] : bk[p];
    const int32_t i0 = sw ? bi[p] : bi[p - 1], i1 = sw ? bi[p - 1] : bi[p];
    bk[p - 1] = k0; bk[p] = k1;
    bi[p - 1] = i0; bi[p] = i1;
  }
}

template <int KD, int KC, int CS>
__global__ void __launch_bounds__(kTkThreads, 1) topk_tc_kernel(
    const uint16_t* __restrict__ q_op, int64_t ldq, const __grid_constant__ CUtensorMap tmap_g,
    const __grid_constant__ CUtensorMap tmap_gp, const float* __restrict__ g_sqn, int64_t nq, int64_t ng, int64_t goff, int tiles_per_chunk,
    int nchunks, int kc_out, int32_t* __restrict__ out_idx, float* __restrict__ out_key, int f16, int raster) {
  using Cfg = TopkCfg<KD>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_ring = smem;
  constexpr int CPS = Cfg::kCPS;
  uint8_t* aux = sm_ring + NST * Cfg::kStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* bar_empty = bar_full + NST;
  uint64_t* bar_a = bar_empty + NST;
  uint64_t* bar_sfull = bar_a + 1;
  uint64_t* bar_sempty = bar_sfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_sempty + 2);
  float* gsq_s = reinterpret_cast<float*>(aux + 512);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Rasterisation: clusters become resident in launch order (x fastest).  Launch-linear cluster L sweeps gallery
  // chunk L / pairs for query-block pair L % pairs, so the ~74 resident clusters read the SAME chunk in step
  // and a gallery tile comes from DRAM once per wave, not once per query-block pair (3.9x the operand bytes
  // at 8192 x 262144 with chunk-fastest order).
  const int pairs = (int)(gridDim.y / CS);
  const int64_t lin = (int64_t)blockIdx.x + (int64_t)gridDim.x * (blockIdx.y / CS);
  const int64_t q0 = raster ? ((lin % pairs) * CS + blockIdx.y % CS) * kTileRows : (int64_t)blockIdx.y * kTileRows;
  const int chunk = raster ? (int)(lin / pairs) : (int)blockIdx.x;   // PLK_TOPK_RASTER=0: the old order (A/B)
  const int total_tiles = (int)((ng + kTileRows - 1) / kTileRows);
  const int t_begin = chunk * tiles_per_chunk;
  int t_end = t_begin + tiles_per_chunk;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  if (T == 0) {  // empty chunk: emit empty lists
    for (int e = threadIdx.x; e < kTileRows * 2 * kc_out; e += kTkThreads) {
      const int64_t qi = q0 + e / (2 * kc_out);
      if (qi < nq) {
        const int64_t o = (qi * nchunks + chunk) * 2 * kc_out + e % (2 * kc_out);
        out_idx[o] = -1;
        out_key[o] = CUDART_INF_F;
      }
    }
    return;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CS); }
    mbar_init(bar_a, kTkEpi);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_sfull + b, 1); mbar_init(bar_sempty + b, kTkEpi); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();
  tc_fence_after();
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0;
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int t = 0; t < T; ++t) {
        const int j0 = (t_begin + t) * kTileRows;
        for (int c = 0; c < KD; c += CPS) {
          mbar_wait(bar_empty + st, ph ^ 1);
          mbar_expect_tx(bar_full + st, Cfg::kStageBytes);
#pragma unroll
          for (int cs = 0; cs < CPS; ++cs)
            chunk_load<CS>(sm_ring + st * Cfg::kStageBytes + cs * kChunkBytes, &tmap_g, &tmap_gp, bar_full + st,
                           (c + cs) * kChunkK, j0, cta_rank);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    const uint32_t idesc = umma_idesc_16(128, 128, 0, 0, f16);
    mbar_wait(bar_a, 0);     // the epilogue warps have parked the query rows in TMEM
    tc_fence_after();
    const uint32_t a_tmem0 = tmem_base + 256;
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);
    int st = 0; uint32_t ph = 0;
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      mbar_wait(bar_sempty + buf, ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll 1
      for (int c = 0; c < KD; c += CPS) {
        mbar_wait(bar_full + st, ph);
        tc_fence_after();
        const uint32_t b_lo = b_lo0 + st * (Cfg::kStageBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int cs = 0; cs < CPS; ++cs)
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)   // A: 8 packed columns per K step; B: 32 bytes (>>4 = 2)
              umma_bf16_ts(d_tmem, a_tmem0 + (c + cs) * 32 + k * 8, b_lo + cs * (kChunkBytes >> 4) + 2 * k, idesc,
                           (c | cs | k) != 0);
          ring_release<CS>(bar_empty + st);
          if (c + CPS >= KD) umma_commit(bar_sfull + buf);
        }
        __syncwarp();
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int64_t qi = q0 + r;
    float bk[KC];
    int32_t bi[KC];
#pragma unroll
    for (int e = 0; e < KC; ++e) { bk[e] = CUDART_INF_F; bi[e] = -1; }
    float thresh = CUDART_INF_F;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    {  // park this thread's query row (16-bit operand, padded to KD*64) in TMEM columns 256.. as packed pairs;
       // the two warps of a lane quadrant take alternate 64-element chunks
      const uint4* qrow = reinterpret_cast<const uint4*>(q_op + qi * ldq);
      for (int c = half; c < KD; c += 2) {
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);
          if (qi < nq) w = qrow[c * 8 + v4];
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + lane_addr + 256 + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a);
    }
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = (int64_t)(t_begin + t) * kTileRows;
      if (half == 0) gsq_s[buf * 128 + r] = (j0 + r < ng) ? g_sqn[j0 + r] : CUDART_INF_F;
      named_barrier_sync(1, kTkEpi);
      mbar_wait(bar_sfull + buf, (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int cc = half * 2 + c2;
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + buf * 128 + cc * 32, raw);
        tmem_ld_wait();
        const float4* g4 = reinterpret_cast<const float4*>(gsq_s + buf * 128 + cc * 32);
        float key[32];
        float m0 = CUDART_INF_F, m1 = CUDART_INF_F;
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 gq = g4[e4];
          key[e4 * 4 + 0] = fmaf(-2.0f, __uint_as_float(raw[e4 * 4 + 0]), gq.x);
          key[e4 * 4 + 1] = fmaf(-2.0f, __uint_as_float(raw[e4 * 4 + 1]), gq.y);
          key[e4 * 4 + 2] = fmaf(-2.0f, __uint_as_float(raw[e4 * 4 + 2]), gq.z);
          key[e4 * 4 + 3] = fmaf(-2.0f, __uint_as_float(raw[e4 * 4 + 3]), gq.w);
          m0 = fminf(m0, fminf(key[e4 * 4 + 0], key[e4 * 4 + 1]));
          m1 = fminf(m1, fminf(key[e4 * 4 + 2], key[e4 * 4 + 3]));
        }
        // Slow path, taken by the whole warp while any lane still holds a key below its current
        // KC-th best: extract-min from the 32 register keys, insert, clear, repeat.  One copy of
        // the insertion code (the column-chunk loop is not unrolled), no local memory.
        float m = fminf(m0, m1);
        while (__any_sync(0xffffffffu, m < thresh)) {
          if (m < thresh) {
            int me = 0;
#pragma unroll
            for (int e = 31; e >= 0; --e) me = (key[e] == m) ? e : me;   // lowest column on ties
            list_insert<KC>(bk, bi, m, (int32_t)(goff + j0 + cc * 32) + me);
            thresh = bk[KC - 1];
#pragma unroll
            for (int e = 0; e < 32; ++e) key[e] = (e == me) ? CUDART_INF_F : key[e];
          }
          float a0 = CUDART_INF_F, a1 = CUDART_INF_F;
#pragma unroll
          for (int e = 0; e < 32; e += 2) { a0 = fminf(a0, key[e]); a1 = fminf(a1, key[e + 1]); }
          m = fminf(a0, a1);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_sempty + buf);
    }
    if (qi < nq) {
      const int64_t o = (qi * (nchunks * 2) + chunk * 2 + half) * kc_out;
#pragma unroll
      for (int e = 0; e < KC; ++e)
        if (e < kc_out) {
          out_idx[o + e] = bi[e];
          out_key[o + e] = bk[e];
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int topk_bf16_chunks(int64_t nq, int64_t ng) {
  // (query-block pairs x gallery chunks) clusters run in waves of 74 (148 SMs / cluster of 2).
  // Pick the chunk count (chunks of >= 64 tiles: every chunk restarts the per-thread lists, and the
  // warm-up phase of a list is the slow path) that wastes the least of the last wave.
  const int64_t clusters = ceil_div(ceil_div(nq, kTileRows), 2);
  const int64_t tiles = ceil_div(ng, kTileRows);
  int64_t maxc = tiles / 64;
  if (maxc < 1) maxc = 1;
  if (maxc > 16) maxc = 16;
  int best = 1;
  double best_eff = 0.0;
  for (int64_t c = 1; c <= maxc; ++c) {
    const int64_t work = clusters * c;
    const double eff = (double)work / (double)(ceil_div(work, 74) * 74);
    if (eff > best_eff + 0.02) { best_eff = eff; best = (int)c; }
  }
  return best;
}

size_t topk_ws_tc16(int64_t nq, int64_t ng, int64_t d, int kc) {
  const int c = topk_bf16_chunks(nq, ng);
  return (size_t)nq * c * 2 * kc * (sizeof(int32_t) + sizeof(float));  // two column-half lists per chunk
}

static int topk_raster() {
  static const int v = [] { const char* e = getenv("PLK_TOPK_RASTER"); return (e && e[0] == '0') ? 0 : 1; }();
  return v;
}

template <int KD, int KC, int CS>
static int launch_topk(const void* q, int64_t ldq, const CUtensorMap& tg, const CUtensorMap& tgp, dim3 grid,
                       const float* g_sqn, int64_t nq, int64_t ng, int64_t goff, int tpc, int nchunks,
                       int kc, int32_t* o_idx, float* o_key, int f16, cudaStream_t st) {
  auto kern = topk_tc_kernel<KD, KC, CS>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TopkCfg<KD>::kSmem));
    configured = true;
  }
  int rc = launch_kernel(kern, grid, dim3(kTkThreads), TopkCfg<KD>::kSmem, st, CS, (const uint16_t*)q, ldq, tg, tgp, g_sqn, nq, ng,
                         goff, tpc, nchunks, kc, o_idx, o_key, f16, topk_raster());
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}

int topk_candidates_tc16(const void* q, const void* g, int f16, int64_t ld,
                         const float* g_sqn, int64_t nq, int64_t ng, int64_t d, int kc, int64_t goff,
                         int32_t* cand_idx, float* cand_key, void* ws, size_t ws_bytes,
                         cudaStream_t st) {
  PLK_REQUIRE(ld % kChunkK == 0 && ld >= d && ld - d < kChunkK, PLK_ERR_INVALID,
              "16-bit operands must be zero-padded to ld = ceil(d/64)*64 (d=%lld ld=%lld)", (long long)d, (long long)ld);
  PLK_REQUIRE(ld <= 512, PLK_ERR_UNSUPPORTED, "tensor-core path supports d <= 512 (got %lld)", (long long)d);
  PLK_REQUIRE(kc <= 32, PLK_ERR_UNSUPPORTED, "tensor-core path keeps at most 32 candidates per query (got %d)", kc);
  PLK_REQUIRE(plk_device_supports_tc(), PLK_ERR_ARCH, "the tensor-core path needs an sm_100 device");
  int rc;
  PLK_REQUIRE(((uintptr_t)q & 15) == 0, PLK_ERR_INVALID, "query operand must be 16-byte aligned");
  CUtensorMap tg, tgp;
  if ((rc = make_tmap_bf16(&tg, g, ng, ld, ld, kTileRows))) return rc;
  if ((rc = make_tmap_bf16(&tgp, g, ng, ld, ld, kTileRows / 2))) return rc;
  const int nchunks = topk_bf16_chunks(nq, ng);
  const int tpc = (int)ceil_div(ceil_div(ng, kTileRows), nchunks);
  int32_t* o_idx = (int32_t*)ws;
  float* o_key = (float*)((char*)ws + (size_t)nq * nchunks * 2 * kc * sizeof(int32_t));
  int64_t qblocks = ceil_div(nq, kTileRows);
  const int cs = qblocks >= 2 ? 2 : 1;   // pairs of query blocks share every gallery chunk (multicast)
  qblocks = ceil_div(qblocks, cs) * cs;
  dim3 grid((unsigned)nchunks, (unsigned)qblocks, 1);
  const int kd = (int)(ld / kChunkK);
  rc = PLK_ERR_UNSUPPORTED;
#define PLK_TK(KD, KC, CS) launch_topk<KD, KC, CS>(q, ld, tg, tgp, grid, g_sqn, nq, ng, goff, tpc, nchunks, kc, o_idx, o_key, f16, st)
#define PLK_CASE(KD)                                                              \
  case KD:                                                                        \
    rc = kc <= 16 ? (cs == 2 ? PLK_TK(KD, 16, 2) : PLK_TK(KD, 16, 1))              \
                  : (cs == 2 ? PLK_TK(KD, 32, 2) : PLK_TK(KD, 32, 1));             \
    break;
  switch (kd) { PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8) }
#undef PLK_TK
#undef PLK_CASE
  if (rc) return rc;
  return select_candidates(o_idx, o_key, nq, nchunks * 2 * kc, kc, cand_idx, cand_key, st);
}

}  // namespace plk
