// tcgen05 (sm_100a) path of the fused similarity + InfoNCE forward and recompute backward.
//
// Both kernels are warp-specialised:
//   warp 0      TMA producer   (one elected lane; cp.async.bulk.tensor into a 16 KiB-stage ring)
//   warp 1      MMA issuer     (one elected lane; tcgen05.mma, accumulators in TMEM) + TMEM alloc
//   warps 2..17 epilogue       (tcgen05.ld 32x32b: thread = one logit row, 32 columns per load;
//                               warps w, w+4, w+8, w+12 share a TMEM lane quadrant and split the columns)
// A CTA owns 128 rows of the "a" operand (resident in shared memory, K-major SWIZZLE_128B) and
// streams [128 x 64] chunks of the "b" operand.  The B x B logits only ever exist as 128x128 fp32
// tiles in TMEM (double buffered so the epilogue of tile t overlaps the MMAs of tile t+1).
//
// Forward epilogue  : E = exp2(acc*s*log2e - (s - 64)*log2e); row sums in registers, diagonal pick; column sums by
//                     storing E back into its tensor-memory chunk and re-reading it as 16x256b fragments (four rows
//                     per thread), then 7 shuffles per 32 columns (PLK_FWD_TT=0: the 31-shuffle transpose-reduce).
// Backward epilogue : G = E*(1/rs_i + 1/cs_j) -> bf16 -> shared memory (swizzled K-major A operand),
//                     then a second tcgen05.mma  acc[128 x 64*c] += G . b_chunk  with the streamed
//                     chunk reused as an MN-major B operand; acc (<= 256 fp32 columns) stays in TMEM
//                     for the whole column sweep.
#include "tc_common.cuh"

namespace plk {
using namespace tc;

// Development aid (-DPLK_TRACE, see tools/trace_tc.py): per-CTA clock stamps of the pipeline events.
#ifdef PLK_TRACE
constexpr int kTraceSlots = 128;
static __device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int slot) {
  if (g_trace == nullptr) return;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  g_trace[(size_t)cta * kTraceSlots + slot] = clock64();
  if (slot == 0 || slot == 6) {   // wall-clock (ns) at CTA entry / exit: slots 40 / 41
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_trace[(size_t)cta * kTraceSlots + (slot == 0 ? 40 : 41)] = (long long)gt;
  }
}
__device__ __forceinline__ void trace_value(int slot, long long v) {
  if (g_trace == nullptr) return;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  g_trace[(size_t)cta * kTraceSlots + slot] = v;
}
#define TR(slot) trace_stamp(slot)
#define TRV(slot, v) trace_value(slot, v)   // used by the -DPLK_TRACE_PROBE blocks (serialised MMA durations)
#else
#define TR(slot) ((void)0)
#define TRV(slot, v) ((void)0)
#endif

constexpr int kNumThreads = 576;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int kEpiThreads = 512;   // four warps per SM sub-partition (latency hiding): each takes one
                                   // 32-column chunk of a tile's 128 columns
constexpr int kAuxBytes = 8192;  // barriers + tmem pointer (first 512 B), per-tile column scratch

__device__ __forceinline__ void row_block_cols(int64_t i0, int64_t n_rows, int64_t row_offset,
                                               int64_t bs, int64_t n_cols, int64_t& jlo,
                                               int64_t& jhi) {
  int64_t last = i0 + kTileRows - 1;
  if (last >= n_rows) last = n_rows - 1;
  int64_t lo, hi, lo2, hi2;
  bucket_range(row_offset + i0, bs, n_cols, lo, hi);
  bucket_range(row_offset + last, bs, n_cols, lo2, hi2);
  jlo = lo;
  jhi = hi2;
}

#ifndef PLK_FWD_TT
#define PLK_FWD_TT 1   // forward column sums through a tensor-memory transposition (see infonce_fwd_tc); 0: 31-shuffle transpose-reduce
#endif

// column sums over the 32 lanes of a warp for 32 columns held one-row-per-lane:
// on return v[0] of lane l is the sum of column l.
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? v[i + h] : v[i];
      const float send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p)));
  return v;
}

// Backward epilogue for one 32-column chunk of a logits tile held one-row-per-thread:
//   G = ex2(a*c1 + c0) * (1/rs_i + 1/cs_j)  -> packed 16-bit pairs (A operand of G.V).
// The epilogue is instruction-issue bound (16 k elements per 128x128 tile against ~2 k cycles of
// tensor work), so the common case -- chunk fully inside the row's bucket, no diagonal column, no
// sum G*S wanted -- is a separate instantiation with 5.5 instructions per element; masks use
// 32-bit chunk-relative bounds.
template <bool F16, bool EDGE, bool GS>
__device__ __forceinline__ void grad_chunk(const uint32_t (&raw)[32], uint32_t (&packed)[16],
                                           const float* rc_smem, float rrs, float c1, float c0,
                                           int lo_rel, int hi_rel, int dei, float& gs_local) {
  // element pairs on the packed fp32 pipe: args = a*c1 + c0 (FFMA2), weights = 1/rs + 1/cs (FADD2),
  // G = E * weights (FMUL2), sum G*a (FFMA2); the two exponentials of a pair stay MUFU.EX2
  const uint64_t c1p = f2_pack(c1, c1), c0p = f2_pack(c0, c0), rrsp = f2_pack(rrs, rrs);
  uint64_t gsp = f2_pack(0.f, 0.f);
#pragma unroll
  for (int e4 = 0; e4 < 8; ++e4) {
    const float4 rc = lds_f4(rc_smem + e4 * 4);
    const uint64_t rcp[2] = {f2_pack(rc.x, rc.y), f2_pack(rc.z, rc.w)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = e4 * 4 + h * 2;
      const uint64_t a2 = f2_pack_u(raw[e], raw[e + 1]);
      const uint64_t x2 = f2_fma(a2, c1p, c0p);
      uint64_t e2;
      if (exp_pair_on_fma(e4 * 2 + h)) {   // compile-time choice (the loops are fully unrolled)
        e2 = ex2_poly2(x2);
      } else {
        float x0, x1;
        f2_unpack(x2, x0, x1);
        e2 = f2_pack(ex2_approx(x0), ex2_approx(x1));
      }
      uint64_t g2 = f2_mul(e2, f2_add(rcp[h], rrsp));
      if constexpr (EDGE) {
        float g0, g1;
        f2_unpack(g2, g0, g1);
        g0 = (e >= lo_rel && e < hi_rel) ? g0 : 0.f;
        g1 = (e + 1 >= lo_rel && e + 1 < hi_rel) ? g1 : 0.f;
        g2 = f2_pack(g0, g1);
      }
      if constexpr (GS) gsp = f2_fma(g2, a2, gsp);
      float g0, g1;
      f2_unpack(g2, g0, g1);
      // the j == i term is added in fp32 by plk_infonce_grad_finish (it dominates a peaked softmax
      // and nearly cancels against the -2*delta term): drop it from the 16-bit operand
      if constexpr (EDGE) {
        g0 = (e == dei) ? 0.f : g0;
        g1 = (e + 1 == dei) ? 0.f : g1;
      }
      packed[e4 * 2 + h] = pack_16x2<F16>(g0, g1);
    }
  }
  if constexpr (GS) {
    float s0, s1;
    f2_unpack(gsp, s0, s1);
    gs_local += s0 + s1;
  }
}

// SigLIP variant (reference src/coordination.py:85-93): G = sigmoid(z) off the diagonal, z = s a + bias,
// given as zl = a*c1 + c0 with c1 = s log2(e), c0 = bias log2(e).  No row / column statistics.
// gs_local += G a (sum G*S after the final * s), gsum_local += G (for d bias).
template <bool F16, bool EDGE, bool GS>
__device__ __forceinline__ void sig_chunk(const uint32_t (&raw)[32], uint32_t (&packed)[16], float c1, float c0,
                                          int lo_rel, int hi_rel, int dei, float& gs_local, float& gsum_local) {
#pragma unroll
  for (int e2 = 0; e2 < 16; ++e2) {
    float gg[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int e = e2 * 2 + x;
      const float a = __uint_as_float(raw[e]);
      float G = sigmoid_l2(fmaf(a, c1, c0));
      if constexpr (EDGE) G = (e >= lo_rel && e < hi_rel && e != dei) ? G : 0.f;
      if constexpr (GS) {
        gs_local = fmaf(G, a, gs_local);
        gsum_local += G;
      }
      gg[x] = G;
    }
    packed[e2] = pack_16x2<F16>(gg[0], gg[1]);
  }
}

// chunk-relative validity window [lo_rel, hi_rel) and diagonal position of a row for the 32 columns
// starting at global column jc0 (all 32-bit)
__device__ __forceinline__ void chunk_window(int64_t lo, int64_t hi, int64_t gi, int64_t jc0, int& lo_rel,
                                             int& hi_rel, int& dei) {
  const int64_t l = lo - jc0, h = hi - jc0, de = gi - jc0;
  lo_rel = l < 0 ? 0 : (l > 32 ? 32 : (int)l);
  hi_rel = h < 0 ? 0 : (h > 32 ? 32 : (int)h);
  dei = (de >= 0 && de < 32) ? (int)de : -1;
}

template <bool F16>
__device__ __forceinline__ void grad_chunk_dispatch(const uint32_t (&raw)[32], uint32_t (&packed)[16],
                                                    const float* rc_smem, float rrs, float c1, float c0,
                                                    int64_t lo, int64_t hi, int64_t gi, int64_t jc0,
                                                    bool want_gs, float& gs_local, bool siglip,
                                                    float& gsum_local) {
  int lo_rel, hi_rel, dei;
  chunk_window(lo, hi, gi, jc0, lo_rel, hi_rel, dei);
  const bool plain = __all_sync(0xffffffffu, lo_rel == 0 && hi_rel == 32 && dei < 0);
  if (siglip) {   // warp-uniform (a launch parameter)
    if (plain) {
      if (want_gs) sig_chunk<F16, false, true>(raw, packed, c1, c0, 0, 32, -1, gs_local, gsum_local);
      else sig_chunk<F16, false, false>(raw, packed, c1, c0, 0, 32, -1, gs_local, gsum_local);
    } else {
      if (want_gs) sig_chunk<F16, true, true>(raw, packed, c1, c0, lo_rel, hi_rel, dei, gs_local, gsum_local);
      else sig_chunk<F16, true, false>(raw, packed, c1, c0, lo_rel, hi_rel, dei, gs_local, gsum_local);
    }
    return;
  }
  if (plain) {
    if (want_gs) grad_chunk<F16, false, true>(raw, packed, rc_smem, rrs, c1, c0, 0, 32, -1, gs_local);
    else grad_chunk<F16, false, false>(raw, packed, rc_smem, rrs, c1, c0, 0, 32, -1, gs_local);
  } else {
    if (want_gs) grad_chunk<F16, true, true>(raw, packed, rc_smem, rrs, c1, c0, lo_rel, hi_rel, dei, gs_local);
    else grad_chunk<F16, true, false>(raw, packed, rc_smem, rrs, c1, c0, lo_rel, hi_rel, dei, gs_local);
  }
}

// =============================================================================================
// forward
// =============================================================================================
template <int KD>
struct FwdCfg {
  static constexpr int kCPS = (KD % 2 == 0) ? 2 : 1;             // 64-wide K chunks per ring stage
  static constexpr int kStageBytes = kCPS * kChunkBytes;
  static constexpr int kStagesMax = (kMaxSmem - 1024 - kAuxBytes) / kStageBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kAuxBytes;
  static_assert(kStages >= 2, "not enough shared memory for the ring");
};

template <int KD, int CS>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_fwd_tc(
    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
    const __grid_constant__ CUtensorMap tmap_bp, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t bs, int tiles_per_seg,
    const float* __restrict__ ls, float* __restrict__ row_sumexp, float* __restrict__ col_sumexp,
    float* __restrict__ diag, int f16, const float* __restrict__ sig_bias, double* __restrict__ sig_sums) {
  // sig_bias != nullptr: SigLIP epilogue (reference src/coordination.py:85-93) on the same mainloop --
  // no row / column statistics; sig_sums[0..2] += (loss terms, sum_i G_ii S_ii, sum_i G_ii)
  using Cfg = FwdCfg<KD>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int CPS = Cfg::kCPS;
  uint8_t* sm_ring = smem;
  uint8_t* aux = sm_ring + NST * Cfg::kStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);  // [NST]
  uint64_t* bar_empty = bar_full + NST;                   // [NST]
  uint64_t* bar_a = bar_empty + NST;                      // [1]
  uint64_t* bar_sfull = bar_a + 1;                        // [2]
  uint64_t* bar_sempty = bar_sfull + 2;                   // [2]
  uint64_t* bar_afull = bar_sempty + 2;                   // [1] owned rows landed in their staging slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_afull + 1);
  // The owned rows arrive by TMA in the LAST KD chunk slots of the ring, are moved to TMEM by the
  // epilogue warps, and only then does the producer let streamed chunks into those slots.
  constexpr int kASlot0 = NST * CPS - KD;
  constexpr int kAStage0 = kASlot0 / CPS;
  static_assert(kASlot0 >= 0, "ring too small to stage the owned rows");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  int64_t jlo = 0, jhi = n_cols;  // clusters are only launched for the single-bucket case
  if constexpr (CS == 1) row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  if (t_begin >= t_end) return;  // uniform across the CTA (and across the cluster)
  const int T = t_end - t_begin;
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) TR(0);
  pdl_trigger();   // the loss kernel may become resident; it waits for this grid's completion

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CS); }
    // owned rows parked in TMEM -- in EVERY CTA of the cluster: a peer's multicast writes into this CTA's
    // ring slots as well, including the tail slots the owned rows are staged in
    mbar_init(bar_a, CS * (kEpiThreads / 32));
    mbar_init(bar_afull, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_sfull + b, 1); mbar_init(bar_sempty + b, kEpiThreads / 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();   // peers' barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up (barriers, tensor memory, descriptor prefetch) may have run under the tail of the
  // normalisation kernel; the operands and the zero-filled accumulators are its products
  pdl_wait();
  if (threadIdx.x == 0) TR(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_afull, KD * kChunkBytes);
      for (int c = 0; c < KD; ++c)
        tma_load_2d(sm_ring + (kASlot0 + c) * kChunkBytes, &tmap_a, bar_afull, c * kChunkK, (int)i0);
      bool a_parked = false;
      int st = 0; uint32_t ph = 0;
      for (int t = 0; t < T; ++t) {
        const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
        for (int c = 0; c < KD; c += CPS) {
          if (!a_parked && st >= kAStage0) {   // first use of a stage that overlaps the staging slots
            mbar_wait(bar_a, 0);
            a_parked = true;
          }
          mbar_wait(bar_empty + st, ph ^ 1);
          mbar_expect_tx(bar_full + st, Cfg::kStageBytes);
#pragma unroll
          for (int cs = 0; cs < CPS; ++cs)
            chunk_load<CS>(sm_ring + st * Cfg::kStageBytes + cs * kChunkBytes, &tmap_b, &tmap_bp, bar_full + st,
                           (c + cs) * kChunkK, j0, cta_rank);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
        if (t < 16) TR(48 + t);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    const uint32_t idesc = umma_idesc_16(128, 128, 0, 0, f16);
    mbar_wait(bar_a, 0);     // the epilogue warps have parked the owned rows in TMEM (columns 256..)
    tc_fence_after();
    if (lane == 0) TR(3);
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
            for (int k = 0; k < kChunkK / kUmmaK; ++k)   // A: 8 packed TMEM columns per K step; B: 32 bytes
              umma_bf16_ts(d_tmem, a_tmem0 + (c + cs) * 32 + k * 8, b_lo + cs * (kChunkBytes >> 4) + 2 * k, idesc,
                           (c | cs | k) != 0);
          ring_release<CS>(bar_empty + st);
          if (c + CPS >= KD) umma_commit(bar_sfull + buf);
        }
        __syncwarp();
        if (lane == 0 && t == 0 && c == 0) TR(4);
        if (++st == NST) { st = 0; ph ^= 1; }
      }
      if (lane == 0 && t < 16) TR(64 + t);
    }
  } else {
    // ---------------- epilogue: 16 warps, thread = one logit row x 32 columns ----------------
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int cc = (warp - 2) >> 2;          // which 32-column chunk of the tile
    const int r = q * 32 + lane;             // row inside the 128-row block
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    if (i < n_rows) bucket_range(gi, bs, n_cols, lo, hi);
    const float s = expf(*ls);
    const bool siglip = sig_bias != nullptr;
    const float c1 = s * kLog2e, c0 = siglip ? *sig_bias * kLog2e : (kShiftK - s) * kLog2e;
    float rsum = 0.f;
    float sg_loss = 0.f, sg_gs = 0.f, sg_g = 0.f;
    {  // park this thread's owned row (16-bit operand, padded to KD*64) in TMEM columns 256.. as packed
       // pairs -- the A operand of the TS-mode MMA; the four warps of a lane quadrant take alternate chunks.
       // Source: the swizzled staging slots (rows past n_rows were zero-filled by the TMA unit).
      mbar_wait(bar_afull, 0);
      for (int c = cc; c < KD; c += 4) {
        const uint8_t* rowp = sm_ring + (kASlot0 + c) * kChunkBytes + r * 128;
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const uint4 w = lds_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4));
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + 256 + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_a);
        if constexpr (CS > 1) mbar_arrive_remote(bar_a, cta_rank ^ 1);
      }
      if (threadIdx.x == 64) TR(2);
    }
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows + cc * 32;   // first column of this chunk
      mbar_wait(bar_sfull + buf, (t >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && t < 16) TR(80 + t);
      uint32_t raw[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + cc * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      // the accumulator buffer goes back to the MMA warp right away (one arrive per warp: 512 arrives on one word cost
      // MIO slots) -- except for InfoNCE under PLK_FWD_TT, where the chunk is reused for the column-sum transposition
      if (lane == 0 && (!PLK_FWD_TT || siglip)) mbar_arrive(bar_sempty + buf);
      if (threadIdx.x == 64 && t < 8) TR(8 + t);              // trace: first epilogue warp has its logits
      if (threadIdx.x == 17 * 32 && t < 8) TR(24 + t);        // trace: last epilogue warp has its logits
      const bool full = (j0 >= lo) && (j0 + 32 <= hi);
      const bool warp_full = __all_sync(0xffffffffu, full);
      const bool has_diag = __any_sync(0xffffffffu, gi >= j0 && gi < j0 + 32 && i < n_rows);
      if (siglip) {   // warp-uniform (a launch parameter)
        float part = 0.f;
        if (warp_full) {
#pragma unroll
          for (int e = 0; e < 32; ++e) part += softplus_l2(fmaf(__uint_as_float(raw[e]), c1, c0));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int64_t j = j0 + e;
            const float sp = softplus_l2(fmaf(__uint_as_float(raw[e]), c1, c0));
            part += (j >= lo && j < hi) ? sp : 0.f;
          }
        }
        if (has_diag) {
          const int64_t de = gi - j0;
          if (de >= 0 && de < 32 && i < n_rows) {
            const int dei = (int)de;
            float dv = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) dv = fmaf(__uint_as_float(raw[e]), (e == dei) ? 1.0f : 0.0f, dv);
            const float S = s * dv;
            diag[i] = S;
            // the diagonal term is softplus(-z), not softplus(z): swap it (same bits were added above)
            const float zl = fmaf(dv, c1, c0);
            part += softplus_l2(-zl) - softplus_l2(zl);
            const float gd = -sigmoid_l2(-zl);
            sg_gs = fmaf(gd, S, sg_gs);
            sg_g += gd;
          }
        }
        sg_loss += part;
        if (threadIdx.x == 64 && t < 16) TR(96 + t);
        continue;
      }
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {   // one pair in four on the FMA pipe (ex2_poly2), the rest MUFU.EX2
        if (exp_pair_on_fma(e >> 1)) {
          f2_unpack(ex2_poly2(f2_fma(f2_pack_u(raw[e], raw[e + 1]), f2_pack(c1, c1), f2_pack(c0, c0))), v[e], v[e + 1]);
        } else {
          v[e] = ex2_approx(fmaf(__uint_as_float(raw[e]), c1, c0));
          v[e + 1] = ex2_approx(fmaf(__uint_as_float(raw[e + 1]), c1, c0));
        }
      }
      if (!warp_full) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int64_t j = j0 + e;
          v[e] = (j >= lo && j < hi) ? v[e] : 0.f;
        }
      }
      if (has_diag) {
        const int64_t de = gi - j0;
        if (de >= 0 && de < 32 && i < n_rows) {
          // mask-FMA pick (keeps raw[] in registers: a select chain gets turned into a
          // dynamically indexed local-memory array by the compiler)
          const int dei = (int)de;
          float dv = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) dv = fmaf(__uint_as_float(raw[e]), (e == dei) ? 1.0f : 0.0f, dv);
          diag[i] = s * dv;
        }
      }
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
      for (int e = 0; e < 32; e += 4) { p0 += v[e]; p1 += v[e + 1]; p2 += v[e + 2]; p3 += v[e + 3]; }
      rsum += (p0 + p1) + (p2 + p3);
#if PLK_FWD_TT
      {
        // Column sums through tensor memory: E goes back into the chunk it came from (32x32b: thread = row) and is
        // read as two 16x256b fragments (thread = 4 rows x 8 columns), so 24 of the 31 cross-lane steps become
        // in-thread adds: 7 shuffles per 32 x 32 block instead of 31 (the MIO queue is shared with the 32 MUFU.EX2).
        const uint32_t chunk_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + cc * 32;
        uint32_t eb[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) eb[e] = __float_as_uint(v[e]);
        tmem_st32(chunk_addr, eb);
        tmem_st_wait();
        uint32_t f0[16], f1[16];
        tmem_ld_16x256b_x4(chunk_addr, f0);                      // lanes q*32 + 0..15
        tmem_ld_16x256b_x4(chunk_addr + (16u << 16), f1);        // lanes q*32 + 16..31
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sempty + buf);
        // r[4n + 2h + b] = (row t/4 + 8h, column 8n + 2(t%4) + b): add the thread's four rows
        float c8[8];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int b = 0; b < 2; ++b)
            c8[n * 2 + b] = (__uint_as_float(f0[4 * n + b]) + __uint_as_float(f0[4 * n + 2 + b])) +
                            (__uint_as_float(f1[4 * n + b]) + __uint_as_float(f1[4 * n + 2 + b]));
        // the 8 lanes t%4 + 4k share these 8 columns: transpose-reduce over lane bits 4, 3, 2
#pragma unroll
        for (int h = 4; h >= 1; h >>= 1) {
          const bool up = (lane & (h * 4)) != 0;
#pragma unroll
          for (int x = 0; x < h; ++x) {
            const float keep = up ? c8[x + h] : c8[x];
            const float send = up ? c8[x] : c8[x + h];
            c8[x] = keep + __shfl_xor_sync(0xffffffffu, send, h * 4);
          }
        }
        // lane bits 4, 3 picked n, bit 2 picked b
        const int col = 8 * (((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)) + 2 * (lane & 3) + ((lane >> 2) & 1);
        const int64_t j = j0 + col;
        if (j < n_cols && c8[0] != 0.f) atomicAdd(col_sumexp + j, c8[0]);
      }
#else
      // column sums over this warp's 32 rows, then one 128-byte reduction per warp
      warp_transpose_reduce(v, lane);
      const int64_t j = j0 + lane;
      if (j < n_cols && v[0] != 0.f) atomicAdd(col_sumexp + j, v[0]);
#endif
      if (threadIdx.x == 64 && t < 16) TR(96 + t);
      if (threadIdx.x == 17 * 32 && t < 8) TR(32 + t);        // trace: last epilogue warp done
    }
    if (siglip) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sg_loss += __shfl_xor_sync(0xffffffffu, sg_loss, o);
        sg_gs += __shfl_xor_sync(0xffffffffu, sg_gs, o);
        sg_g += __shfl_xor_sync(0xffffffffu, sg_g, o);
      }
      if (lane == 0) {
        atomicAdd(sig_sums, (double)sg_loss);
        if (sg_g != 0.f) {
          atomicAdd(sig_sums + 1, (double)sg_gs);
          atomicAdd(sig_sums + 2, (double)sg_g);
        }
      }
    } else if (i < n_rows) {
      atomicAdd(row_sumexp + i, rsum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TR(5);
  if constexpr (CS > 1) cluster_sync_exec();   // no CTA leaves while a peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) TR(6);
}

// =============================================================================================
// backward.  One launch serves one direction (ndir = 1) or both (ndir = 2: blockIdx.z % 2 picks
// the operand set), so that at small batches all SMs are busy with half as many column segments
// and a single prologue / drain.
// =============================================================================================
struct GradDir {
  CUtensorMap ta, tb, tbp;   // owned rows (box 128), streamed rows (box 128), streamed rows (box 64, multicast)
  CUtensorMap tacc;          // acc as fp32 [nseg][n_rows][d], box 128 x 32 (valid when use_tacc)
  int use_tacc;
  const float* rs;           // sum-exp along the owned rows
  const float* cs;           // sum-exp along the streamed rows
  float* acc;                // [nseg][n_rows][d] partial accumulators
  float* gs;                 // nullable: += sum G*S
};
// Gradient tail fused into infonce_grad_tc4 (InfoNCE, single bucket, d % 128 == 0): the LAST column segment of a
// (row block, direction) to finish adds the partial slabs and runs the tail on the 128 rows; the last tail of
// the grid produces d logit_scale.  Counters are zero on entry and left zero.
struct GradTail {
  int enabled;
  int nseg;                 // column segments (= partial slabs) per row block and direction
  int64_t ldx, batch;
  const float* x[2];        // RAW rows of direction k's own modality (the partner is x[1 - k])
  const float* inv_den[2];
  const float* nrm[2];
  float* dx[2];
  const float* diag;        // S_ii, row sums R and column sums C of the owned rows (plk_infonce_grad_finish)
  const float* R;
  const float* C;
  const float* grad_out_emb;
  const float* grad_out;
  float emb_scale;
  const float* diag_sum;
  float* dls_out;
  int* counters;            // [ndir * row_blocks] arrivals per (direction, row block), then [1] finished tails
};
struct GradArgs {
  GradDir dir[2];
  int ndir;
  const float* bias;   // non-null: SigLIP weights (rs / cs unused, gs -> float[2] = (sum G*S, sum G))
  GradTail tail;
};
template <int KD, int DNC>
struct GradCfg {
  static constexpr int kResident = KD * kChunkBytes;
  static constexpr int kGBytes = 2 * kChunkBytes;  // G tile: two [128 x 64] bf16 sub-tiles
  static constexpr int kStagesMax = (kMaxSmem - 1024 - kAuxBytes - kResident - kGBytes) / kChunkBytes;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr int kSmem = 1024 + kResident + kGBytes + kStages * kChunkBytes + kAuxBytes;
  static_assert(kStages >= 2, "not enough shared memory for the ring");
  static_assert(DNC >= 1 && DNC <= 4, "at most 256 accumulator columns per CTA");
};

template <int KD, int DNC, int CS, bool F16>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, int tiles_per_seg,
    const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = GradCfg<KD, DNC>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_a = smem;
  uint8_t* sm_g = smem + Cfg::kResident;
  uint8_t* sm_ring = sm_g + Cfg::kGBytes;
  uint8_t* aux = sm_ring + NST * kChunkBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);  // [NST]
  uint64_t* bar_empty = bar_full + NST;                   // [NST]
  uint64_t* bar_a = bar_empty + NST;                      // [1]
  uint64_t* bar_sfull = bar_a + 1;                        // [2]
  uint64_t* bar_sempty = bar_sfull + 2;                   // [2]
  uint64_t* bar_gfull = bar_sempty + 2;                   // [1]
  uint64_t* bar_gempty = bar_gfull + 1;                   // [1]
  uint64_t* bar_accfull = bar_gempty + 1;                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);     // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  const int h = blockIdx.z / ga.ndir;  // which block of DNC*64 output columns
  int64_t jlo = 0, jhi = n_cols;  // clusters are only launched for the single-bucket case
  if constexpr (CS == 1) row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0;
  float* acc_out = g.acc + (int64_t)blockIdx.x * n_rows * d;
  if (T == 0) {  // this segment has no tiles: its partial is zero (uniform across the CTA)
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DNC * 64; e += kNumThreads) {
      const int64_t rr = i0 + e / (DNC * 64), col = (int64_t)h * DNC * 64 + e % (DNC * 64);
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tb);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CS); }
    mbar_init(bar_a, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_sfull + b, 1); mbar_init(bar_sempty + b, kEpiThreads); }
    mbar_init(bar_gfull, kEpiThreads);
    mbar_init(bar_gempty, 1);
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();   // peers' barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kAccCol = 256;

  // Chunk order on the ring (producer and MMA agree): S(0), S(1), B2(0), S(2), B2(1), ... , B2(T-1)
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_a, Cfg::kResident);
      for (int c = 0; c < KD; ++c) tma_load_2d(sm_a + c * kChunkBytes, &g.ta, bar_a, c * kChunkK, (int)i0);
      int st = 0; uint32_t ph = 0;
      auto push = [&](int col_chunk, int j0) {
        mbar_wait(bar_empty + st, ph ^ 1);
        ring_load<CS>(sm_ring + st * kChunkBytes, &g.tb, &g.tbp, bar_full + st, col_chunk * kChunkK, j0, cta_rank);
        if (++st == NST) { st = 0; ph ^= 1; }
      };
      for (int t = 0; t <= T; ++t) {
        if (t < T) {
          const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
          for (int c = 0; c < KD; ++c) push(c, j0);
        }
        if (t >= 1) {
          const int j0 = (int)(jlo + (int64_t)(t_begin + t - 1) * kTileRows);
          for (int dc = 0; dc < DNC; ++dc) push(h * DNC + dc, j0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_16(128, 128, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(128, 64, 0, 1, F16);  // B = streamed chunk, MN-major
    mbar_wait(bar_a, 0);
    tc_fence_after();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(sm_a), 16), b_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);
    int st = 0; uint32_t ph = 0;
    const uint32_t g_lo0 = umma_desc_lo(smem_u32(sm_g), 16);
    const uint32_t b2_lo0 = umma_desc_lo(smem_u32(sm_ring), kChunkBytes);
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        const int buf = t & 1;
        mbar_wait(bar_sempty + buf, ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll 1
        for (int c = 0; c < KD; ++c) {
          mbar_wait(bar_full + st, ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + c * (kChunkBytes >> 4);
          const uint32_t b_lo = b_lo0 + st * (kChunkBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma_bf16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc_s, (c | k) != 0);
            ring_release<CS>(bar_empty + st);
            if (c == KD - 1) umma_commit(bar_sfull + buf);
          }
          __syncwarp();
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
      if (t >= 1) {
        const int u = t - 1;
        mbar_wait(bar_gfull, u & 1);
        tc_fence_after();
#pragma unroll 1
        for (int dc = 0; dc < DNC; ++dc) {
          mbar_wait(bar_full + st, ph);
          tc_fence_after();
          const uint32_t b_lo = b2_lo0 + st * (kChunkBytes >> 4);
          const uint32_t d_tmem = tmem_base + kAccCol + dc * 64;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTileRows / kUmmaK; ++k) {
              // A = G[128 x 128] K-major: K 0..63 in sub-tile 0, 64..127 in sub-tile 1
              // B = chunk[128 j x 64 cols] read MN-major: 16 K-rows (j) = 2048 bytes per step
              umma_bf16_lo(d_tmem, g_lo0 + (k >> 2) * (kChunkBytes >> 4) + (k & 3) * 2, b_lo + k * (2048 >> 4),
                           idesc_g, (u | k) != 0);
            }
            ring_release<CS>(bar_empty + st);
            if (dc == DNC - 1) umma_commit(bar_gempty);
          }
          __syncwarp();
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
    }
    if (elect_one()) umma_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cc = (warp - 2) >> 2;          // which 32-column chunk of the tile this warp owns
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    const bool siglip = ga.bias != nullptr;
    if (i < n_rows) {
      bucket_range(gi, bs, n_cols, lo, hi);
      if (!siglip) rrs = 1.0f / g.rs[i];
    }
    const float s = expf(*ls);
    const float c1 = s * kLog2e, c0 = siglip ? *ga.bias * kLog2e : (kShiftK - s) * kLog2e;
    const bool want_gs = (g.gs != nullptr) && (h == 0);
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows;
      if (cc == 0 && !siglip) rcs_s[buf * 128 + r] = (j0 + r < n_cols) ? 1.0f / g.cs[j0 + r] : 0.f;
      named_barrier_sync(1, kEpiThreads);
      mbar_wait(bar_sfull + buf, (t >> 1) & 1);
      tc_fence_after();
      const bool full = (j0 + cc * 32 >= lo) && (j0 + cc * 32 + 32 <= hi);
      const bool warp_full = __all_sync(0xffffffffu, full);
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + buf * 128 + cc * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_sempty + buf);
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_s + buf * 128 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      mbar_wait(bar_gempty, (t & 1) ^ 1);  // MMA2 of the previous tile has read G
      // G[r][cc*32 .. +32): sub-tile cc/2, logical 16-byte chunks (cc%2)*4 .. +4, 128B swizzle
      const uint32_t grow = smem_u32(sm_g + (cc >> 1) * kChunkBytes + r * 128);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {   // explicit st.shared: a generic store goes through the global path (stall_lg)
        const int chunk = ((cc & 1) * 4 + c16) ^ (r & 7);
        sts_v4(grow + chunk * 16, packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_gfull);
    }
    // drain the resident accumulator: 2*DNC 32-column chunks over the four warps of a quadrant
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int ch = cc; ch < DNC * 2; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
      tmem_ld_wait();
      const int64_t col0 = (int64_t)h * DNC * 64 + ch * 32;
      if (i < n_rows) {
        float* dst = acc_out + i * d + col0;
        if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + e) =
                make_float4(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]),
                            __uint_as_float(raw[e + 2]), __uint_as_float(raw[e + 3]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col0 + e < d) dst[e] = __uint_as_float(raw[e]);
        }
      }
    }
    if (want_gs) {
      gs_local *= s;  // sum G * S with S = s * (u.v)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();   // no CTA leaves while a peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =============================================================================================
// backward, d <= 256: the streamed operand tile (<= 64 KiB) is kept in shared memory for BOTH
// GEMMs of a tile, and G never touches shared memory -- the epilogue warps overwrite the fp32
// logits tile in TMEM with the packed bf16 G (tcgen05.st), which the second tcgen05.mma reads as
// its A operand straight from tensor memory (B = the same tile viewed MN-major, N = d):
//     S(t)    = a . b_t^T                 SS-mode, D = logits buffer t&1
//     G(t)    = E (1/rs + 1/cs)           epilogue warps, in place
//     acc    += G(t) . b_t                TS-mode, D = resident accumulator
// Issue order S(0), S(1), GV(0), S(2), GV(1), ...: the in-order tensor pipe protects buffer t&1
// (S(t+2) is issued after GV(t)), so the only barriers are tile-full / G-ready / tile-free.
// =============================================================================================
template <int KD>
struct Grad2Cfg {
  static constexpr int kResident = KD * kChunkBytes;
  static constexpr int kTileBuf = KD * kChunkBytes;
  static constexpr int kSmem = 1024 + kResident + 2 * kTileBuf + kAuxBytes;
  static_assert(kSmem <= kMaxSmem, "tile-buffer backward needs d <= 256");
};

template <int KD, int CS, bool F16, bool SIG>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc2(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols,
    int64_t d, int64_t bs, int tiles_per_seg, const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = Grad2Cfg<KD>;
  constexpr int DN = KD * 64;  // accumulator columns = padded d
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_a = smem;
  uint8_t* sm_y = smem + Cfg::kResident;                  // [2][KD chunks]
  uint8_t* aux = sm_y + 2 * Cfg::kTileBuf;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(aux);     // [1]
  uint64_t* bar_yfull = bar_a + 1;                        // [2]
  uint64_t* bar_yempty = bar_yfull + 2;                   // [2]
  uint64_t* bar_sfull = bar_yempty + 2;                   // [2]
  uint64_t* bar_gfull = bar_sfull + 2;                    // [2]
  uint64_t* bar_accfull = bar_gfull + 2;                  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);     // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  int64_t jlo = 0, jhi = n_cols;
  if constexpr (CS == 1) row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0;
  float* acc_out = g.acc + (int64_t)blockIdx.x * n_rows * d;
  if (T == 0) {
    griddep_wait();   // global writes only after the kernel queued before this one is done
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DN; e += kNumThreads) {
      const int64_t rr = i0 + e / DN, col = e % DN;
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) TR(0);
  pdl_trigger();   // the gradient-tail kernel may become resident; it waits for this grid's completion

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tb);
    tma_prefetch_desc(&g.tbp);
    mbar_init(bar_a, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_yfull + b, 1);
      mbar_init(bar_yempty + b, CS);
      mbar_init(bar_sfull + b, 1);
      mbar_init(bar_gfull + b, kEpiThreads / 32);   // one elected arrival per epilogue warp
    }
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kAccCol = 256;
  if (threadIdx.x == 0) TR(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_a, Cfg::kResident);
      for (int c = 0; c < KD; ++c) tma_load_2d(sm_a + c * kChunkBytes, &g.ta, bar_a, c * kChunkK, (int)i0);
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
        mbar_wait(bar_yempty + b, ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_yfull + b, Cfg::kTileBuf);
        for (int c = 0; c < KD; ++c)
          chunk_load<CS>(sm_y + (b * KD + c) * kChunkBytes, &g.tb, &g.tbp, bar_yfull + b, c * kChunkK, j0, cta_rank);
        if (t < 16) TR(48 + t);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_16(128, 128, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(128, DN, 0, 1, F16);   // A from TMEM (K-major), B MN-major
    mbar_wait(bar_a, 0);
    tc_fence_after();
    if (lane == 0) TR(3);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(sm_a), 16);
    const uint32_t y_lo0 = umma_desc_lo(smem_u32(sm_y), 16);               // K-major view (S)
    const uint32_t y2_lo0 = umma_desc_lo(smem_u32(sm_y), kChunkBytes);     // MN-major view (G.V), LBO = chunk
    // Issue order: S(0) S(1) | GV(0) GV(1) S(2) S(3) | GV(2) GV(3) S(4) S(5) | ...
    // Tile buffer t&1 is released when GV(t) retires; issuing the two GVs back to back lets the
    // TMA refill of buffer 0 (~1000 cycles) overlap GV(2p+1) and the refill of buffer 1 overlap
    // S(2p+2), instead of stalling the in-order issue stream once per tile.
    for (int t0 = 0; t0 < T + 2; t0 += 2) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {          // GV of the previous pair
        const int u = t0 - 2 + h2;
        if (u < 0 || u >= T) continue;
        const int b = u & 1;
        mbar_wait(bar_gfull + b, (u >> 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = y2_lo0 + b * KD * (kChunkBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileRows / kUmmaK; ++k) {
            // A: packed G, K elements 32c..32c+31 live in columns 32c .. 32c+15 of logits buffer b
            const uint32_t a_tmem = tmem_base + b * 128 + (k >> 1) * 32 + (k & 1) * 8;
            umma_bf16_ts(tmem_base + kAccCol, a_tmem, b_lo + k * (2048 >> 4), idesc_g, (u | k) != 0);
          }
          ring_release<CS>(bar_yempty + b);   // tile buffer b (and logits buffer b) free once these retire
        }
        __syncwarp();
        if (lane == 0 && u < 16) TR(112 + u);
      }
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {          // S of this pair
        const int t = t0 + h2;
        if (t >= T) continue;
        const int b = t & 1;
        mbar_wait(bar_yfull + b, (t >> 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * 128;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < KD; ++c) {
            const uint32_t a_lo = a_lo0 + c * (kChunkBytes >> 4);
            const uint32_t b_lo = y_lo0 + (b * KD + c) * (kChunkBytes >> 4);
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma_bf16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc_s, (c | k) != 0);
          }
          umma_commit(bar_sfull + b);
        }
        __syncwarp();
        if (lane == 0 && t < 16) TR(64 + t);
      }
    }
    if (elect_one()) umma_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cc = (warp - 2) >> 2;          // which 32-column chunk of the tile this warp owns
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    constexpr bool siglip = SIG;   // compile-time here: the InfoNCE instantiation carries no extra state
    if (i < n_rows) {
      bucket_range(gi, bs, n_cols, lo, hi);
      if (!siglip) rrs = 1.0f / g.rs[i];
    }
    const float s = expf(*ls);
    const float c1 = s * kLog2e, c0 = siglip ? *ga.bias * kLog2e : (kShiftK - s) * kLog2e;
    const bool want_gs = g.gs != nullptr;
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // 1/cs of the NEXT tile is fetched one tile ahead (its global-load latency hides behind the
    // exp work of the current tile) and parked in the other half of rcs_s
    float rc_next = 0.f;
    if (cc == 0 && !siglip) {
      const int64_t jc = jlo + (int64_t)t_begin * kTileRows + r;
      rcs_s[r] = (jc < n_cols) ? 1.0f / g.cs[jc] : 0.f;
    }
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows;
      named_barrier_sync(1, kEpiThreads);   // rcs_s[buf] visible; everyone is done with tile t-1
      if (cc == 0 && t + 1 < T && !siglip) {
        const int64_t jc = j0 + kTileRows + r;
        rc_next = (jc < n_cols) ? g.cs[jc] : 0.f;
      }
      mbar_wait(bar_sfull + buf, (t >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && t < 16) TR(80 + t);
      const bool full = (j0 + cc * 32 >= lo) && (j0 + cc * 32 + 32 <= hi);
      const bool warp_full = __all_sync(0xffffffffu, full);
      const uint32_t col0 = tmem_base + lane_addr + buf * 128 + cc * 32;
      uint32_t raw[32];
      tmem_ld32(col0, raw);
      tmem_ld_wait();
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_s + buf * 128 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      // G overwrites the first 16 of this warp's own 32 logits columns (all 32 were read above)
      tmem_st16(col0, packed);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull + buf);   // 16 arrivals per tile instead of 512 serialized smem atomics
      if (threadIdx.x == 64 && t < 16) TR(96 + t);
      if (cc == 0 && t + 1 < T) rcs_s[(buf ^ 1) * 128 + r] = (rc_next != 0.f) ? 1.0f / rc_next : 0.f;
    }
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TR(2);
    // With an overlapped launch (launch_kernel_ex) everything above only read what the forward left
    // behind; global writes start here, after the kernel queued before this one has completed.
    griddep_wait();
    if (g.use_tacc) {
      // Drain through shared memory + TMA stores: a thread owns a row, so direct stores would touch 32
      // different lines per warp instruction (8192 sixteen-byte requests per CTA); staged as 128-byte
      // swizzled rows in the (now idle) tile buffers, each 32-column chunk leaves as ONE bulk store of
      // full lines.  The four warps of a chunk (lane quadrants 0..3) synchronise on a named barrier.
#pragma unroll 1
      for (int ch = cc; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        uint8_t* stage = sm_y + ch * kChunkBytes;     // 128 rows x 128 B
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
          sts_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4), raw[v4 * 4], raw[v4 * 4 + 1], raw[v4 * 4 + 2], raw[v4 * 4 + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(2 + cc, 128);
        if (q == 0 && lane == 0 && i0 < n_rows) {
          tma_store_3d(&g.tacc, stage, ch * 32, (int)i0, (int)blockIdx.x);
          tma_store_commit();
        }
      }
      if (q == 0 && lane == 0) tma_store_wait_read();
    } else {
#pragma unroll 1
    for (int ch = cc; ch < 2 * KD; ch += 4) {
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
      tmem_ld_wait();
      const int64_t col0 = (int64_t)ch * 32;
      if (i < n_rows) {
        float* dst = acc_out + i * d + col0;
        if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + e) =
                make_float4(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]),
                            __uint_as_float(raw[e + 2]), __uint_as_float(raw[e + 3]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col0 + e < d) dst[e] = __uint_as_float(raw[e]);
        }
      }
    }
    }
    if (want_gs) {
      gs_local *= s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);   // zeroed by the forward's last kernel (waited for above)
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
  }
  griddep_wait();   // no thread block outlives the kernel queued before it (see launch_kernel_ex)
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TR(5);
  if constexpr (CS > 1) cluster_sync_exec();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) TR(6);
}


// =============================================================================================
// backward, d <= 256, BOTH GEMMs with the A operand in tensor memory (TS mode).
//
// infonce_grad_tc2 above reads the owned rows from shared memory for S = a . b_t^T: an SS-mode
// 128x128x16 MMA pulls 8 KiB of operands per 64 cycles -- all of the 128 B/clk a shared memory delivers --
// and the TMA refill of the tile buffers comes on top, so a 128-column tile took 3100 cycles for 2048 cycles
// of tensor work.  Here the owned rows are parked in TMEM once (as in the forward) and the logits are
// produced 64 columns at a time, which is what makes everything fit in the 512 TMEM columns:
//     [0, 64)     S        fp32 logits of the current 64-column tile            (one buffer)
//     [64, 128)   G ring   2 x 32 columns: packed 16-bit weights of tiles t, t+1 (A operand of G . b_t)
//     [128, 256)  A        owned rows, packed 16-bit pairs, KD x 32 columns
//     [256, 512)  acc      [128 x d] fp32, resident for the whole sweep
// Shared memory only holds streamed tiles ([64 x d] 16-bit, a ring of six) and is read at 64 B/clk by either
// MMA (+ 32 B/clk of TMA refill).  Issue order S(0) | S(1) GV(0) | S(2) GV(1) | ... on the in-order tensor
// pipe: the epilogue of tile t (tcgen05.ld, one exp per logit, tcgen05.st of G) has the whole of GV(t-1) and
// S(t+1) -- 1024 cycles -- before GV(t) needs its result, and it releases the single S buffer right after its
// tcgen05.ld, long before S(t+1) is due.  The sixteen epilogue warps form two groups that take alternate tiles
// (8 warps = 4 lane quadrants x 2 chunks of 32 columns), so a warp has two tile periods for one chunk.
// G ring slot t&1 is rewritten by the same group two tiles later, after S(t+2) -- issued behind GV(t) -- has
// completed: no extra barrier.  A tile slot is released by the commit that follows GV(t).  No cluster /
// multicast here: at 32 B/clk per SM the streamed tiles use half of the L2 bandwidth, and pairing row blocks
// measured as neutral in infonce_grad_tc2.
// =============================================================================================
constexpr int kG3Slots = 6;
constexpr int kG3Cols = 64;                     // logits columns per tile
constexpr int kG3ChunkBytes = kChunkBytes / 2;  // [64 x 64] 16-bit = 8 KiB
template <int KD>
struct Grad3Cfg {
  static constexpr int kSlotBytes = KD * kG3ChunkBytes;
  static constexpr int kRing = kG3Slots * kSlotBytes;
  static constexpr int kSmem = 1024 + kRing + kAuxBytes;
  static constexpr int kAStage = (kG3Slots - 2) * kSlotBytes;   // owned rows (KD x 16 KiB) arrive in the last two slots
  static_assert(kSmem <= kMaxSmem, "ring too large");
  static_assert(4 * kSlotBytes >= kTileRows * KD * 64 * 4, "accumulator drain is staged in four slots");
};

template <int KD, bool F16, bool SIG>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc3(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols,
    int64_t d, int64_t bs, int tiles_per_seg, const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = Grad3Cfg<KD>;
  constexpr int NSL = kG3Slots;
  constexpr int DN = KD * 64;  // accumulator columns = padded d
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_ring = smem;
  uint8_t* aux = sm_ring + Cfg::kRing;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);   // [NSL] tile landed
  uint64_t* bar_empty = bar_full + NSL;                     // [NSL] tile consumed by S and G.V
  uint64_t* bar_afull = bar_empty + NSL;                    // [1] owned rows landed in their staging slots
  uint64_t* bar_a = bar_afull + 1;                          // [1] owned rows parked in TMEM
  // One barrier per epilogue group (= tile parity) for each hand-over: a parity wait only tells "the phase
  // before the current one is complete", so two groups waiting on ONE barrier for alternate completions
  // would see each other's phases.
  uint64_t* bar_sfull = bar_a + 1;                          // [2] logits tile t complete            (index t&1)
  uint64_t* bar_sempty = bar_sfull + 2;                     // [2] logits tile t read by its group    (index t&1)
  uint64_t* bar_gfull = bar_sempty + 2;                     // [2] packed weights of tile t in ring slot t&1
  uint64_t* bar_accfull = bar_gfull + 2;                    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);       // [2 groups][2 buffers][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  int64_t jlo = 0, jhi = n_cols;
  row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kG3Cols - 1) / kG3Cols);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  float* acc_out = g.acc + (int64_t)blockIdx.x * n_rows * d;
  if (T == 0) {
    griddep_wait();   // global writes only after the kernel queued before this one is done
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DN; e += kNumThreads) {
      const int64_t rr = i0 + e / DN, col = e % DN;
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) TR(0);
  pdl_trigger();   // the gradient-tail kernel may become resident; it waits for this grid's completion

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tbp);
    for (int s = 0; s < NSL; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    mbar_init(bar_afull, 1);
    mbar_init(bar_a, kEpiThreads);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_sfull + b, 1);
      mbar_init(bar_sempty + b, kEpiThreads / 64);      // one elected arrival per warp of the tile's group
      mbar_init(bar_gfull + b, kEpiThreads / 64);
    }
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kSCol = 0, kGCol = 64, kACol = 128, kAccCol = 256;
  if (threadIdx.x == 0) TR(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_afull, KD * kChunkBytes);
      for (int c = 0; c < KD; ++c)
        tma_load_2d(sm_ring + Cfg::kAStage + c * kChunkBytes, &g.ta, bar_afull, c * kChunkK, (int)i0);
      bool a_parked = false;
      int st = 0; uint32_t ph = 0;
      for (int t = 0; t < T; ++t) {
        const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kG3Cols);
        if (!a_parked && st >= NSL - 2) {   // first use of a slot that overlaps the staged owned rows
          mbar_wait(bar_a, 0);
          a_parked = true;
        }
        mbar_wait(bar_empty + st, ph ^ 1);
        mbar_expect_tx(bar_full + st, Cfg::kSlotBytes);
        uint8_t* slot = sm_ring + st * Cfg::kSlotBytes;
#pragma unroll
        for (int c = 0; c < KD; ++c)   // g.tbp: boxes of 64 rows
          tma_load_2d(slot + c * kG3ChunkBytes, &g.tbp, bar_full + st, c * kChunkK, j0);
        if (++st == NSL) { st = 0; ph ^= 1; }
        if (t < 16) TR(48 + t);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_16(128, kG3Cols, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(128, DN, 0, 1, F16);   // A from TMEM (K-major), B MN-major
    mbar_wait(bar_a, 0);     // the epilogue warps have parked the owned rows in TMEM
    tc_fence_after();
    if (lane == 0) TR(3);
    const uint32_t s_tmem = tmem_base + kSCol, a_tmem0 = tmem_base + kACol, acc_tmem = tmem_base + kAccCol;
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);              // K-major view (S)
    const uint32_t mn_lo0 = umma_desc_lo(smem_u32(sm_ring), kG3ChunkBytes);  // MN-major view (G.V): LBO = chunk
    int st = 0; uint32_t ph = 0;   // slot of tile t (S side)
    int st_g = 0;                  // slot of tile t-1 (G.V side)
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        mbar_wait(bar_full + st, ph);
        if (t >= 1) mbar_wait(bar_sempty + ((t - 1) & 1), ((t - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = k_lo0 + st * (Cfg::kSlotBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < KD; ++c)
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma_bf16_ts(s_tmem, a_tmem0 + c * 32 + k * 8, b_lo + c * (kG3ChunkBytes >> 4) + 2 * k, idesc_s,
                           (c | k) != 0);
          umma_commit(bar_sfull + (t & 1));
        }
        __syncwarp();
        if (lane == 0 && t < 16) TR(64 + t);
        if (++st == NSL) { st = 0; ph ^= 1; }
      }
      if (t >= 1) {
        const int u = t - 1;
        mbar_wait(bar_gfull + (u & 1), (u >> 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = mn_lo0 + st_g * (Cfg::kSlotBytes >> 4);
        const uint32_t g_tmem = tmem_base + kGCol + (u & 1) * 32;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kG3Cols / kUmmaK; ++k)   // A: 8 packed columns per K step; B: 16 tile rows = 2 KiB
            umma_bf16_ts(acc_tmem, g_tmem + k * 8, b_lo + k * (2048 >> 4), idesc_g, (u | k) != 0);
          umma_commit(bar_empty + st_g);   // the tile slot is free once S(u) and G.V(u) have retired
        }
        __syncwarp();
        if (lane == 0 && u < 16) TR(112 + u);
        if (++st_g == NSL) st_g = 0;
      }
    }
    if (elect_one()) umma_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int e = (warp - 2) >> 2;           // 0..3
    const int grp = e >> 1;                  // tiles t = grp (mod 2)
    const int cc = e & 1;                    // which 32-column chunk of the 64-column tile
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    constexpr bool siglip = SIG;
    if (i < n_rows) {
      bucket_range(gi, bs, n_cols, lo, hi);
      if (!siglip) rrs = 1.0f / g.rs[i];
    }
    const float s = expf(*ls);
    const float c1 = s * kLog2e, c0 = siglip ? *ga.bias * kLog2e : (kShiftK - s) * kLog2e;
    const bool want_gs = g.gs != nullptr;
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    {  // park this thread's owned row (16-bit operand, padded to KD*64) in TMEM as packed pairs: the A operand
       // of S.  Source: the swizzled staging slots (rows past n_rows were zero-filled by the TMA unit).
      mbar_wait(bar_afull, 0);
      for (int c = e; c < KD; c += 4) {
        const uint8_t* rowp = sm_ring + Cfg::kAStage + c * kChunkBytes + r * 128;
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const uint4 w = lds_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4));
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + lane_addr + kACol + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a);
      if (threadIdx.x == 64) TR(2);
    }
    // 1/cs of the group's NEXT tile is fetched one tile ahead and parked in the other buffer of rcs_s
    const bool filler = (cc == 0) && (q < 2);          // 64 threads per group: one per tile column
    float* rcs_g = rcs_s + grp * 128;
    float rc_next = 0.f;
    if (filler && !siglip && grp < T) {
      const int64_t jc = jlo + (int64_t)(t_begin + grp) * kG3Cols + r;
      rcs_g[r] = (jc < n_cols) ? 1.0f / g.cs[jc] : 0.f;
    }
    for (int t = grp; t < T; t += 2) {
      const int it = t >> 1;                           // tile ordinal inside the group
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kG3Cols;
      named_barrier_sync(1 + grp, kEpiThreads / 2);    // rcs of this tile visible; the group is done with tile t-2
      if (filler && !siglip && t + 2 < T) {
        const int64_t jc = j0 + 2 * kG3Cols + r;
        rc_next = (jc < n_cols) ? g.cs[jc] : 0.f;
      }
      mbar_wait(bar_sfull + grp, it & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && t < 16) TR(80 + t);
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kSCol + cc * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + grp);    // the logits buffer goes back to the MMA warp right away
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_g + (it & 1) * 64 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      tmem_st16(tmem_base + lane_addr + kGCol + (t & 1) * 32 + cc * 16, packed);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull + grp);
      if (threadIdx.x == 64 && t < 16) TR(96 + t);
      if (filler && !siglip && t + 2 < T) rcs_g[((it + 1) & 1) * 64 + r] = (rc_next != 0.f) ? 1.0f / rc_next : 0.f;
    }
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TR(2);
    // With an overlapped launch everything above only read what the forward left behind; global writes
    // start here, after the kernel queued before this one has completed.
    griddep_wait();
    if (g.use_tacc) {
      // Drain through shared memory + TMA stores (see infonce_grad_tc2): 128-byte swizzled rows staged in the
      // idle tile slots, one bulk store of full lines per 32-column chunk.
#pragma unroll 1
      for (int ch = e; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        uint8_t* stage = sm_ring + ch * kChunkBytes;     // 128 rows x 128 B
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
          sts_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4), raw[v4 * 4], raw[v4 * 4 + 1], raw[v4 * 4 + 2], raw[v4 * 4 + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(3 + e, 128);
        if (q == 0 && lane == 0 && i0 < n_rows) {
          tma_store_3d(&g.tacc, stage, ch * 32, (int)i0, (int)blockIdx.x);
          tma_store_commit();
        }
      }
      if (q == 0 && lane == 0) tma_store_wait_read();
    } else {
#pragma unroll 1
      for (int ch = e; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        const int64_t col0 = (int64_t)ch * 32;
        if (i < n_rows) {
          float* dst = acc_out + i * d + col0;
          if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              *reinterpret_cast<float4*>(dst + x) =
                  make_float4(__uint_as_float(raw[x]), __uint_as_float(raw[x + 1]),
                              __uint_as_float(raw[x + 2]), __uint_as_float(raw[x + 3]));
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
          }
        }
      }
    }
    if (want_gs) {
      gs_local *= s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);   // zeroed by the forward's last kernel (waited for above)
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
  }
  griddep_wait();   // no thread block outlives the kernel queued before it (see launch_kernel_ex)
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TR(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) TR(6);
}


// =============================================================================================
// backward, d <= 256, logits in TS mode at N = 128 and G through shared memory ("tc4").
//
// Measured on B200 (tools/trace_tc.py): a TS-mode tcgen05.mma takes 64 cycles however small N is, and an
// SS-mode 128x128x16 MMA 128 cycles (both operands through the 64 B/clk shared-memory operand port).  So
// S = a . b_t^T costs 16 x 128 cycles per 128-column tile with the owned rows in shared memory
// (infonce_grad_tc2: 3100 cycles per tile together with G . b_t), and 2 x 16 x 64 when the columns are
// halved to make room in TMEM (infonce_grad_tc3: the same 3070).  Here the owned rows sit in TMEM and S
// keeps N = 128 (16 x 64 = 1024 cycles); what no longer fits in TMEM is a second logits buffer or G, so the
// epilogue hands the logits buffer back right after its tcgen05.ld and writes the packed 16-bit G to a
// 32 KiB shared-memory buffer, from where G . b_t runs in SS mode against the whole [128 x d] tile as ONE
// MN-major operand (8 MMAs of N = d, ~1100 cycles).
//     TMEM  [0,128) S | [128,256) owned rows (packed pairs) | [256,512) acc [128 x d] fp32
//     smem  3 tile buffers [128 x d] (a 64 KiB tile takes ~1450 cycles to arrive and is held from S(t) to the
//           end of GV(t): with two buffers the refill was exposed every tile) | G [128 x 128] 16-bit
//           (3 x 64 + 32 KiB + 2 KiB of barriers = exactly the 227 KiB a CTA may have at d = 256)
// Issue order S(0) | S(1) GV(0) | S(2) GV(1) | ...: S(t+1) needs the logits buffer back (tcgen05.ld of tile
// t, ~400 cycles into GV(t-1)); GV(t) needs G(t), which the epilogue stores after GV(t-1) has released the
// G buffer; the commit behind GV(t) also releases tile buffer t % 3 for tile t + 3.
// =============================================================================================
constexpr int kG4Aux = 2048;
template <int KD>
struct Grad4Cfg {
  static constexpr int kTileBuf = KD * kChunkBytes;
  static constexpr int kNBuf = 3;
  static constexpr int kGBuf = 2 * kChunkBytes;            // [128 x 128] 16-bit as two K-major [128 x 64] sub-tiles
  static constexpr int kSmem = 1024 + kNBuf * kTileBuf + kGBuf + kG4Aux;
  static_assert(kSmem <= kMaxSmem, "tc4 needs d <= 256");
};

template <int KD, bool F16, bool SIG>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc4(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols,
    int64_t d, int64_t bs, int tiles_per_seg, const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = Grad4Cfg<KD>;
  constexpr int NB = Cfg::kNBuf;
  constexpr int DN = KD * 64;  // accumulator columns = padded d
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_y = smem;                                    // [NB][KD chunks]
  uint8_t* sm_g = smem + NB * Cfg::kTileBuf;               // [2 sub-tiles]
  uint8_t* aux = sm_g + Cfg::kGBuf;
  uint64_t* bar_afull = reinterpret_cast<uint64_t*>(aux);  // [1] owned rows landed in the last tile buffer
  uint64_t* bar_a = bar_afull + 1;                         // [1] owned rows parked in TMEM
  uint64_t* bar_yfull = bar_a + 1;                         // [NB] tile landed
  uint64_t* bar_yempty = bar_yfull + NB;                   // [NB] tile consumed by S and G.V
  uint64_t* bar_sfull = bar_yempty + NB;                   // [1] logits complete
  uint64_t* bar_sempty = bar_sfull + 1;                    // [1] logits read by the epilogue
  uint64_t* bar_gfull = bar_sempty + 1;                    // [1] G written
  uint64_t* bar_gempty = bar_gfull + 1;                    // [1] G consumed by G.V
  uint64_t* bar_accfull = bar_gempty + 1;                  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);      // [2][128]
  uint8_t* sm_astage = sm_y + (NB - 1) * Cfg::kTileBuf;    // owned rows arrive here (first used by tile NB-1)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  int64_t jlo = 0, jhi = n_cols;
  row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  float* acc_out = g.acc + (int64_t)blockIdx.x * n_rows * d;
  if (T == 0) {
    griddep_wait();   // global writes only after the kernel queued before this one is done
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DN; e += kNumThreads) {
      const int64_t rr = i0 + e / DN, col = e % DN;
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) TR(0);
  pdl_trigger();   // the gradient-tail kernel may become resident; it waits for this grid's completion

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tb);
    mbar_init(bar_afull, 1);
    mbar_init(bar_a, kEpiThreads);
    for (int b = 0; b < NB; ++b) {
      mbar_init(bar_yfull + b, 1);
      mbar_init(bar_yempty + b, 1);
    }
    mbar_init(bar_sfull, 1);
    mbar_init(bar_sempty, kEpiThreads / 32);   // one elected arrival per epilogue warp
    mbar_init(bar_gfull, kEpiThreads / 32);
    mbar_init(bar_gempty, 1);
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kSCol = 0, kACol = 128, kAccCol = 256;
  if (threadIdx.x == 0) TR(1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_afull, KD * kChunkBytes);
      for (int c = 0; c < KD; ++c) tma_load_2d(sm_astage + c * kChunkBytes, &g.ta, bar_afull, c * kChunkK, (int)i0);
      int b = 0; uint32_t ph = 0;
      for (int t = 0; t < T; ++t) {
        const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
        if (t == NB - 1) mbar_wait(bar_a, 0);   // the staged owned rows have left the last buffer
        mbar_wait(bar_yempty + b, ph ^ 1);
        mbar_expect_tx(bar_yfull + b, Cfg::kTileBuf);
        for (int c = 0; c < KD; ++c)
          tma_load_2d(sm_y + (b * KD + c) * kChunkBytes, &g.tb, bar_yfull + b, c * kChunkK, j0);
        if (t < 16) TR(48 + t);
        if (++b == NB) { b = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_16(128, 128, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(128, DN, 0, 1, F16);   // A = G (K-major, smem), B = tile MN-major
    mbar_wait(bar_a, 0);     // the epilogue warps have parked the owned rows in TMEM
    tc_fence_after();
    if (lane == 0) TR(3);
    const uint32_t s_tmem = tmem_base + kSCol, a_tmem0 = tmem_base + kACol, acc_tmem = tmem_base + kAccCol;
    const uint32_t y_lo0 = umma_desc_lo(smem_u32(sm_y), 16);               // K-major view (S)
    const uint32_t y2_lo0 = umma_desc_lo(smem_u32(sm_y), kChunkBytes);     // MN-major view (G.V), LBO = chunk
    const uint32_t g_lo0 = umma_desc_lo(smem_u32(sm_g), 16);
    int b = 0; uint32_t ph = 0;   // buffer / phase of tile t (S side)
    int bg = 0;                   // buffer of tile t-1 (G.V side)
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        mbar_wait(bar_yfull + b, ph);
        if (t >= 1) mbar_wait(bar_sempty, (t - 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = y_lo0 + b * KD * (kChunkBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < KD; ++c)
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma_bf16_ts(s_tmem, a_tmem0 + c * 32 + k * 8, b_lo + c * (kChunkBytes >> 4) + 2 * k, idesc_s,
                           (c | k) != 0);
          umma_commit(bar_sfull);
        }
        __syncwarp();
        if (lane == 0 && t < 16) TR(64 + t);
        if (++b == NB) { b = 0; ph ^= 1; }
      }
      if (t >= 1) {
        const int u = t - 1;
        mbar_wait(bar_gfull, u & 1);
        tc_fence_after();
        const uint32_t b_lo = y2_lo0 + bg * KD * (kChunkBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileRows / kUmmaK; ++k)
            // A = G[128 x 128] K-major: K 0..63 in sub-tile 0, 64..127 in sub-tile 1 (32 bytes per K step)
            // B = tile[128 j x d] read MN-major: 16 K-rows (j) = 2048 bytes per step
            umma_bf16_lo(acc_tmem, g_lo0 + (k >> 2) * (kChunkBytes >> 4) + (k & 3) * 2, b_lo + k * (2048 >> 4), idesc_g,
                         (u | k) != 0);
          umma_commit(bar_yempty + bg);   // tile buffer free once S(u) and G.V(u) have retired
          umma_commit(bar_gempty);        // ... and so is the G buffer
        }
        __syncwarp();
        if (lane == 0 && u < 16) TR(112 + u);
        if (++bg == NB) bg = 0;
      }
    }
    if (elect_one()) umma_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cc = (warp - 2) >> 2;          // which 32-column chunk of the tile this warp owns
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    constexpr bool siglip = SIG;
    if (i < n_rows) {
      bucket_range(gi, bs, n_cols, lo, hi);
      if (!siglip) rrs = 1.0f / g.rs[i];
    }
    const float s = expf(*ls);
    const float c1 = s * kLog2e, c0 = siglip ? *ga.bias * kLog2e : (kShiftK - s) * kLog2e;
    const bool want_gs = g.gs != nullptr;
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    {  // park this thread's owned row in TMEM as packed pairs (A operand of S); source: the staged rows
      mbar_wait(bar_afull, 0);
      for (int c = cc; c < KD; c += 4) {
        const uint8_t* rowp = sm_astage + c * kChunkBytes + r * 128;
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const uint4 w = lds_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4));
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + lane_addr + kACol + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a);
      if (threadIdx.x == 64) TR(2);
    }
    // 1/cs of the NEXT tile is fetched one tile ahead and parked in the other half of rcs_s
    float rc_next = 0.f;
    if (cc == 0 && !siglip) {
      const int64_t jc = jlo + (int64_t)t_begin * kTileRows + r;
      rcs_s[r] = (jc < n_cols) ? 1.0f / g.cs[jc] : 0.f;
    }
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows;
      if (threadIdx.x == 64 && t == 4) TR(14);
      named_barrier_sync(1, kEpiThreads);   // rcs_s[buf] visible; everyone is done with tile t-1
      if (threadIdx.x == 64 && t == 4) TR(15);
      if (cc == 0 && t + 1 < T && !siglip) {
        const int64_t jc = j0 + kTileRows + r;
        rc_next = (jc < n_cols) ? g.cs[jc] : 0.f;
      }
      mbar_wait(bar_sfull, t & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && t < 16) TR(80 + t);
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kSCol + cc * 32, raw);
      tmem_ld_wait();
      if (threadIdx.x == 64 && t == 4) TR(8);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty);   // the logits buffer goes back to the MMA warp right away
      if (threadIdx.x == 64 && t == 4) TR(9);
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_s + buf * 128 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      if (threadIdx.x == 64 && t == 4) TR(10);
      if (t >= 1) mbar_wait(bar_gempty, (t - 1) & 1);   // G.V of the previous tile has read the G buffer
      if (threadIdx.x == 64 && t == 4) TR(11);
      // G[r][cc*32 .. +32): sub-tile cc/2, logical 16-byte chunks (cc%2)*4 .. +4, 128B swizzle
      const uint32_t grow = smem_u32(sm_g + (cc >> 1) * kChunkBytes + r * 128);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {   // explicit st.shared: a generic store goes through the global path (stall_lg)
        const int chunk = ((cc & 1) * 4 + c16) ^ (r & 7);
        sts_v4(grow + chunk * 16, packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
      }
      if (threadIdx.x == 64 && t == 4) TR(12);
      fence_proxy_async_smem();
      if (threadIdx.x == 64 && t == 4) TR(13);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull);
      if (threadIdx.x == 64 && t < 16) TR(96 + t);
      if (cc == 0 && t + 1 < T) rcs_s[(buf ^ 1) * 128 + r] = (rc_next != 0.f) ? 1.0f / rc_next : 0.f;
    }
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TR(2);
    griddep_wait();   // global writes start here (see infonce_grad_tc2)
    if (g.use_tacc) {
#pragma unroll 1
      for (int ch = cc; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        uint8_t* stage = sm_y + ch * kChunkBytes;     // 128 rows x 128 B in the idle tile buffers
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
          sts_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4), raw[v4 * 4], raw[v4 * 4 + 1], raw[v4 * 4 + 2], raw[v4 * 4 + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(2 + cc, 128);
        if (q == 0 && lane == 0 && i0 < n_rows) {
          tma_store_3d(&g.tacc, stage, ch * 32, (int)i0, (int)blockIdx.x);
          tma_store_commit();
        }
      }
      if (q == 0 && lane == 0) {
        if (ga.tail.enabled) tma_store_wait_all();   // the slab must be complete before this segment is counted
        else tma_store_wait_read();
      }
    } else {
#pragma unroll 1
      for (int ch = cc; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        const int64_t col0 = (int64_t)ch * 32;
        if (i < n_rows) {
          float* dst = acc_out + i * d + col0;
          if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              *reinterpret_cast<float4*>(dst + x) =
                  make_float4(__uint_as_float(raw[x]), __uint_as_float(raw[x + 1]),
                              __uint_as_float(raw[x + 2]), __uint_as_float(raw[x + 3]));
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
          }
        }
      }
    }
    if (want_gs) {
      gs_local *= s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);   // zeroed by the forward's last kernel (waited for above)
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
    if constexpr (!SIG && (KD % 2 == 0)) {
      if (ga.tail.enabled) {
        // ---- fused gradient tail: count this segment; the last one of the (direction, row block) finishes it
        const GradTail& tl = ga.tail;
        const int kdir = blockIdx.z % ga.ndir;
        int* flag_s = reinterpret_cast<int*>(aux + 256);
        named_barrier_sync(1, kEpiThreads);      // every warp's stores are complete, every sum G*S is added
        if (threadIdx.x == 64) {
          __threadfence();
          const int old = atomicAdd(tl.counters + kdir * gridDim.y + blockIdx.y, 1);
          *flag_s = (old == tl.nseg - 1);
        }
        named_barrier_sync(1, kEpiThreads);
        if (*flag_s) {
          __threadfence();
          constexpr int NV = KD / 2;
          const float go = (*tl.grad_out_emb) * tl.emb_scale;
          const float* xo = tl.x[kdir];
          const float* xp = tl.x[1 - kdir];
          const int64_t slab4 = n_rows * d / 4;
#pragma unroll 1
          for (int rr = warp - 2; rr < kTileRows; rr += kEpiThreads / 32) {
            const int64_t row = i0 + rr;
            if (row >= n_rows) break;
            float coef, dterm;
            tail_terms(nullptr, tl.diag[row], tl.R, tl.C, row, s, go, tl.batch, coef, dterm);
            finish_row_vec<NV>(g.acc + row * d, tl.nseg, slab4, xo + row * tl.ldx, xp + row * tl.ldx, coef, dterm,
                               tl.inv_den[kdir][row], tl.inv_den[1 - kdir][row], !(tl.nrm[kdir][row] > kNormEps),
                               tl.dx[kdir] + row * d, lane);
          }
          named_barrier_sync(1, kEpiThreads);
          if (threadIdx.x == 64) {
            const int total = ga.ndir * (int)gridDim.y;
            __threadfence();
            const int done = atomicAdd(tl.counters + total, 1);
            tl.counters[kdir * gridDim.y + blockIdx.y] = 0;          // left zero for the next backward
            if (done == total - 1) {   // the last tail of the grid: every sum G*S has been added
              __threadfence();
              const float gsv = *reinterpret_cast<volatile float*>(ga.dir[0].gs);
              *tl.dls_out = (float)((double)(*tl.grad_out) / (2.0 * (double)tl.batch) *
                                    ((double)gsv - 2.0 * (double)(*tl.diag_sum)));
              *ga.dir[0].gs = 0.f;     // consumed: the accumulator is back to its zero-initialised state
              tl.counters[total] = 0;
            }
          }
        }
      }
    }
  }
  griddep_wait();   // no thread block outlives the kernel queued before it (see launch_kernel_ex)
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TR(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) TR(6);
}

// =============================================================================================
// backward, d = 128 or 256, one MMA over a PAIR of row blocks ("tc5", tcgen05.mma.cta_group::2).
//
// infonce_grad_tc4 is bound by the shared-memory operand port: G . b_t reads G (4 KiB per K step) and the tile
// (8 KiB per K step) at 64 B/clk = 192 cycles per MMA, 1536 per tile, after 1024 for S.  With cta_group::2
// the two CTAs of a cluster own row blocks 2p and 2p+1 and every MMA has M = 256: each CTA contributes ITS
// 128 rows (A: TMEM for S, its G buffer for G . b_t), the B operand is split along N between the two shared
// memories, and each CTA's tensor pipe runs its own 128 rows against the whole of B.  Per CTA and K step that
// halves the B bytes through the operand port: S reads 64 of the 128 logits columns (TS mode: 64 cycles
// either way), G . b_t reads 4 KiB of G + 4 KiB of tile = 128 cycles: 1024 + 1024 per tile instead of 2560.
//   per tile each CTA loads  KD x [64 j x 64 k]   its half of the tile's ROWS (= logits columns), all of d  (S)
//                            KD/2 x [128 j x 64]  all rows of the tile, its half of d                      (G . b_t)
//   (the same 64 KiB per tile and CTA as tc4 at d = 256; a quarter of it is fetched twice, from L2)
//   TMEM (both CTAs, allocated as a pair)  [0,128) S | [128,256) owned rows | [256,512) acc [128 x d]
// The leader (cluster rank 0) issues every MMA; its commits are multicast to the barrier of the same name in
// both CTAs; the epilogue warps of both CTAs arrive on the LEADER's sempty / gfull barriers; the idle MMA
// warp of the peer forwards "my operands have landed" to the leader.
// Needs every row block to sweep the same columns (single bucket), like the multicast clusters.
// =============================================================================================
template <int KD>
struct Grad5Cfg {
  static_assert(KD == 2 || KD == 4, "tc5: d = 128 or 256");
  static constexpr int KH = KD / 2;
  static constexpr int kSBytes = KD * (kChunkBytes / 2);   // [KD][64 x 64]
  static constexpr int kVBytes = KH * kChunkBytes;         // [KH][128 x 64]
  // Two rings instead of three whole-tile buffers (same 3 * KD * 16 KiB): the logits half of a tile is dead as soon
  // as S(t) has retired, the G.V half lives until G.V(t) has -- a whole-tile buffer was held for both, and every
  // third tile waited ~500 cycles for its operands (G.V(t) done -> load of tile t+3 -> ~1450 cycles to arrive).
  static constexpr int kNS = 2;                            // logits operands: S(t) .. S(t+1)
  static constexpr int kNV = 4;                            // G.V operands: held from arrival to the end of G.V(t)
  static constexpr int kRingBytes = kNS * kSBytes + kNV * kVBytes;
  static constexpr int kGBuf = 2 * kChunkBytes;
  static constexpr int kSmem = 1024 + kRingBytes + kGBuf + kG4Aux;
  static_assert(kSmem <= kMaxSmem, "tc5 needs d <= 256");
  static_assert(2 * kVBytes == KD * kChunkBytes, "the owned rows are staged in the last two G.V slots");
  static_assert(kRingBytes >= 2 * KD * kChunkBytes, "accumulator drain staging");
};

template <int KD, bool F16, bool SIG>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc5(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols,
    int64_t d, int64_t bs, int tiles_per_seg, const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = Grad5Cfg<KD>;
  constexpr int NS = Cfg::kNS, NV = Cfg::kNV, KH = Cfg::KH;
  constexpr int DN = KD * 64;  // accumulator columns = padded d
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_y = smem;                                    // (also the accumulator drain's staging area)
  uint8_t* sm_s = smem;                                    // [NS] logits operands   [KD][64 x 64]
  uint8_t* sm_v = smem + NS * Cfg::kSBytes;                // [NV] G.V operands      [KH][128 x 64]
  uint8_t* sm_g = smem + Cfg::kRingBytes;                  // [2 sub-tiles]
  uint8_t* aux = sm_g + Cfg::kGBuf;
  uint64_t* bar_afull = reinterpret_cast<uint64_t*>(aux);  // [1] own rows landed in the last tile buffer
  uint64_t* bar_aloc = bar_afull + 1;                      // [1] own rows parked in TMEM (this CTA's warps)
  uint64_t* bar_aall = bar_aloc + 1;                       // [1] leader: both CTAs' rows parked
  uint64_t* bar_yfull = bar_aall + 1;                      // [NS] own logits operands landed
  uint64_t* bar_pfull = bar_yfull + NS;                    // [NS] leader: the peer's logits operands landed
  uint64_t* bar_yempty = bar_pfull + NS;                   // [NS] logits operands consumed (multicast commit)
  uint64_t* bar_vfull = bar_yempty + NS;                   // [NV] own G.V operands landed
  uint64_t* bar_pvfull = bar_vfull + NV;                   // [NV] leader: the peer's G.V operands landed
  uint64_t* bar_vempty = bar_pvfull + NV;                  // [NV] G.V operands consumed (multicast commit)
  uint64_t* bar_sfull = bar_vempty + NV;                   // [1] logits complete (multicast commit)
  uint64_t* bar_sempty = bar_sfull + 1;                    // [1] leader: logits read by both epilogues
  uint64_t* bar_gfull = bar_sempty + 1;                    // [1] leader: both G buffers written
  uint64_t* bar_gempty = bar_gfull + 1;                    // [1] G consumed (multicast commit)
  uint64_t* bar_accfull = bar_gempty + 1;                  // [1] (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);      // [2][128]
  uint8_t* sm_astage = sm_v + (NV - 2) * Cfg::kVBytes;     // own rows arrive here (first used by tile NV-2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader; grid = (row blocks, segments, z),
                                                           // clusters of {2,1,1}: the pair must be adjacent in x
  const int64_t i0 = (int64_t)blockIdx.x * kTileRows;
  const int64_t jlo = 0, jhi = n_cols;                     // single bucket: every row block sweeps every column
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.y * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;     // the same in both CTAs of the pair
  float* acc_out = g.acc + (int64_t)blockIdx.y * n_rows * d;
  if (T == 0) {
    griddep_wait();
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DN; e += kNumThreads) {
      const int64_t rr = i0 + e / DN, col = e % DN;
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) TR(0);
  pdl_trigger();
  // Scalars of the epilogue (temperature, bias, this thread's row sum): requested at entry, first used after the
  // owned rows are parked -- the two dependent global round trips used to sit between the set-up and the parking
  // (~2000 cycles of the prologue; inputs of this kernel were written before the loss reduction let it start).
  float pre_ls = 0.f, pre_bias = 0.f, pre_rs = 1.f;
  if (warp >= 2) {
    pre_ls = *ls;
    if (SIG) pre_bias = *ga.bias;
    const int64_t ip = i0 + (warp & 3) * 32 + lane;
    if (!SIG && ip < n_rows) pre_rs = g.rs[ip];
  }

  constexpr int kEpiWarps = kEpiThreads / 32;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tb);
    tma_prefetch_desc(&g.tbp);
    mbar_init(bar_afull, 1);
    mbar_init(bar_aloc, kEpiWarps);
    mbar_init(bar_aall, 2 * kEpiWarps);
    for (int b = 0; b < NS; ++b) {
      mbar_init(bar_yfull + b, 1);
      mbar_init(bar_pfull + b, 1);
      mbar_init(bar_yempty + b, 1);
    }
    for (int b = 0; b < NV; ++b) {
      mbar_init(bar_vfull + b, 1);
      mbar_init(bar_pvfull + b, 1);
      mbar_init(bar_vempty + b, 1);
    }
    mbar_init(bar_sfull, 1);
    mbar_init(bar_sempty, 2 * kEpiWarps);
    mbar_init(bar_gfull, 2 * kEpiWarps);
    mbar_init(bar_gempty, 1);
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  // One tile's operands: this CTA's 64 logits columns for every K chunk, then every tile row for its half of d.
  auto load_s = [&](int t) {
    const int b = t % NS;
    const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
    uint8_t* buf = sm_s + b * Cfg::kSBytes;
    mbar_expect_tx(bar_yfull + b, Cfg::kSBytes);
    for (int c = 0; c < KD; ++c)
      tma_load_2d(buf + c * (kChunkBytes / 2), &g.tbp, bar_yfull + b, c * kChunkK, j0 + 64 * (int)rank);
    if (t < 16) TR(48 + t);
  };
  auto load_v = [&](int t) {
    const int b = t % NV;
    const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
    uint8_t* buf = sm_v + b * Cfg::kVBytes;
    mbar_expect_tx(bar_vfull + b, Cfg::kVBytes);
    for (int c = 0; c < KH; ++c)
      tma_load_2d(buf + c * kChunkBytes, &g.tb, bar_vfull + b, ((int)rank * KH + c) * kChunkK, j0);
  };
  // The owned rows and the first tiles land in this CTA's own shared memory and complete on its own barriers:
  // they are requested BEFORE the cluster rendezvous and the tensor-memory allocation (the operands were written
  // two kernels earlier -- the loss reduction between waits for the forward before it lets this grid start), so
  // ~1300 cycles of set-up run under the first loads' latency instead of in front of it.
  constexpr int kEarlyMax = NS < NV - 2 ? NS : NV - 2;    // slots that are free before anything has been consumed
  const int kEarly = T < kEarlyMax ? T : kEarlyMax;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_afull, KD * kChunkBytes);
    for (int c = 0; c < KD; ++c) tma_load_2d(sm_astage + c * kChunkBytes, &g.ta, bar_afull, c * kChunkK, (int)i0);
    for (int t = 0; t < kEarly; ++t) { load_s(t); load_v(t); }
  }
  if (warp == 1) tmem_alloc2<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_exec();      // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kSCol = 0, kACol = 128, kAccCol = 256;
  if (threadIdx.x == 0) TR(1);

  if (warp == 0) {
    if (lane == 0) {
      for (int t = kEarly; t < T; ++t) {   // use n of a slot waits for the slot's (n-1)-th release: parity (n-1) & 1
        mbar_wait(bar_yempty + t % NS, ((t / NS) & 1) ^ 1);
        load_s(t);
        if (t == NV - 2) mbar_wait(bar_aloc, 0);   // the staged rows have left the last two G.V slots
        mbar_wait(bar_vempty + t % NV, ((t / NV) & 1) ^ 1);
        load_v(t);
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank != 0) {
    // peer: tell the leader when this CTA's operands of tile t are in shared memory
    for (int t = 0; t < T; ++t) {
      mbar_wait(bar_yfull + t % NS, (t / NS) & 1);
      if (lane == 0) mbar_arrive_remote(bar_pfull + t % NS, 0);
      __syncwarp();
      mbar_wait(bar_vfull + t % NV, (t / NV) & 1);
      if (lane == 0) mbar_arrive_remote(bar_pvfull + t % NV, 0);
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_16(256, 128, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(256, DN, 0, 1, F16);   // A = G (K-major, smem), B = tile MN-major
    mbar_wait(bar_aall, 0);     // both CTAs' epilogue warps have parked their rows in TMEM
    tc_fence_after();
    if (lane == 0) TR(3);
    const uint32_t s_tmem = tmem_base + kSCol, a_tmem0 = tmem_base + kACol, acc_tmem = tmem_base + kAccCol;
    const uint32_t y_lo0 = umma_desc_lo(smem_u32(sm_s), 16);                              // K-major view (S)
    const uint32_t y2_lo0 = umma_desc_lo(smem_u32(sm_v), kChunkBytes);                    // MN-major view (G.V)
    const uint32_t g_lo0 = umma_desc_lo(smem_u32(sm_g), 16);
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        const int b = t % NS;
        const uint32_t ph = (t / NS) & 1;
        mbar_wait(bar_yfull + b, ph);
        mbar_wait(bar_pfull + b, ph);
        if (t >= 1) mbar_wait(bar_sempty, (t - 1) & 1);
        tc_fence_after();
        const uint32_t b_lo = y_lo0 + b * (Cfg::kSBytes >> 4);
#ifdef PLK_TRACE_PROBE
        const long long ts0 = clock64();
#endif
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < KD; ++c)
#pragma unroll
            for (int k = 0; k < kChunkK / kUmmaK; ++k)
              umma2_bf16_ts(s_tmem, a_tmem0 + c * 32 + k * 8, b_lo + c * (kChunkBytes >> 5) + 2 * k, idesc_s,
                            (c | k) != 0);
          umma2_commit(bar_sfull);
          umma2_commit(bar_yempty + b);   // the logits operands are free as soon as S(t) has retired
        }
        __syncwarp();
        if (lane == 0 && t < 16) TR(64 + t);
#ifdef PLK_TRACE_PROBE
        mbar_wait(bar_sfull, t & 1);     // probe: serialise, measure issue -> completion of the 16 S MMAs
        if (lane == 0 && t < 16) TRV(24 + (t & 7), clock64() - ts0);
#endif
      }
      if (t >= 1) {
        const int u = t - 1;
        const int bg = u % NV;
        mbar_wait(bar_vfull + bg, (u / NV) & 1);
        mbar_wait(bar_pvfull + bg, (u / NV) & 1);
        mbar_wait(bar_gfull, u & 1);
        tc_fence_after();
        const uint32_t b_lo = y2_lo0 + bg * (Cfg::kVBytes >> 4);
#ifdef PLK_TRACE_PROBE
        const long long tg0 = clock64();
#endif
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileRows / kUmmaK; ++k)
            umma2_bf16_lo(acc_tmem, g_lo0 + (k >> 2) * (kChunkBytes >> 4) + (k & 3) * 2, b_lo + k * (2048 >> 4),
                          idesc_g, (u | k) != 0);
          umma2_commit(bar_vempty + bg);
          umma2_commit(bar_gempty);
        }
        __syncwarp();
        if (lane == 0 && u < 16) TR(112 + u);
#ifdef PLK_TRACE_PROBE
        mbar_wait(bar_gempty, u & 1);    // probe: issue -> completion of the 8 G.V MMAs
        if (lane == 0 && u < 16) TRV(32 + (u & 7), clock64() - tg0);
#endif
      }
    }
    if (elect_one()) umma2_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cc = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    constexpr bool siglip = SIG;
    if (i < n_rows) bucket_range(gi, bs, n_cols, lo, hi);
    const bool want_gs = g.gs != nullptr;
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    {
      mbar_wait(bar_afull, 0);
      for (int c = cc; c < KD; c += 4) {
        const uint8_t* rowp = sm_astage + c * kChunkBytes + r * 128;
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const uint4 w = lds_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4));
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + lane_addr + kACol + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_aloc);
        mbar_arrive_remote(bar_aall, 0);
      }
    }
    if (i < n_rows && !siglip) rrs = 1.0f / pre_rs;
    const float s = expf(pre_ls);
    const float c1 = s * kLog2e, c0 = siglip ? pre_bias * kLog2e : (kShiftK - s) * kLog2e;
    // 1 / (column sum) of the tile's columns, double buffered.  The four warps that share a 32-column chunk all
    // write the SAME 32 values (identical stores) and each reads only behind its own write + __syncwarp, so no
    // CTA-wide barrier per tile is needed (it made every warp wait for the slowest one: 10 % of the epilogue's
    // stall samples).  A warp cannot overwrite a buffer a slower warp still reads: it stores after the gempty(t-1)
    // wait, i.e. after every warp has arrived on gfull(t-1), which follows that warp's last read of tile t-1.
    float rc_next = 0.f;
    if (!siglip) {
      const int64_t jc = jlo + (int64_t)t_begin * kTileRows + cc * 32 + lane;
      rcs_s[cc * 32 + lane] = (jc < n_cols) ? 1.0f / g.cs[jc] : 0.f;
      __syncwarp();
    }
    // (Tried: software-pipelining this loop by one tile -- tcgen05.ld of tile t+1 issued before the fence and
    // the gfull arrive of tile t.  It delays gfull(t) until S(t+1) has completed, which opens a ~900-cycle
    // bubble in the tensor pipe between S(t+1) and G.V(t): 2900 cycles per tile against 2600.)
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows;
      if (t + 1 < T && !siglip) {
        const int64_t jc = j0 + kTileRows + cc * 32 + lane;
        rc_next = (jc < n_cols) ? g.cs[jc] : 0.f;
      }
      mbar_wait(bar_sfull, t & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && t < 16) TR(80 + t);
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kSCol + cc * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar_sempty, 0);
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_s + buf * 128 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      if (t >= 1) mbar_wait(bar_gempty, (t - 1) & 1);
      const uint32_t grow = smem_u32(sm_g + (cc >> 1) * kChunkBytes + r * 128);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {   // explicit st.shared: a generic store goes through the global path (stall_lg)
        const int chunk = ((cc & 1) * 4 + c16) ^ (r & 7);
        sts_v4(grow + chunk * 16, packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar_gfull, 0);
      if (threadIdx.x == 64 && t < 16) TR(96 + t);
      if (t + 1 < T && !siglip) {
        rcs_s[(buf ^ 1) * 128 + cc * 32 + lane] = (rc_next != 0.f) ? 1.0f / rc_next : 0.f;
        __syncwarp();
      }
    }
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
    if (threadIdx.x == 64) TR(2);
    griddep_wait();
    if (g.use_tacc) {
#pragma unroll 1
      for (int ch = cc; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        uint8_t* stage = sm_y + ch * kChunkBytes;     // 128 rows x 128 B in the idle tile buffers
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
          sts_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4), raw[v4 * 4], raw[v4 * 4 + 1], raw[v4 * 4 + 2], raw[v4 * 4 + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(2 + cc, 128);
        if (q == 0 && lane == 0 && i0 < n_rows) {
          tma_store_3d(&g.tacc, stage, ch * 32, (int)i0, (int)blockIdx.y);
          tma_store_commit();
        }
      }
      if (q == 0 && lane == 0) tma_store_wait_read();
    } else {
#pragma unroll 1
      for (int ch = cc; ch < 2 * KD; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        const int64_t col0 = (int64_t)ch * 32;
        if (i < n_rows) {
          float* dst = acc_out + i * d + col0;
#pragma unroll
          for (int x = 0; x < 32; ++x)
            if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
        }
      }
    }
    if (want_gs) {
      gs_local *= s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
  }
  griddep_wait();
  tc_fence_before();
  if (threadIdx.x == 0) TR(5);
  cluster_sync_exec();     // neither CTA frees the pair's tensor memory (or exits) while the other still reads it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<512>(tmem_base);
  }
  if (threadIdx.x == 0) TR(6);
}

// =============================================================================================
// backward, 448 < d <= 512 (eight 64-element chunks): "tc8".
//
// The [128 x 512] fp32 accumulator alone is all of TMEM, so a CTA owns one 256-column half of d (blockIdx.z)
// and recomputes the logits for it, as infonce_grad_tc does.  What changes against that kernel (8300 cycles per
// 128 x 128 tile) follows from the measured operand rates (see infonce_grad_tc4):
//   * the first four K chunks of the owned rows are parked in TMEM (128 columns) and used in TS mode (64 cycles
//     per MMA), only the last four stay in shared memory (SS mode, 128 cycles): S costs 3072 cycles, not 4096;
//   * ONE logits buffer, handed back right after the epilogue's tcgen05.ld, makes room for them;
//   * G . b_t consumes the streamed chunks in PAIRS, as MN-major operands of N = 128 (adjacent ring slots,
//     LBO = one slot): 2 x 8 MMAs of 128 cycles instead of 4 x 8 of 96;
//   * G is double buffered in shared memory, so the epilogue of tile t+1 never waits for G . b_t.
//     TMEM  [0,128) S | [128,256) owned rows, K chunks 0..3 | [256,512) acc [128 x 256] fp32
//     smem  owned rows, K chunks 4..7 (64 KiB) | 2 G buffers (64 KiB; K chunks 0..3 are staged here first)
//           | ring of six 16 KiB chunks (cluster multicast) | 2 KiB barriers  = 227 KiB
// Ring order (producer and MMA agree): S(0), S(1), GV(0), S(2), GV(1), ..., GV(T-1), eight chunks per S,
// four per GV -- the slot index stays even at every pair.
// =============================================================================================
struct Grad8Cfg {
  static constexpr int KD = 8, KA = 4, DNC = 4, NST = 6;
  static constexpr int kAHi = (KD - KA) * kChunkBytes;
  static constexpr int kGBuf = 2 * kChunkBytes;
  static constexpr int kAux = 2048;
  static constexpr int kSmem = 1024 + kAHi + 2 * kGBuf + NST * kChunkBytes + kAux;
  static_assert(kSmem <= kMaxSmem, "tc8 shared memory");
  static_assert(KA * kChunkBytes <= 2 * kGBuf, "K chunks 0..3 of the owned rows are staged in the G buffers");
  static_assert(2 * kGBuf + NST * kChunkBytes >= kTileRows * DNC * 64 * 4, "accumulator drain staging");
};

template <int CS, bool F16>
__global__ void __launch_bounds__(kNumThreads, 1) infonce_grad_tc8(
    const __grid_constant__ GradArgs ga, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs,
    int tiles_per_seg, const float* __restrict__ ls) {
  const GradDir& g = ga.dir[blockIdx.z % ga.ndir];
  using Cfg = Grad8Cfg;
  constexpr int KD = Cfg::KD, KA = Cfg::KA, DNC = Cfg::DNC, NST = Cfg::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_a = smem;                                   // K chunks KA..KD-1 of the owned rows
  uint8_t* sm_g = smem + Cfg::kAHi;                       // [2] G buffers
  uint8_t* sm_ring = sm_g + 2 * Cfg::kGBuf;
  uint8_t* aux = sm_ring + NST * kChunkBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(aux);  // [NST]
  uint64_t* bar_empty = bar_full + NST;                   // [NST]
  uint64_t* bar_afull = bar_empty + NST;                  // [1] K chunks 0..3 landed in the G buffers
  uint64_t* bar_ahi = bar_afull + 1;                      // [1] K chunks 4..7 resident
  uint64_t* bar_a = bar_ahi + 1;                          // [1] K chunks 0..3 parked in TMEM
  uint64_t* bar_sfull = bar_a + 1;                        // [1]
  uint64_t* bar_sempty = bar_sfull + 1;                   // [1]
  uint64_t* bar_gfull = bar_sempty + 1;                   // [2]
  uint64_t* bar_accfull = bar_gfull + 2;                  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfull + 1);
  float* rcs_s = reinterpret_cast<float*>(aux + 512);     // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.y * kTileRows;
  const int h = blockIdx.z / ga.ndir;  // which 256-column half of d
  int64_t jlo = 0, jhi = n_cols;  // clusters are only launched for the single-bucket case
  if constexpr (CS == 1) row_block_cols(i0, n_rows, row_offset, bs, n_cols, jlo, jhi);
  const int total_tiles = (int)((jhi - jlo + kTileRows - 1) / kTileRows);
  const int t_begin = blockIdx.x * tiles_per_seg;
  int t_end = t_begin + tiles_per_seg;
  if (t_end > total_tiles) t_end = total_tiles;
  const int T = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0;
  float* acc_out = g.acc + (int64_t)blockIdx.x * n_rows * d;
  if (T == 0) {  // this segment has no tiles: its partial is zero (uniform across the CTA and the cluster)
    for (int64_t e = threadIdx.x; e < (int64_t)kTileRows * DNC * 64; e += kNumThreads) {
      const int64_t rr = i0 + e / (DNC * 64), col = (int64_t)h * DNC * 64 + e % (DNC * 64);
      if (rr < n_rows && col < d) acc_out[rr * d + col] = 0.f;
    }
    return;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.ta);
    tma_prefetch_desc(&g.tb);
    for (int s = 0; s < NST; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CS); }
    mbar_init(bar_afull, 1);
    mbar_init(bar_ahi, 1);
    mbar_init(bar_a, kEpiThreads);
    mbar_init(bar_sfull, 1);
    mbar_init(bar_sempty, kEpiThreads / 32);   // one elected arrival per epilogue warp
    mbar_init(bar_gfull, kEpiThreads / 32);
    mbar_init(bar_gfull + 1, kEpiThreads / 32);
    mbar_init(bar_accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();   // peers' barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kSCol = 0, kACol = 128, kAccCol = 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_afull, KA * kChunkBytes);
      for (int c = 0; c < KA; ++c) tma_load_2d(sm_g + c * kChunkBytes, &g.ta, bar_afull, c * kChunkK, (int)i0);
      mbar_expect_tx(bar_ahi, Cfg::kAHi);
      for (int c = KA; c < KD; ++c)
        tma_load_2d(sm_a + (c - KA) * kChunkBytes, &g.ta, bar_ahi, c * kChunkK, (int)i0);
      int st = 0; uint32_t ph = 0;
      auto push = [&](int col_chunk, int j0) {
        mbar_wait(bar_empty + st, ph ^ 1);
        ring_load<CS>(sm_ring + st * kChunkBytes, &g.tb, &g.tbp, bar_full + st, col_chunk * kChunkK, j0, cta_rank);
        if (++st == NST) { st = 0; ph ^= 1; }
      };
      for (int t = 0; t <= T; ++t) {
        if (t < T) {
          const int j0 = (int)(jlo + (int64_t)(t_begin + t) * kTileRows);
          for (int c = 0; c < KD; ++c) push(c, j0);
        }
        if (t >= 1) {
          const int j0 = (int)(jlo + (int64_t)(t_begin + t - 1) * kTileRows);
          for (int dc = 0; dc < DNC; ++dc) push(h * DNC + dc, j0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // All 32 lanes run this loop (warp-uniform control flow, see elect_one); one elected lane issues.
    constexpr uint32_t idesc_s = umma_idesc_16(128, 128, 0, 0, F16);
    constexpr uint32_t idesc_g = umma_idesc_16(128, 128, 0, 1, F16);  // B = two adjacent chunks, MN-major
    mbar_wait(bar_a, 0);
    mbar_wait(bar_ahi, 0);
    tc_fence_after();
    const uint32_t s_tmem = tmem_base + kSCol, a_tmem0 = tmem_base + kACol;
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(sm_a), 16), b_lo0 = umma_desc_lo(smem_u32(sm_ring), 16);
    const uint32_t g_lo0 = umma_desc_lo(smem_u32(sm_g), 16);
    const uint32_t b2_lo0 = umma_desc_lo(smem_u32(sm_ring), kChunkBytes);   // LBO = the next ring slot
    int st = 0; uint32_t ph = 0;
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        if (t >= 1) mbar_wait(bar_sempty, (t - 1) & 1);
#pragma unroll 1
        for (int c = 0; c < KD; ++c) {
          mbar_wait(bar_full + st, ph);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + st * (kChunkBytes >> 4);
          if (elect_one()) {
            if (c < KA) {
#pragma unroll
              for (int k = 0; k < kChunkK / kUmmaK; ++k)
                umma_bf16_ts(s_tmem, a_tmem0 + c * 32 + k * 8, b_lo + 2 * k, idesc_s, (c | k) != 0);
            } else {
              const uint32_t a_lo = a_lo0 + (c - KA) * (kChunkBytes >> 4);
#pragma unroll
              for (int k = 0; k < kChunkK / kUmmaK; ++k)
                umma_bf16_lo(s_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc_s, 1u);
            }
            ring_release<CS>(bar_empty + st);
            if (c == KD - 1) umma_commit(bar_sfull);
          }
          __syncwarp();
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
      if (t >= 1) {
        const int u = t - 1;
        mbar_wait(bar_gfull + (u & 1), (u >> 1) & 1);
        tc_fence_after();
        const uint32_t a_lo = g_lo0 + (u & 1) * (Cfg::kGBuf >> 4);
#pragma unroll 1
        for (int p = 0; p < DNC / 2; ++p) {
          mbar_wait(bar_full + st, ph);
          mbar_wait(bar_full + st + 1, ph);        // NST is even and st is even here: the pair never wraps
          tc_fence_after();
          const uint32_t b_lo = b2_lo0 + st * (kChunkBytes >> 4);
          const uint32_t d_tmem = tmem_base + kAccCol + p * 128;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTileRows / kUmmaK; ++k)
              // A = G[128 x 128] K-major: K 0..63 in sub-tile 0, 64..127 in sub-tile 1
              // B = two chunks [128 j x 64 cols] read MN-major as N = 128: 16 K-rows (j) = 2048 bytes per step
              umma_bf16_lo(d_tmem, a_lo + (k >> 2) * (kChunkBytes >> 4) + (k & 3) * 2, b_lo + k * (2048 >> 4), idesc_g,
                           (u | k) != 0);
            ring_release<CS>(bar_empty + st);
            ring_release<CS>(bar_empty + st + 1);
          }
          __syncwarp();
          st += 2;
          if (st == NST) { st = 0; ph ^= 1; }
        }
      }
    }
    if (elect_one()) umma_commit(bar_accfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cc = (warp - 2) >> 2;          // which 32-column chunk of the tile this warp owns
    const int r = q * 32 + lane;
    const int64_t i = i0 + r;
    const int64_t gi = row_offset + i;
    int64_t lo = 0, hi = 0;
    float rrs = 0.f;
    const bool siglip = ga.bias != nullptr;
    if (i < n_rows) {
      bucket_range(gi, bs, n_cols, lo, hi);
      if (!siglip) rrs = 1.0f / g.rs[i];
    }
    const float s = expf(*ls);
    const float c1 = s * kLog2e, c0 = siglip ? *ga.bias * kLog2e : (kShiftK - s) * kLog2e;
    const bool want_gs = (g.gs != nullptr) && (h == 0);
    float gs_local = 0.f, gsum_local = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    {  // park K chunks 0..3 of this thread's owned row in TMEM as packed pairs
      mbar_wait(bar_afull, 0);
      for (int c = cc; c < KA; c += 4) {
        const uint8_t* rowp = sm_g + c * kChunkBytes + r * 128;
        uint32_t pk[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const uint4 w = lds_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4));
          pk[v4 * 4 + 0] = w.x; pk[v4 * 4 + 1] = w.y; pk[v4 * 4 + 2] = w.z; pk[v4 * 4 + 3] = w.w;
        }
        tmem_st32(tmem_base + lane_addr + kACol + c * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a);
    }
    float rc_next = 0.f;
    if (cc == 0 && !siglip) {
      const int64_t jc = jlo + (int64_t)t_begin * kTileRows + r;
      rcs_s[r] = (jc < n_cols) ? 1.0f / g.cs[jc] : 0.f;
    }
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int64_t j0 = jlo + (int64_t)(t_begin + t) * kTileRows;
      named_barrier_sync(1, kEpiThreads);   // rcs_s[buf] visible; everyone is done with tile t-1
      if (cc == 0 && t + 1 < T && !siglip) {
        const int64_t jc = j0 + kTileRows + r;
        rc_next = (jc < n_cols) ? g.cs[jc] : 0.f;
      }
      mbar_wait(bar_sfull, t & 1);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + kSCol + cc * 32, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty);   // the logits buffer goes back to the MMA warp right away
      uint32_t packed[16];
      grad_chunk_dispatch<F16>(raw, packed, rcs_s + buf * 128 + cc * 32, rrs, c1, c0, lo, hi, gi, j0 + cc * 32,
                               want_gs, gs_local, siglip, gsum_local);
      // G buffer t&1 was last read by G.V(t-2), which the in-order tensor pipe ran before S(t) completed
      const uint32_t grow = smem_u32(sm_g + buf * Cfg::kGBuf + (cc >> 1) * kChunkBytes + r * 128);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {   // explicit st.shared: a generic store goes through the global path (stall_lg)
        const int chunk = ((cc & 1) * 4 + c16) ^ (r & 7);
        sts_v4(grow + chunk * 16, packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull + buf);
      if (cc == 0 && t + 1 < T) rcs_s[(buf ^ 1) * 128 + r] = (rc_next != 0.f) ? 1.0f / rc_next : 0.f;
    }
    // drain the resident accumulator: 2*DNC 32-column chunks over the four warps of a lane quadrant
    mbar_wait(bar_accfull, 0);
    tc_fence_after();
    if (g.use_tacc) {
#pragma unroll 1
      for (int ch = cc; ch < DNC * 2; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        uint8_t* stage = sm_g + ch * kChunkBytes;     // 128 rows x 128 B in the idle G buffers + ring
        uint8_t* rowp = stage + r * 128;
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
          sts_v4(smem_u32(rowp) + ((v4 ^ (r & 7)) << 4), raw[v4 * 4], raw[v4 * 4 + 1], raw[v4 * 4 + 2], raw[v4 * 4 + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(2 + cc, 128);
        if (q == 0 && lane == 0 && i0 < n_rows) {
          tma_store_3d(&g.tacc, stage, h * DNC * 64 + ch * 32, (int)i0, (int)blockIdx.x);
          tma_store_commit();
        }
      }
      if (q == 0 && lane == 0) tma_store_wait_read();
    } else {
#pragma unroll 1
      for (int ch = cc; ch < DNC * 2; ch += 4) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_addr + kAccCol + ch * 32, raw);
        tmem_ld_wait();
        const int64_t col0 = (int64_t)h * DNC * 64 + ch * 32;
        if (i < n_rows) {
          float* dst = acc_out + i * d + col0;
          if (col0 + 32 <= d && (d & 3) == 0) {
#pragma unroll
            for (int x = 0; x < 32; x += 4)
              *reinterpret_cast<float4*>(dst + x) =
                  make_float4(__uint_as_float(raw[x]), __uint_as_float(raw[x + 1]),
                              __uint_as_float(raw[x + 2]), __uint_as_float(raw[x + 3]));
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x)
              if (col0 + x < d) dst[x] = __uint_as_float(raw[x]);
          }
        }
      }
    }
    if (want_gs) {
      gs_local *= s;  // sum G * S with S = s * (u.v)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs_local += __shfl_xor_sync(0xffffffffu, gs_local, o);
        gsum_local += __shfl_xor_sync(0xffffffffu, gsum_local, o);
      }
      if (lane == 0) {
        atomicAdd(g.gs, gs_local);
        if (siglip) atomicAdd(g.gs + 1, gsum_local);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_exec();   // no CTA leaves while a peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =============================================================================================
// host launchers
// =============================================================================================
// widest column range [jlo, jhi) a 128-row block can need: the union of the buckets its rows touch
static int64_t max_col_span(int64_t bs, int64_t n_cols) {
  int64_t span = (ceil_div(kTileRows - 1, bs) + 1) * bs;
  return span > n_cols ? n_cols : span;
}

static int pick_segments(int64_t row_blocks, int64_t max_tiles, int z, int64_t cap = 16) {
  // Column segments per (row block, direction / d-half) work item.  A CTA costs its tiles plus a fixed
  // set-up + drain of about four tiles; the grid runs in waves of 148 CTAs (one per SM).  Pick the split
  // that minimises waves x (tiles per CTA + 4): it fills the SMs at small batches (C2: 32 row blocks ->
  // 4 segments forward, 2 x 2 backward) and trims the last partial wave at large ones (512 items: one
  // segment = 3.46 -> 4 waves of 260, two segments = 6.92 -> 7 waves of 132).
  const int64_t items = row_blocks * z;
  int64_t best = 1, best_cost = -1;
  // `cap`: the backward writes one fp32 partial slab per segment, which the gradient tail reads back
  const int64_t nmax = max_tiles < cap ? max_tiles : cap;
  for (int64_t n = 1; n <= nmax; ++n) {
    const int64_t waves = ceil_div(items * n, 148);
    const int64_t cost = waves * (ceil_div(max_tiles, n) + 4);
    if (best_cost < 0 || cost < best_cost) { best = n; best_cost = cost; }
  }
  return (int)best;
}

static int check_tc_shape(int64_t ld, int64_t d) {
  PLK_REQUIRE(ld % kChunkK == 0 && ld >= d && ld - d < kChunkK, PLK_ERR_INVALID,
              "16-bit operands must be zero-padded to ld = ceil(d/64)*64 (d=%lld ld=%lld)", (long long)d,
              (long long)ld);
  PLK_REQUIRE(ld <= 512, PLK_ERR_UNSUPPORTED, "tensor-core path supports d <= 512 (got %lld)", (long long)d);
  PLK_REQUIRE(plk_device_supports_tc(), PLK_ERR_ARCH, "the tensor-core path needs an sm_100 device");
  return PLK_OK;
}

// Clusters of 2 row blocks share every streamed chunk (multicast) when all row blocks sweep the
// same columns, i.e. in the single-bucket case.
static int pick_cluster(int64_t row_blocks, int64_t bs, int64_t n_cols) {
  return (bs >= n_cols && row_blocks >= 2) ? 2 : 1;
}

template <int KD, int CS>
static int launch_fwd(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbp, dim3 grid,
                      int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t bs, int tps,
                      const float* ls, float* rsum, float* csum, float* diag, int f16, cudaStream_t st,
                      const float* sig_bias, double* sig_sums) {
  auto kern = infonce_fwd_tc<KD, CS>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<KD>::kSmem));
    configured = true;
  }
  int rc = launch_kernel_ex(kern, grid, dim3(kNumThreads), FwdCfg<KD>::kSmem, st, CS, true, ta, tb, tbp, n_rows,
                         row_offset, n_cols, bs, tps, ls, rsum, csum, diag, f16, sig_bias, sig_sums);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}

static int fwd_tc16_impl(const void* u, const void* v, int f16, int64_t ld, int64_t n_rows,
                         int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* ls,
                         float* row_sumexp, float* col_sumexp, float* diag, int sums_zeroed, cudaStream_t st,
                         const float* sig_bias, double* sig_sums) {
  int rc = check_tc_shape(ld, d);
  if (rc) return rc;
  int64_t row_blocks = ceil_div(n_rows, kTileRows);
  const int cs = pick_cluster(row_blocks, bs, n_cols);
  PLK_REQUIRE(((uintptr_t)u & 15) == 0, PLK_ERR_INVALID, "operand must be 16-byte aligned");
  CUtensorMap ta, tb, tbp;
  if ((rc = make_tmap_bf16(&ta, u, n_rows, ld, ld, kTileRows))) return rc;
  if ((rc = make_tmap_bf16(&tb, v, n_cols, ld, ld, kTileRows))) return rc;
  if ((rc = make_tmap_bf16(&tbp, v, n_cols, ld, ld, kTileRows / 2))) return rc;
  // sums are accumulated with atomics -> zero first; diag needs no init (every owned row has its
  // diagonal column inside its bucket, so it is always written)
  if (!sums_zeroed && sig_bias == nullptr && (rc = zero2(row_sumexp, n_rows, col_sumexp, n_cols, st))) return rc;
  const int64_t max_tiles = ceil_div(max_col_span(bs, n_cols), kTileRows);
  const int nseg = pick_segments(row_blocks, max_tiles, 1);
  const int tps = (int)ceil_div(max_tiles, nseg);
  row_blocks = ceil_div(row_blocks, cs) * cs;   // phantom row block (all rows masked) pads the last cluster
  dim3 grid((unsigned)nseg, (unsigned)row_blocks, 1);
  switch (ld / kChunkK) {
#define PLK_CASE(KD)                                                                                         \
  case KD:                                                                                                   \
    return cs == 2 ? launch_fwd<KD, 2>(ta, tb, tbp, grid, n_rows, row_offset, n_cols, bs, tps, ls, row_sumexp, col_sumexp, diag, f16, st, sig_bias, sig_sums) \
                   : launch_fwd<KD, 1>(ta, tb, tbp, grid, n_rows, row_offset, n_cols, bs, tps, ls, row_sumexp, col_sumexp, diag, f16, st, sig_bias, sig_sums);
    PLK_CASE(1) PLK_CASE(2) PLK_CASE(3) PLK_CASE(4) PLK_CASE(5) PLK_CASE(6) PLK_CASE(7) PLK_CASE(8)
#undef PLK_CASE
  }
  set_error("unsupported padded width %lld", (long long)ld);
  return PLK_ERR_UNSUPPORTED;
}

int infonce_fwd_tc16(const void* u, const void* v, int f16, int64_t ld, int64_t n_rows,
                     int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* ls,
                     float* row_sumexp, float* col_sumexp, float* diag, int sums_zeroed, cudaStream_t st) {
  return fwd_tc16_impl(u, v, f16, ld, n_rows, row_offset, n_cols, d, bs, ls, row_sumexp, col_sumexp, diag,
                       sums_zeroed, st, nullptr, nullptr);
}

int siglip_fwd_tc16(const void* u, const void* v, int f16, int64_t ld, int64_t n_rows, int64_t row_offset,
                    int64_t n_cols, int64_t d, int64_t bs, const float* ls, const float* bias, float* diag,
                    double* sums, cudaStream_t st) {
  return fwd_tc16_impl(u, v, f16, ld, n_rows, row_offset, n_cols, d, bs, ls, nullptr, nullptr, diag, 1, st, bias,
                       sums);
}

template <int KD, int DNC, int CS, bool F16>
static int launch_grad(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                       int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st) {
  auto kern = infonce_grad_tc<KD, DNC, CS, F16>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GradCfg<KD, DNC>::kSmem));
    configured = true;
  }
  int rc = launch_kernel(kern, grid, dim3(kNumThreads), GradCfg<KD, DNC>::kSmem, st, CS, ga, n_rows, row_offset,
                         n_cols, d, bs, tps, ls);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}

template <int KD, int CS, bool F16, bool SIG>
static int launch_grad2_m(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                          int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  auto kern = infonce_grad_tc2<KD, CS, F16, SIG>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Grad2Cfg<KD>::kSmem));
    configured = true;
  }
  int rc = launch_kernel_ex(kern, grid, dim3(kNumThreads), Grad2Cfg<KD>::kSmem, st, CS, overlap_prev, ga, n_rows,
                            row_offset, n_cols, d, bs, tps, ls);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}
template <int KD, int CS, bool F16>
static int launch_grad2(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                        int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  return ga.bias != nullptr
             ? launch_grad2_m<KD, CS, F16, true>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev)
             : launch_grad2_m<KD, CS, F16, false>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
}

template <int KD, bool F16, bool SIG>
static int launch_grad3_m(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                          int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  auto kern = infonce_grad_tc3<KD, F16, SIG>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Grad3Cfg<KD>::kSmem));
    configured = true;
  }
  int rc = launch_kernel_ex(kern, grid, dim3(kNumThreads), Grad3Cfg<KD>::kSmem, st, 1, overlap_prev, ga, n_rows,
                            row_offset, n_cols, d, bs, tps, ls);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}
template <int KD, bool F16>
static int launch_grad3(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                        int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  return ga.bias != nullptr
             ? launch_grad3_m<KD, F16, true>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev)
             : launch_grad3_m<KD, F16, false>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
}

template <int KD, bool F16, bool SIG>
static int launch_grad4_m(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                          int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  auto kern = infonce_grad_tc4<KD, F16, SIG>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Grad4Cfg<KD>::kSmem));
    configured = true;
  }
  int rc = launch_kernel_ex(kern, grid, dim3(kNumThreads), Grad4Cfg<KD>::kSmem, st, 1, overlap_prev, ga, n_rows,
                            row_offset, n_cols, d, bs, tps, ls);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}
template <int KD, bool F16>
static int launch_grad4(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                        int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  return ga.bias != nullptr
             ? launch_grad4_m<KD, F16, true>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev)
             : launch_grad4_m<KD, F16, false>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
}

template <int KD, bool F16, bool SIG>
static int launch_grad5_m(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                          int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  auto kern = infonce_grad_tc5<KD, F16, SIG>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Grad5Cfg<KD>::kSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Grad5Cfg<KD>::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_prev ? 2 : 1;
  PLK_CUDA(cudaLaunchKernelEx(&cfg, kern, ga, n_rows, row_offset, n_cols, d, bs, tps, ls));
  PLK_LAUNCHED(1);
  return PLK_OK;
}
template <int KD, bool F16>
static int launch_grad5(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                        int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st, bool overlap_prev) {
  return ga.bias != nullptr
             ? launch_grad5_m<KD, F16, true>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev)
             : launch_grad5_m<KD, F16, false>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
}
// The paired-CTA backward (cta_group::2) is the default where it applies (padded d = 128 / 256, one bucket, at
// least two row blocks); PLK_GRAD_TC5=0 selects infonce_grad_tc4 there.
static bool use_grad_tc5() {
  static const bool v = getenv("PLK_GRAD_TC5") == nullptr || getenv("PLK_GRAD_TC5")[0] != '0';
  return v;
}

template <int CS, bool F16>
static int launch_grad8(const GradArgs& ga, dim3 grid, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                        int64_t d, int64_t bs, int tps, const float* ls, cudaStream_t st) {
  auto kern = infonce_grad_tc8<CS, F16>;
  static bool configured = false;
  if (!configured) {
    PLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Grad8Cfg::kSmem));
    configured = true;
  }
  int rc = launch_kernel(kern, grid, dim3(kNumThreads), Grad8Cfg::kSmem, st, CS, ga, n_rows, row_offset, n_cols, d, bs,
                         tps, ls);
  if (rc) return rc;
  PLK_LAUNCHED(1);
  return PLK_OK;
}
// PLK_GRAD_TC8=0: the previous streaming kernel (infonce_grad_tc<8,4>) for 448 < d <= 512
static bool use_grad_tc8() {
  static const bool v = getenv("PLK_GRAD_TC8") == nullptr || getenv("PLK_GRAD_TC8")[0] != '0';
  return v;
}

// PLK_GRAD_TC2=1 selects the previous d <= 256 backward (owned rows read from shared memory, 128-column
// tiles) for A/B measurements; the default is infonce_grad_tc3.
static bool use_grad_tc2() {
  static const bool v = getenv("PLK_GRAD_TC2") != nullptr && getenv("PLK_GRAD_TC2")[0] == '1';
  return v;
}
// PLK_GRAD_TC3=1: infonce_grad_tc3 (64-column tiles, G in TMEM); default for d <= 256: infonce_grad_tc4
static bool use_grad_tc3() {
  static const bool v = getenv("PLK_GRAD_TC3") != nullptr && getenv("PLK_GRAD_TC3")[0] == '1';
  return v;
}
// logits columns per tile of the backward that serves a padded width of kd 64-element chunks
static int grad_tile_cols(int kd) { return (kd <= 4 && use_grad_tc3() && !use_grad_tc2()) ? kG3Cols : kTileRows; }

// number of partial accumulators per direction for this shape (bf16 path); ndir = 1 or 2 directions per launch
int grad_parts_tc16(int64_t n_rows, int64_t n_cols, int64_t d, int64_t bs, int ndir) {
  const int64_t ld = ceil_div(d, kChunkK) * kChunkK;
  const int z = (ld > 256 ? 2 : 1) * ndir;
  const int64_t row_blocks = ceil_div(n_rows, kTileRows);
  const int64_t max_tiles = ceil_div(max_col_span(bs, n_cols), grad_tile_cols((int)(ld / kChunkK)));
  const int64_t nseg = pick_segments(row_blocks, max_tiles, z, 8);
  return (int)ceil_div(max_tiles, ceil_div(max_tiles, nseg));   // no empty trailing segment
}

static int fill_dir(GradDir& g, const void* a, const void* b, int64_t ld, int64_t n_rows,
                    int64_t n_cols, const float* rs, const float* cs, float* acc, float* gs) {
  int rc;
  if ((rc = make_tmap_bf16(&g.ta, a, n_rows, ld, ld, kTileRows))) return rc;
  if ((rc = make_tmap_bf16(&g.tb, b, n_cols, ld, ld, kTileRows))) return rc;
  if ((rc = make_tmap_bf16(&g.tbp, b, n_cols, ld, ld, kTileRows / 2))) return rc;
  g.rs = rs; g.cs = cs; g.acc = acc; g.gs = gs;
  g.use_tacc = 0;
  return PLK_OK;
}

template <bool F16>
static int grad_launch_16(const GradArgs& ga_in, int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols,
                          int64_t d, int64_t bs, const float* ls, cudaStream_t st, bool overlap_prev = false) {
  GradArgs ga = ga_in;
  int64_t row_blocks = ceil_div(n_rows, kTileRows);
  const int csz = pick_cluster(row_blocks, bs, n_cols);
  const int kd = (int)(ld / kChunkK);
  const int z = (kd > 4 ? 2 : 1) * ga.ndir;
  const int64_t max_tiles = ceil_div(max_col_span(bs, n_cols), grad_tile_cols(kd));
  int nseg = pick_segments(row_blocks, max_tiles, z, 8);
  const int tps = (int)ceil_div(max_tiles, nseg);
  nseg = (int)ceil_div(max_tiles, tps);   // no empty trailing segment (grad_parts_tc16 reports the same count)
  if (kd <= 4) {   // d <= 256: the whole [128 x d] accumulator is TMEM-resident, G never leaves tensor memory
    if (d % 32 == 0) {   // accumulator drain by TMA store (full 128-byte lines)
      for (int k = 0; k < ga.ndir; ++k) {
        if (((uintptr_t)ga.dir[k].acc & 15) != 0) continue;
        int rc = make_tmap_f32_slabs(&ga.dir[k].tacc, ga.dir[k].acc, nseg, n_rows, d);
        if (rc) return rc;
        ga.dir[k].use_tacc = 1;
      }
      if (ga.ndir == 1) ga.dir[1] = ga.dir[0];
    }
    if (!use_grad_tc2() && !use_grad_tc3() && use_grad_tc5() && (kd == 2 || kd == 4) && csz == 2 &&
        !ga.tail.enabled) {
      dim3 grid5((unsigned)(ceil_div(row_blocks, 2) * 2), (unsigned)nseg, (unsigned)z);
      return kd == 4 ? launch_grad5<4, F16>(ga, grid5, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev)
                     : launch_grad5<2, F16>(ga, grid5, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
    }
    if (!use_grad_tc2() && !use_grad_tc3()) {   // S in TS mode at N = 128, G through shared memory, no clusters
      dim3 grid4((unsigned)nseg, (unsigned)row_blocks, (unsigned)z);
      switch (kd) {
#define PLK_CASE4(KD) \
  case KD: return launch_grad4<KD, F16>(ga, grid4, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
        PLK_CASE4(1) PLK_CASE4(2) PLK_CASE4(3) PLK_CASE4(4)
#undef PLK_CASE4
      }
    }
    if (!use_grad_tc2()) {   // both GEMMs in TS mode, 64-column tiles, no clusters
      dim3 grid3((unsigned)nseg, (unsigned)row_blocks, (unsigned)z);
      switch (kd) {
#define PLK_CASE3(KD) \
  case KD: return launch_grad3<KD, F16>(ga, grid3, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
        PLK_CASE3(1) PLK_CASE3(2) PLK_CASE3(3) PLK_CASE3(4)
#undef PLK_CASE3
      }
    }
  }
  row_blocks = ceil_div(row_blocks, csz) * csz;
  dim3 grid((unsigned)nseg, (unsigned)row_blocks, (unsigned)z);
  if (kd <= 4) {
    switch (kd) {
#define PLK_CASE2(KD)                                                                                   \
  case KD:                                                                                              \
    return csz == 2 ? launch_grad2<KD, 2, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev) \
                    : launch_grad2<KD, 1, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st, overlap_prev);
      PLK_CASE2(1) PLK_CASE2(2) PLK_CASE2(3) PLK_CASE2(4)
#undef PLK_CASE2
    }
  }
  if (kd == 8 && use_grad_tc8()) {
    if (d % 32 == 0) {   // accumulator drain by TMA store (full 128-byte lines)
      for (int k = 0; k < ga.ndir; ++k) {
        if (((uintptr_t)ga.dir[k].acc & 15) != 0) continue;
        int rc = make_tmap_f32_slabs(&ga.dir[k].tacc, ga.dir[k].acc, nseg, n_rows, d);
        if (rc) return rc;
        ga.dir[k].use_tacc = 1;
      }
      if (ga.ndir == 1) ga.dir[1] = ga.dir[0];
    }
    return csz == 2 ? launch_grad8<2, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st)
                    : launch_grad8<1, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st);
  }
  switch (kd) {
#define PLK_CASE(KD, DNC)                                                                               \
  case KD:                                                                                              \
    return csz == 2 ? launch_grad<KD, DNC, 2, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st) \
                    : launch_grad<KD, DNC, 1, F16>(ga, grid, n_rows, row_offset, n_cols, d, bs, tps, ls, st);
    PLK_CASE(5, 3) PLK_CASE(6, 3) PLK_CASE(7, 4) PLK_CASE(8, 4)
#undef PLK_CASE
  }
  set_error("unsupported padded width %lld", (long long)ld);
  return PLK_ERR_UNSUPPORTED;
}

int infonce_grad_tc16(const void* a, const void* b, int f16, int64_t ld, int64_t n_rows,
                      int64_t row_offset, int64_t n_cols, int64_t d, int64_t bs, const float* ls,
                      const float* rs, const float* cs, float* acc, float* gs, cudaStream_t st) {
  int rc = check_tc_shape(ld, d);
  if (rc) return rc;
  GradArgs ga;
  ga.ndir = 1;
  ga.bias = nullptr;
  ga.tail = GradTail{};
  if ((rc = fill_dir(ga.dir[0], a, b, ld, n_rows, n_cols, rs, cs, acc, gs))) return rc;
  ga.dir[1] = ga.dir[0];
  return f16 ? grad_launch_16<true>(ga, ld, n_rows, row_offset, n_cols, d, bs, ls, st)
             : grad_launch_16<false>(ga, ld, n_rows, row_offset, n_cols, d, bs, ls, st);
}

// Can the gradient tail run inside the backward kernel?  (InfoNCE, one bucket, d <= 256 with d % 128 == 0, fp32
// rows that the vectorised tail accepts.)  OFF unless PLK_FUSE_TAIL=1: measured at B = 4096, d = 256 the fused
// step takes 73.0 us against 69.6 us with the separate tail kernel -- the tail is latency-bound row work, and
// inside the backward only the 64 last-arriving CTAs (1024 warps, eight rows each, one after the other) do it,
// on the kernel's critical path, where the separate launch spreads 8192 rows over every SM at once.
bool grad_tail_fusable(int64_t n_rows, int64_t n_cols, int64_t d, int64_t bs, int64_t ldx, const void* x, const void* y,
                       const void* dx, const void* dy, const void* acc) {
  static const bool on = getenv("PLK_FUSE_TAIL") != nullptr && getenv("PLK_FUSE_TAIL")[0] == '1';
  if (!on || use_grad_tc2() || use_grad_tc3()) return false;
  if (bs < n_cols || n_rows != n_cols || d % 128 != 0 || d > 256 || (ldx & 3)) return false;
  return ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)dx | (uintptr_t)dy | (uintptr_t)acc) & 15) == 0);
}

int infonce_grad_pair_tc16_tail(const void* a0, const void* b0, const void* a1, const void* b1, int f16, int64_t ld,
                                int64_t n_rows, int64_t d, const float* ls, const float* rs0, const float* cs0,
                                const float* rs1, const float* cs1, float* acc0, float* acc1, float* gs, cudaStream_t st,
                                int overlap_prev, const GradTailHost& th) {
  int rc = check_tc_shape(ld, d);
  if (rc) return rc;
  GradArgs ga;
  ga.ndir = 2;
  ga.bias = nullptr;
  if ((rc = fill_dir(ga.dir[0], a0, b0, ld, n_rows, n_rows, rs0, cs0, acc0, gs))) return rc;
  if ((rc = fill_dir(ga.dir[1], a1, b1, ld, n_rows, n_rows, rs1, cs1, acc1, nullptr))) return rc;
  GradTail& t = ga.tail;
  t.enabled = 1;
  t.nseg = grad_parts_tc16(n_rows, n_rows, d, n_rows, 2);
  t.ldx = th.ldx; t.batch = th.batch;
  t.x[0] = th.x; t.x[1] = th.y;
  t.inv_den[0] = th.inv_den_x; t.inv_den[1] = th.inv_den_y;
  t.nrm[0] = th.nrm_x; t.nrm[1] = th.nrm_y;
  t.dx[0] = th.dx; t.dx[1] = th.dy;
  t.diag = th.diag; t.R = rs0; t.C = cs0;
  t.grad_out_emb = th.grad_out_emb; t.grad_out = th.grad_out; t.emb_scale = th.emb_scale;
  t.diag_sum = th.diag_sum; t.dls_out = th.dls_out; t.counters = th.counters;
  return f16 ? grad_launch_16<true>(ga, ld, n_rows, 0, n_rows, d, n_rows, ls, st, overlap_prev != 0)
             : grad_launch_16<false>(ga, ld, n_rows, 0, n_rows, d, n_rows, ls, st, overlap_prev != 0);
}

int infonce_grad_pair_tc16(const void* a0, const void* b0, const void* a1,
                           const void* b1, int f16, int64_t ld, int64_t n_rows, int64_t row_offset,
                           int64_t n_cols, int64_t d, int64_t bs, const float* ls, const float* rs0,
                           const float* cs0, const float* rs1, const float* cs1, float* acc0, float* acc1,
                           float* gs, cudaStream_t st, int overlap_prev, const float* bias) {
  int rc = check_tc_shape(ld, d);
  if (rc) return rc;
  GradArgs ga;
  ga.ndir = 2;
  ga.bias = bias;
  ga.tail = GradTail{};
  if ((rc = fill_dir(ga.dir[0], a0, b0, ld, n_rows, n_cols, rs0, cs0, acc0, gs))) return rc;
  if ((rc = fill_dir(ga.dir[1], a1, b1, ld, n_rows, n_cols, rs1, cs1, acc1, nullptr))) return rc;
  return f16 ? grad_launch_16<true>(ga, ld, n_rows, row_offset, n_cols, d, bs, ls, st, overlap_prev != 0)
             : grad_launch_16<false>(ga, ld, n_rows, row_offset, n_cols, d, bs, ls, st, overlap_prev != 0);
}

}  // namespace plk

#ifdef PLK_TRACE
extern "C" int plk_debug_set_trace(long long* buf) {
  return cudaMemcpyToSymbol(plk::g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif
