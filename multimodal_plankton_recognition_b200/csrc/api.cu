// extern "C" surface of libplk.so: argument validation + dispatch on the operand dtype.
// Declarations and the reference lines each entry point replaces: include/plk.h.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace plk {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace plk

using namespace plk;

extern "C" {

int plk_version(void) { return 100; }
const char* plk_last_error(void) { return g_err; }
int64_t plk_launch_count(void) { return g_launches.load(); }

int plk_device_supports_tc(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

static int check_common(const void* a, const void* b, int op_dtype, int64_t ld, int64_t n_rows,
                        int64_t n_cols, int64_t d, int64_t bs) {
  PLK_REQUIRE(a && b, PLK_ERR_INVALID, "null operand pointer");
  PLK_REQUIRE(op_dtype == PLK_F32 || op_dtype == PLK_BF16 || op_dtype == PLK_F16, PLK_ERR_INVALID,
              "op_dtype must be PLK_F32, PLK_BF16 or PLK_F16 (got %d)", op_dtype);
  PLK_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0 && ld >= d, PLK_ERR_INVALID,
              "bad shape n_rows=%lld n_cols=%lld d=%lld ld=%lld", (long long)n_rows,
              (long long)n_cols, (long long)d, (long long)ld);
  PLK_REQUIRE(bs > 0, PLK_ERR_INVALID, "bucket_size must be positive");
  return PLK_OK;
}

int plk_infonce_fwd(const void* u, const void* v, int op_dtype, int64_t ld, int64_t n_rows,
                    int64_t row_offset, int64_t n_cols, int64_t d, int64_t bucket_size,
                    const float* logit_scale, float* row_sumexp, float* col_sumexp, float* diag,
                    int sums_zeroed, void* stream) {
  int rc = check_common(u, v, op_dtype, ld, n_rows, n_cols, d, bucket_size);
  if (rc) return rc;
  PLK_REQUIRE(logit_scale && row_sumexp && col_sumexp && diag, PLK_ERR_INVALID, "null output pointer");
  PLK_REQUIRE(row_offset >= 0 && row_offset + n_rows <= n_cols, PLK_ERR_INVALID,
              "owned rows [%lld,%lld) outside the global batch %lld", (long long)row_offset,
              (long long)(row_offset + n_rows), (long long)n_cols);
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == PLK_F32)
    return infonce_fwd_f32((const float*)u, (const float*)v, ld, n_rows, row_offset, n_cols, d,
                           bucket_size, logit_scale, row_sumexp, col_sumexp, diag, sums_zeroed, st);
  return infonce_fwd_tc16(u, v, op_dtype == PLK_F16, ld, n_rows, row_offset, n_cols, d, bucket_size,
                          logit_scale, row_sumexp, col_sumexp, diag, sums_zeroed, st);
}

int plk_infonce_grad_parts(int op_dtype, int64_t n_rows, int64_t n_cols, int64_t d,
                           int64_t bucket_size) {
  if (op_dtype == PLK_F32 || n_rows <= 0 || n_cols <= 0 || d <= 0 || bucket_size <= 0) return 1;
  return grad_parts_tc16(n_rows, n_cols, d, bucket_size, 1);
}

int plk_infonce_grad_pair_parts(int op_dtype, int64_t n_rows, int64_t n_cols, int64_t d,
                                int64_t bucket_size) {
  if (op_dtype == PLK_F32 || n_rows <= 0 || n_cols <= 0 || d <= 0 || bucket_size <= 0) return 1;
  return grad_parts_tc16(n_rows, n_cols, d, bucket_size, 2);
}

int plk_infonce_grad(const void* a, const void* b, int op_dtype, int64_t ld, int64_t n_rows,
                     int64_t row_offset, int64_t n_cols, int64_t d, int64_t bucket_size,
                     const float* logit_scale, const float* rs, const float* cs, float* acc,
                     float* gs, void* stream) {
  int rc = check_common(a, b, op_dtype, ld, n_rows, n_cols, d, bucket_size);
  if (rc) return rc;
  PLK_REQUIRE(logit_scale && rs && cs && acc, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(row_offset >= 0 && row_offset + n_rows <= n_cols, PLK_ERR_INVALID,
              "owned rows outside the global batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == PLK_F32)
    return infonce_grad_f32((const float*)a, (const float*)b, ld, n_rows, row_offset, n_cols, d,
                            bucket_size, logit_scale, rs, cs, acc, gs, st);
  return infonce_grad_tc16(a, b, op_dtype == PLK_F16, ld, n_rows, row_offset, n_cols, d, bucket_size,
                           logit_scale, rs, cs, acc, gs, st);
}

int plk_infonce_grad_pair(const void* a0, const void* b0, const void* a1, const void* b1, int op_dtype,
                          int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d,
                          int64_t bucket_size, const float* logit_scale, const float* rs0,
                          const float* cs0, const float* rs1, const float* cs1, float* acc0, float* acc1,
                          float* gs, void* stream) {
  int rc = check_common(a0, b0, op_dtype, ld, n_rows, n_cols, d, bucket_size);
  if (rc) return rc;
  PLK_REQUIRE(a1 && b1 && logit_scale && rs0 && cs0 && rs1 && cs1 && acc0 && acc1, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(row_offset >= 0 && row_offset + n_rows <= n_cols, PLK_ERR_INVALID,
              "owned rows outside the global batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == PLK_F32) {
    rc = infonce_grad_f32((const float*)a0, (const float*)b0, ld, n_rows, row_offset, n_cols, d, bucket_size,
                          logit_scale, rs0, cs0, acc0, gs, st);
    if (rc) return rc;
    return infonce_grad_f32((const float*)a1, (const float*)b1, ld, n_rows, row_offset, n_cols, d, bucket_size,
                            logit_scale, rs1, cs1, acc1, nullptr, st);
  }
  return infonce_grad_pair_tc16(a0, b0, a1, b1, op_dtype == PLK_F16, ld, n_rows, row_offset, n_cols, d,
                                bucket_size, logit_scale, rs0, cs0, rs1, cs1, acc0, acc1, gs, st);
}

// ---- the whole single-GPU loss step: layout of the saved state + the two composite calls ----
namespace {
struct ClipState {
  size_t u, v, stats, aux, cnt, bytes;
  int64_t ld, n_cnt;
  ClipState(int op_dtype, int64_t B, int64_t d) {
    const size_t esz = op_dtype == PLK_F32 ? 4 : 2;
    ld = op_dtype == PLK_F32 ? d : (d + 63) / 64 * 64;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    u = 0;
    v = up((size_t)B * ld * esz);
    stats = v + up((size_t)B * ld * esz);
    aux = stats + up((size_t)7 * B * 4);
    cnt = aux + 256;                               // counters of the fused gradient tail (zero between calls)
    n_cnt = 2 * ((B + 127) / 128) + 1;
    bytes = cnt + up((size_t)n_cnt * 4);
  }
};
int clip_parts(int op_dtype, int64_t B, int64_t d, int64_t bs) {
  return plk_infonce_grad_pair_parts(op_dtype, B, B, d, bs);
}
}  // namespace

size_t plk_clip_loss_state_bytes(int op_dtype, int64_t batch, int64_t d) {
  if (batch <= 0 || d <= 0) return 0;
  return ClipState(op_dtype, batch, d).bytes;
}

size_t plk_clip_loss_workspace_bytes(int op_dtype, int64_t batch, int64_t d, int64_t bucket_size) {
  if (batch <= 0 || d <= 0 || bucket_size <= 0) return 0;
  return (size_t)2 * clip_parts(op_dtype, batch, d, bucket_size) * batch * d * sizeof(float);
}

static int clip_forward_impl(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype,
                             int64_t bucket_size, int64_t batch_global, const float* logit_scale, void* state,
                             float* loss_out, float* partial_out, void* const* peer_bufs, int rank, int world,
                             unsigned* epoch, float* out2, void* stream) {
  PLK_REQUIRE(x && y && logit_scale && state && loss_out, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype >= PLK_F32 && op_dtype <= PLK_F16, PLK_ERR_INVALID, "bad op_dtype %d", op_dtype);
  PLK_REQUIRE(batch > 0 && d > 0 && ldx >= d && bucket_size > 0 && batch % bucket_size == 0 && batch_global >= batch,
              PLK_ERR_INVALID, "bad shape batch=%lld d=%lld ldx=%lld bucket_size=%lld batch_global=%lld",
              (long long)batch, (long long)d, (long long)ldx, (long long)bucket_size, (long long)batch_global);
  PLK_REQUIRE(((uintptr_t)state & 255) == 0, PLK_ERR_INVALID, "state must be 256-byte aligned");
  const ClipState L(op_dtype, batch, d);
  char* base = (char*)state;
  float* st = (float*)(base + L.stats);     // rows: 1/den_x, |x|, 1/den_y, |y|, row sum-exp, col sum-exp, diag
  float* aux = (float*)(base + L.aux);      // (sum of diagonal logits, gs accumulator)
  const int64_t B = batch;
  // one launch also zero-fills the two sum-exp rows (adjacent: 2B floats) and the fused tail's counters
  int rc = plk_l2norm_pair_fwd(x, y, B, d, ldx, base + L.u, base + L.v, op_dtype, L.ld, st, st + B, st + 2 * B,
                               st + 3 * B, st + 4 * B, 2 * B, (float*)(base + L.cnt), L.n_cnt, stream);
  if (rc) return rc;
  rc = plk_infonce_fwd(base + L.u, base + L.v, op_dtype, L.ld, B, 0, B, d, bucket_size, logit_scale, st + 4 * B,
                       st + 5 * B, st + 6 * B, 1, stream);
  if (rc) return rc;
  if (peer_bufs != nullptr)
    return plk_infonce_loss_xgpu(st + 4 * B, st + 5 * B, st + 6 * B, logit_scale, B, batch_global, loss_out, aux,
                                 aux + 1, partial_out, peer_bufs, rank, world, epoch, out2, stream);
  return plk_infonce_loss(st + 4 * B, st + 5 * B, st + 6 * B, logit_scale, B, batch_global, loss_out, aux, aux + 1,
                          stream);
}

int plk_clip_loss_forward(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx,
                          int op_dtype, int64_t bucket_size, int64_t batch_global,
                          const float* logit_scale, void* state, float* loss_out, void* stream) {
  return clip_forward_impl(x, y, batch, d, ldx, op_dtype, bucket_size, batch_global, logit_scale, state, loss_out,
                           nullptr, nullptr, 0, 1, nullptr, nullptr, stream);
}

int plk_clip_loss_forward_xgpu(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx,
                               int op_dtype, int64_t bucket_size, int64_t batch_global,
                               const float* logit_scale, void* state, float* loss_out, float* partial_out,
                               void* const* peer_bufs, int rank, int world, unsigned* epoch, float* out2,
                               void* stream) {
  PLK_REQUIRE(partial_out && peer_bufs && epoch && out2, PLK_ERR_INVALID, "null pointer");
  return clip_forward_impl(x, y, batch, d, ldx, op_dtype, bucket_size, batch_global, logit_scale, state, loss_out,
                           partial_out, peer_bufs, rank, world, epoch, out2, stream);
}

static int clip_backward_impl(const float* grad_out, const float* grad_out_emb, float emb_scale, const float* x, const float* y,
                              int64_t batch, int64_t d, int64_t ldx, int op_dtype, int64_t bucket_size,
                              int64_t batch_global, const float* logit_scale, void* state, void* workspace,
                              float* dx, float* dy, float* dls, const float* loss_partial, void* const* peer_bufs,
                              int rank, int world, unsigned* epoch, float* out2, void* stream) {
  PLK_REQUIRE(grad_out && grad_out_emb && x && y && logit_scale && state && workspace && dx && dy && dls,
              PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype >= PLK_F32 && op_dtype <= PLK_F16, PLK_ERR_INVALID, "bad op_dtype %d", op_dtype);
  PLK_REQUIRE(batch > 0 && d > 0 && ldx >= d && bucket_size > 0 && batch % bucket_size == 0 && batch_global >= batch,
              PLK_ERR_INVALID, "bad shape");
  const ClipState L(op_dtype, batch, d);
  char* base = (char*)state;
  float* st = (float*)(base + L.stats);
  float* aux = (float*)(base + L.aux);
  const int64_t B = batch;
  const int parts = clip_parts(op_dtype, B, d, bucket_size);
  float* acc_x = (float*)workspace;
  float* acc_y = acc_x + (size_t)parts * B * d;
  const float *rs = st + 4 * B, *cs = st + 5 * B;
  int rc;
  static const bool overlap = getenv("PLK_PDL") == nullptr || getenv("PLK_PDL")[0] != '0';
  if (op_dtype != PLK_F32 && peer_bufs == nullptr && batch_global == batch &&
      grad_tail_fusable(B, B, d, bucket_size, ldx, x, y, dx, dy, workspace)) {
    // d <= 256, one bucket: the gradient tail runs inside the recompute kernel (the last column segment of a
    // row block adds the partial slabs and finishes its 128 rows; the last tail produces d logit_scale) --
    // no second launch, and the slabs are read back while they are still being produced elsewhere
    GradTailHost th;
    th.x = x; th.y = y; th.ldx = ldx; th.batch = batch_global;
    th.inv_den_x = st; th.nrm_x = st + B; th.inv_den_y = st + 2 * B; th.nrm_y = st + 3 * B; th.diag = st + 6 * B;
    th.grad_out_emb = grad_out_emb; th.grad_out = grad_out; th.emb_scale = emb_scale;
    th.diag_sum = aux; th.dx = dx; th.dy = dy; th.dls_out = dls; th.counters = (int*)(base + L.cnt);
    return infonce_grad_pair_tc16_tail(base + L.u, base + L.v, base + L.v, base + L.u, op_dtype == PLK_F16, L.ld, B, d,
                                       logit_scale, rs, cs, cs, rs, acc_x, acc_y, aux + 1, (cudaStream_t)stream,
                                       overlap ? 1 : 0, th);
  }
  if (op_dtype != PLK_F32) {
    // Everything this launch reads was produced by the forward call, at least one kernel back in the
    // stream; only the sum G*S accumulator is zeroed by the forward's last kernel.  The grid may
    // therefore start under that kernel (or whatever ran in between) -- programmatic serialization.
    rc = infonce_grad_pair_tc16(base + L.u, base + L.v, base + L.v, base + L.u, op_dtype == PLK_F16, L.ld, B, 0, B,
                                d, bucket_size, logit_scale, rs, cs, cs, rs, acc_x, acc_y, aux + 1,
                                (cudaStream_t)stream, overlap ? 1 : 0);
  } else {
    rc = plk_infonce_grad_pair(base + L.u, base + L.v, base + L.v, base + L.u, op_dtype, L.ld, B, 0, B, d,
                               bucket_size, logit_scale, rs, cs, cs, rs, acc_x, acc_y, aux + 1, stream);
  }
  if (rc) return rc;
  return finish_pair_scaled(acc_x, acc_y, parts, x, y, B, d, ldx, st, st + B, st + 2 * B, st + 3 * B, st + 6 * B, rs,
                            cs, logit_scale, grad_out_emb, emb_scale, grad_out, batch_global, aux + 1, aux, dx, dy,
                            dls, loss_partial, peer_bufs, rank, world, epoch, out2, stream);
}

int plk_clip_loss_backward(const float* grad_out, const float* grad_out_emb, float emb_scale, const float* x,
                           const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype,
                           int64_t bucket_size, int64_t batch_global, const float* logit_scale, void* state,
                           void* workspace, float* dx, float* dy, float* dls, void* stream) {
  return clip_backward_impl(grad_out, grad_out_emb, emb_scale, x, y, batch, d, ldx, op_dtype, bucket_size, batch_global,
                            logit_scale, state, workspace, dx, dy, dls, nullptr, nullptr, 0, 1, nullptr, nullptr,
                            stream);
}

int plk_clip_loss_backward_xgpu(const float* grad_out, const float* grad_out_emb, float emb_scale, const float* x,
                                const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype, int64_t bucket_size,
                                int64_t batch_global, const float* logit_scale, void* state, void* workspace,
                                float* dx, float* dy, float* dls, const float* loss_partial,
                                void* const* peer_bufs, int rank, int world, unsigned* epoch, float* out2,
                                void* stream) {
  PLK_REQUIRE(loss_partial && peer_bufs && epoch && out2, PLK_ERR_INVALID, "null pointer");
  return clip_backward_impl(grad_out, grad_out_emb, emb_scale, x, y, batch, d, ldx, op_dtype, bucket_size, batch_global,
                            logit_scale, state, workspace, dx, dy, dls, loss_partial, peer_bufs, rank, world, epoch,
                            out2, stream);
}

// ---- N2: SigLIP loss (reference src/coordination.py:67-95) on the same state layout ----
//   aux region: double sums[3] (loss terms, sum_i G_ii S_ii, sum_i G_ii) at +0, float gs2[2] at +32
int plk_siglip_loss_forward(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype,
                            int64_t bucket_size, const float* logit_scale, const float* bias, void* state,
                            float* loss_out, void* stream) {
  PLK_REQUIRE(x && y && logit_scale && bias && state && loss_out, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype >= PLK_F32 && op_dtype <= PLK_F16, PLK_ERR_INVALID, "bad op_dtype %d", op_dtype);
  PLK_REQUIRE(batch > 0 && d > 0 && ldx >= d && bucket_size > 0 && batch % bucket_size == 0, PLK_ERR_INVALID,
              "bad shape batch=%lld d=%lld ldx=%lld bucket_size=%lld", (long long)batch, (long long)d,
              (long long)ldx, (long long)bucket_size);
  PLK_REQUIRE(((uintptr_t)state & 255) == 0, PLK_ERR_INVALID, "state must be 256-byte aligned");
  const ClipState L(op_dtype, batch, d);
  char* base = (char*)state;
  float* st = (float*)(base + L.stats);
  double* sums = (double*)(base + L.aux);
  const int64_t B = batch;
  cudaStream_t cst = (cudaStream_t)stream;
  int rc = plk_l2norm_pair_fwd(x, y, B, d, ldx, base + L.u, base + L.v, op_dtype, L.ld, st, st + B, st + 2 * B,
                               st + 3 * B, nullptr, 0, (float*)(base + L.aux), 16, stream);
  if (rc) return rc;
  if (op_dtype == PLK_F32)
    rc = siglip_fwd_f32((const float*)(base + L.u), (const float*)(base + L.v), L.ld, B, 0, B, d, bucket_size,
                        logit_scale, bias, st + 6 * B, sums, cst);
  else
    rc = siglip_fwd_tc16(base + L.u, base + L.v, op_dtype == PLK_F16, L.ld, B, 0, B, d, bucket_size, logit_scale, bias,
                         st + 6 * B, sums, cst);
  if (rc) return rc;
  return plk_siglip_loss(sums, B, loss_out, stream);
}

int plk_siglip_loss_backward(const float* grad_out, const float* x, const float* y, int64_t batch, int64_t d,
                             int64_t ldx, int op_dtype, int64_t bucket_size, const float* logit_scale,
                             const float* bias, void* state, void* workspace, float* dx, float* dy, float* dls,
                             float* dbias, void* stream) {
  PLK_REQUIRE(grad_out && x && y && logit_scale && bias && state && workspace && dx && dy && dls && dbias,
              PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype >= PLK_F32 && op_dtype <= PLK_F16, PLK_ERR_INVALID, "bad op_dtype %d", op_dtype);
  PLK_REQUIRE(batch > 0 && d > 0 && ldx >= d && bucket_size > 0 && batch % bucket_size == 0, PLK_ERR_INVALID,
              "bad shape");
  const ClipState L(op_dtype, batch, d);
  char* base = (char*)state;
  float* st = (float*)(base + L.stats);
  const double* sums = (const double*)(base + L.aux);
  float* gs2 = (float*)(base + L.aux + 32);
  const int64_t B = batch;
  const int parts = clip_parts(op_dtype, B, d, bucket_size);
  float* acc_x = (float*)workspace;
  float* acc_y = acc_x + (size_t)parts * B * d;
  cudaStream_t cst = (cudaStream_t)stream;
  int rc;
  if (op_dtype == PLK_F32) {
    rc = siglip_grad_f32((const float*)(base + L.u), (const float*)(base + L.v), L.ld, B, 0, B, d, bucket_size,
                         logit_scale, bias, acc_x, gs2, cst);
    if (rc) return rc;
    rc = siglip_grad_f32((const float*)(base + L.v), (const float*)(base + L.u), L.ld, B, 0, B, d, bucket_size,
                         logit_scale, bias, acc_y, nullptr, cst);
  } else {
    static const bool overlap = getenv("PLK_PDL") == nullptr || getenv("PLK_PDL")[0] != '0';
    rc = infonce_grad_pair_tc16(base + L.u, base + L.v, base + L.v, base + L.u, op_dtype == PLK_F16, L.ld, B, 0, B,
                                d, bucket_size, logit_scale, nullptr, nullptr, nullptr, nullptr, acc_x, acc_y, gs2,
                                cst, overlap ? 1 : 0, bias);
  }
  if (rc) return rc;
  return plk_siglip_grad_finish_pair(acc_x, acc_y, parts, x, y, B, d, ldx, st, st + B, st + 2 * B, st + 3 * B,
                                     st + 6 * B, logit_scale, bias, grad_out, B, gs2, sums, dx, dy, dls, dbias,
                                     stream);
}

size_t plk_topk_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int kc, int op_dtype) {
  return op_dtype == PLK_F32 ? topk_ws_f32(nq, ng, d, kc) : topk_ws_tc16(nq, ng, d, kc);
}

int plk_topk_candidates(const void* q, const void* g, int op_dtype, int64_t ld, const float* g_sqn,
                        int64_t nq, int64_t ng, int64_t d, int kc, int64_t gallery_offset,
                        int32_t* cand_idx, float* cand_key, void* workspace, size_t workspace_bytes,
                        void* stream) {
  PLK_REQUIRE(q && g && g_sqn && cand_idx && cand_key, PLK_ERR_INVALID, "null pointer");
  PLK_REQUIRE(op_dtype == PLK_F32 || op_dtype == PLK_BF16 || op_dtype == PLK_F16, PLK_ERR_INVALID, "bad op_dtype %d", op_dtype);
  PLK_REQUIRE(nq > 0 && ng > 0 && d > 0 && ld >= d, PLK_ERR_INVALID, "bad shape");
  PLK_REQUIRE(kc >= 1 && kc <= 64, PLK_ERR_INVALID, "kc must be in [1,64] (got %d)", kc);
  PLK_REQUIRE(gallery_offset >= 0 && gallery_offset + ng < (int64_t)1 << 31, PLK_ERR_INVALID,
              "gallery indices must fit int32");
  size_t need = plk_topk_workspace_bytes(nq, ng, d, kc, op_dtype);
  PLK_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), PLK_ERR_INVALID,
              "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (op_dtype == PLK_F32)
    return topk_candidates_f32((const float*)q, (const float*)g, ld, g_sqn, nq, ng, d, kc,
                               gallery_offset, cand_idx, cand_key, workspace, workspace_bytes, st);
  return topk_candidates_tc16(q, g, op_dtype == PLK_F16, ld, g_sqn, nq, ng, d, kc, gallery_offset, cand_idx,
                              cand_key, workspace, workspace_bytes, st);
}

}  // extern "C"
