/*
 * plk.h -- C ABI of libplk.so: the B200-native (sm_100a) cross-modal similarity
 * hot path of imveikka/multimodal_plankton_recognition.
 *
 * The path is S = normalize(I) . normalize(P)^T consumed by
 *   (1) the symmetric CLIP-style InfoNCE coordination loss, forward + backward
 *       (reference src/coordination.py:17-47, autograd backward of the same graph), and
 *   (2) euclidean / cosine top-k retrieval + inverse-distance weighted k-NN vote
 *       (reference src/ann.py:6-34, driven by reference scripts/benchmark_cross.py:24-96).
 *
 * Conventions (all entry points)
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - the caller allocates every input, output and workspace; the library never
 *     frees or retains a pointer past the call and never synchronises: all work
 *     is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = PLK_OK, negative = error (see enum); the message is
 *     available from plk_last_error() (thread-local).  No exceptions cross the ABI.
 *   - shapes are row-major; `ld*` arguments are leading dimensions in ELEMENTS.
 *   - operand dtype `op_dtype` selects the arithmetic path:
 *       PLK_F32  : fp32 CUDA-core path (parity mode, <=1e-5 rel vs the reference)
 *       PLK_BF16 : tcgen05 bf16 MMA, fp32 TMEM accumulators (<=2e-3 rel)
 *       PLK_F16  : the same kernels with fp16 operands (unit-norm rows fit fp16; 8x finer rounding)
 *     There is no CPU fallback.
 */
#ifndef PLK_H_
#define PLK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum plk_status {
  PLK_OK = 0,
  PLK_ERR_INVALID = -1,     /* bad shape / dtype / null pointer / alignment       */
  PLK_ERR_UNSUPPORTED = -2, /* valid request this build cannot serve (e.g. d>512 bf16) */
  PLK_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed                */
  PLK_ERR_ARCH = -4         /* device is not sm_100 (tcgen05 path needs it)       */
};

enum plk_dtype { PLK_F32 = 0, PLK_BF16 = 1, PLK_F16 = 2 };

int plk_version(void);
const char* plk_last_error(void);
/* 1 if the current device can run the tcgen05 (PLK_BF16) kernels. */
int plk_device_supports_tc(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t plk_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * a2. L2 normalisation   u = x / max(||x||_2, 1e-12)            reference src/coordination.py:33-34
 *   x        [n, d]      x_dtype in {F32, BF16, F16}, leading dim ldx
 *   u        [n, ldu]    u_dtype in {F32, BF16, F16}; columns d..ldu-1 are written as zero (the
 *                        tcgen05 path wants ldu = d rounded up to 64)
 *   inv_den  [n]  fp32   1 / max(||x||, eps)
 *   nrm      [n]  fp32   ||x||           (needed by the backward to detect the eps clamp)
 *   sqn      [n]  fp32   ||u||^2 of the values actually WRITTEN to u (after rounding);
 *                        nullable.  Retrieval uses it for |g|^2 - 2 q.g ranking.
 * ------------------------------------------------------------------------------------------ */
int plk_l2norm_fwd(const void* x, int x_dtype, int64_t n, int64_t d, int64_t ldx,
                   void* u, int u_dtype, int64_t ldu, float* inv_den, float* nrm, float* sqn,
                   int normalise /* 0: plain cast/copy (retrieval on pre-normalised data) */,
                   void* stream);

/* Both modalities in one launch (fp32 rows in, same shapes), optionally zero-filling two fp32
 * arrays (the sum-exp accumulators plk_infonce_fwd adds into; pass sums_zeroed = 1 there). */
int plk_l2norm_pair_fwd(const float* x, const float* y, int64_t n, int64_t d, int64_t ldx,
                        void* u, void* v, int u_dtype, int64_t ldu,
                        float* inv_den_x, float* nrm_x, float* inv_den_y, float* nrm_y,
                        float* zero_a, int64_t n_zero_a, float* zero_b, int64_t n_zero_b,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * a3-a7. Fused similarity + temperature + online row/column sum-exp + diagonal.
 *        replaces: bmm, *exp(logit_scale), 2x cross_entropy       reference src/coordination.py:36-44
 *
 *   u [n_rows, ld]   normalised rows OWNED by this call (global row index = row_offset + i)
 *   v [n_cols, ld]   normalised rows of the other modality for the WHOLE (global) batch
 *   S_ij = exp(*logit_scale) * u_i . v_j, restricted to j in the bucket of global row i
 *          (bucket b = global_i / bucket_size, columns [b*bs, (b+1)*bs)).
 *   Fixed, range-centred shift: because |u.v| <= 1, E_ij = exp(S_ij - s + 64) <= e^64 never overflows
 *   (nor does a sum of up to e^24 of them), so ONE exp serves the row and the column sums.  A row keeps
 *   a non-zero sum while its best match satisfies s*(1 - cos_max) < 151: any data for
 *   s = exp(logit_scale) <= 75 (logit_scale <= 4.3), and s <= 151 (logit_scale <= 5.0) when every row
 *   has a non-negative best cosine.  Beyond that plk_infonce_loss returns NaN (never a finite-looking
 *   wrong loss); the reference's F.cross_entropy has no such limit.
 *
 *   row_sumexp [n_rows]  OUT  sum_j E_ij                (complete)
 *   col_sumexp [n_cols]  OUT  sum_{i owned} E_ij        (partial when rows are sharded)
 *   diag       [n_rows]  OUT  S_ii (global diagonal)
 *   logit_scale          device pointer to the 0-dim fp32 parameter (no host sync).
 * ------------------------------------------------------------------------------------------ */
int plk_infonce_fwd(const void* u, const void* v, int op_dtype, int64_t ld,
                    int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d,
                    int64_t bucket_size, const float* logit_scale,
                    float* row_sumexp, float* col_sumexp, float* diag,
                    int sums_zeroed /* 1: row/col_sumexp were already zero-filled (plk_l2norm_pair_fwd) */,
                    void* stream);

/* a8. loss partial over owned rows:
 *   *loss_out = (1/(2*B_global)) * sum_i [ 2 (s - 64) + log R_i + log C_i - 2 S_ii ]   reference src/coordination.py:45
 *   col_sumexp_own points at the n_rows entries of the (all-reduced) column sums that
 *   belong to the owned rows.  Also writes sum_i S_ii to *diag_sum_out (used by d logit_scale) and,
 *   if gs_zero is not NULL, stores 0 to *gs_zero (the accumulator plk_infonce_grad adds into). */
int plk_infonce_loss(const float* row_sumexp, const float* col_sumexp_own, const float* diag,
                     const float* logit_scale, int64_t n_rows, int64_t batch_global,
                     float* loss_out, float* diag_sum_out, float* gs_zero, void* stream);

/* Same for a rank of a bucket-aligned sharded step, with the sum over the ranks fused into the kernel
 * (peer-mapped symmetric memory, arguments as plk_infonce_grad_finish_pair_xgpu):
 *   *loss_out = GLOBAL loss (identical bits on every rank), *partial_out = this rank's partial.
 * All ranks must call it the same number of times, interleaved identically with the other _xgpu calls. */
int plk_infonce_loss_xgpu(const float* row_sumexp, const float* col_sumexp_own, const float* diag,
                          const float* logit_scale, int64_t n_rows, int64_t batch_global,
                          float* loss_out, float* diag_sum_out, float* gs_zero, float* partial_out,
                          void* const* peer_bufs, int rank, int world, unsigned* epoch, float* out2,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * a9. Recompute backward, one direction (flash-style: logits are never materialised).
 *   acc_i = sum_{j != i}  E_ij (1/rs_i + 1/cs_j) * b_j     for the owned rows i (global j != i)
 *   a [n_rows, ld] owned normalised rows, b [n_cols, ld] all normalised rows of the other
 *   modality, rs [n_rows] the sum-exp along a's rows, cs [n_cols] the sum-exp along b's rows.
 *   Direction image:   (a,b,rs,cs) = (u, v, row_sumexp, col_sumexp)
 *   Direction profile: (a,b,rs,cs) = (v, u, col_sumexp, row_sumexp)   (S is symmetric in roles)
 *   acc  [parts, n_rows, d] fp32 OUT: `parts` partial sums (column sweep split across CTAs to fill
 *        the 148 SMs; parts = plk_infonce_grad_parts(...)); plk_infonce_grad_finish adds them.
 *   gs   nullable; IN/OUT scalar, ADDED to: sum_ij E_ij (1/rs_i + 1/cs_j) S_ij, j == i included (for
 *        d logit_scale).  The caller zero-initialises it (plk_infonce_loss does, via gs_zero).
 * ------------------------------------------------------------------------------------------ */
int plk_infonce_grad_parts(int op_dtype, int64_t n_rows, int64_t n_cols, int64_t d,
                           int64_t bucket_size);
int plk_infonce_grad(const void* a, const void* b, int op_dtype, int64_t ld,
                     int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d,
                     int64_t bucket_size, const float* logit_scale,
                     const float* rs, const float* cs, float* acc, float* gs, void* stream);

/* Both directions in ONE launch (fills the SMs with half as many column segments at small batches):
 *   direction 0: (a0, b0, rs0, cs0) -> acc0 [parts, n_rows, d]  (+= gs)
 *   direction 1: (a1, b1, rs1, cs1) -> acc1 [parts, n_rows, d]
 * with parts = plk_infonce_grad_pair_parts(...).  All other arguments as plk_infonce_grad. */
int plk_infonce_grad_pair_parts(int op_dtype, int64_t n_rows, int64_t n_cols, int64_t d,
                                int64_t bucket_size);
int plk_infonce_grad_pair(const void* a0, const void* b0, const void* a1, const void* b1, int op_dtype,
                          int64_t ld, int64_t n_rows, int64_t row_offset, int64_t n_cols, int64_t d,
                          int64_t bucket_size, const float* logit_scale,
                          const float* rs0, const float* cs0, const float* rs1, const float* cs1,
                          float* acc0, float* acc1, float* gs, void* stream);

/* a9 (tail). Adds the j == i term and the -2*delta_ij term in fp32, applies g*s/(2B) and the
 * normalisation backward:
 *   acc_i = sum over the `parts` slabs of acc  (plk_infonce_grad leaves the j == i term out);
 *   dU_i = coef * (acc_i + (E_ii (1/rs_i + 1/cs_i) - 2) p_i / den_p_i),  E_ii = exp(diag_i - s + 64),
 *          coef = (*grad_out) * s / (2 B_global)
 *   dx_i = (dU_i - u_i (u_i . dU_i)) / den_i      if ||x_i|| > eps,   else dU_i / eps
 *   x, partner: RAW fp32 embeddings [n, d] of this modality / the other one (same rows);
 *   diag, rs, cs: S_ii and the two sum-exps of the owned rows.  dx [n, d] written in dx_dtype. */
int plk_infonce_grad_finish(const float* acc, int parts, const void* x, const void* partner,
                            int x_dtype, int64_t n, int64_t d, int64_t ldx,
                            const float* inv_den_x, const float* nrm_x, const float* inv_den_p,
                            const float* diag, const float* rs, const float* cs,
                            const float* logit_scale, const float* grad_out, int64_t batch_global,
                            void* dx, int dx_dtype, void* stream);

/* plk_infonce_grad_finish for both modalities + plk_infonce_dls in ONE launch (fp32 rows in/out):
 *   dx from (acc_x, x, partner y), dy from (acc_y, y, partner x); embedding gradients use
 *   *grad_out_emb, d logit_scale uses *grad_out (they differ by the world size under DDP scaling);
 *   *dls_out = *grad_out / (2B) * (*gs - 2 * *diag_sum), after which *gs is reset to 0 (consumed). */
int plk_infonce_grad_finish_pair(const float* acc_x, const float* acc_y, int parts,
                                 const float* x, const float* y, int64_t n, int64_t d, int64_t ldx,
                                 const float* inv_den_x, const float* nrm_x,
                                 const float* inv_den_y, const float* nrm_y,
                                 const float* diag, const float* rs, const float* cs,
                                 const float* logit_scale, const float* grad_out_emb,
                                 const float* grad_out, int64_t batch_global, float* gs,
                                 const float* diag_sum, float* dx, float* dy, float* dls_out,
                                 void* stream);

/* Same, for a rank of a sharded step, with the cross-GPU sum of (loss partial, d logit_scale
 * partial) fused into the kernel: the ranks exchange the two scalars through peer-mapped symmetric
 * memory over NVLink (release/acquire flags), replacing a separate NCCL all-reduce launch.
 *   loss_partial  this rank's loss partial (from plk_infonce_loss)
 *   peer_bufs     DEVICE array [world] of peer-mapped base pointers of one >= 128-byte,
 *                 zero-initialised symmetric buffer per rank (same order on every rank)
 *   epoch         TWO local device counters (uint32[2], zero-initialised once; each is incremented
 *                 per launch: the first block publishes, the last block collects, so rank skew is
 *                 absorbed by the kernel's row work)
 *   out2          OUT (global loss, global d logit_scale), bitwise identical on every rank
 * All ranks must launch it the same number of times (it waits for every peer, bounded by a trap). */
int plk_infonce_grad_finish_pair_xgpu(const float* acc_x, const float* acc_y, int parts,
                                      const float* x, const float* y, int64_t n, int64_t d, int64_t ldx,
                                      const float* inv_den_x, const float* nrm_x,
                                      const float* inv_den_y, const float* nrm_y,
                                      const float* diag, const float* rs, const float* cs,
                                      const float* logit_scale, const float* grad_out_emb,
                                      const float* grad_out, int64_t batch_global, float* gs,
                                      const float* diag_sum, float* dx, float* dy, float* dls_out,
                                      const float* loss_partial, void* const* peer_bufs, int rank,
                                      int world, unsigned* epoch, float* out2, void* stream);

/* d logit_scale partial:  *dls_out = (*grad_out) / (2 B_global) * (*gs - 2 * *diag_sum) */
int plk_infonce_dls(const float* gs, const float* diag_sum, const float* grad_out,
                    int64_t batch_global, float* dls_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a1-a11 in two calls: the whole single-GPU loss step, what `CLIPLoss.forward` and the autograd
 * backward of its result do.                                reference src/coordination.py:26-47
 *   x, y [B, d] raw fp32 embeddings (row stride ldx); logit_scale the learnable scalar (device).
 *   state      plk_clip_loss_state_bytes() bytes of device memory written by the forward and
 *              consumed by the backward (normalised 16-bit/fp32 operands, the per-row statistics,
 *              two scalars); 256-byte aligned.  One allocation instead of five.
 *   workspace  plk_clip_loss_workspace_bytes() bytes of scratch for the backward (the partial
 *              gradient slabs); contents are dead after the call.
 *   forward : *loss_out = loss.     Launches: normalise(both) -> fused similarity/sum-exp -> loss.
 *   backward: dx, dy [B, d] fp32 scaled by *grad_out_emb, *dls = d loss / d logit_scale scaled by
 *             *grad_out.  Launches: recompute backward (both directions, started under the tail of
 *             the preceding kernel -- programmatic serialization) -> gradient tail (both + dls).
 *   The backward may be called more than once on one state (it restores what it consumes).
 *   batch_global: the 1/(2 B) of the loss.  batch_global == batch on one GPU.  A rank of a
 *   bucket-aligned sharded step (every bucket entirely on one rank: the local problem is
 *   complete) passes the global batch: loss / dls are then this rank's PARTIAL sums.  The embedding
 *   gradients are scaled by (*grad_out_emb) * emb_scale; under DDP gradient averaging pass
 *   grad_out_emb = grad_out and emb_scale = world (emb_scale != 1 needs d % 128 == 0, d <= 1024).
 *   plk_clip_loss_forward_xgpu returns the GLOBAL loss (the ranks' partials are exchanged inside the
 *   loss kernel, plk_infonce_loss_xgpu); plk_clip_loss_backward_xgpu additionally sums (loss partial,
 *   dls partial) over the ranks inside the gradient-tail kernel (plk_infonce_grad_finish_pair_xgpu):
 *   a sharded step without a single collective launch.
 * ------------------------------------------------------------------------------------------ */
size_t plk_clip_loss_state_bytes(int op_dtype, int64_t batch, int64_t d);
size_t plk_clip_loss_workspace_bytes(int op_dtype, int64_t batch, int64_t d, int64_t bucket_size);
int plk_clip_loss_forward(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx,
                          int op_dtype, int64_t bucket_size, int64_t batch_global,
                          const float* logit_scale, void* state, float* loss_out, void* stream);
int plk_clip_loss_forward_xgpu(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx,
                               int op_dtype, int64_t bucket_size, int64_t batch_global,
                               const float* logit_scale, void* state, float* loss_out,
                               float* partial_out, void* const* peer_bufs, int rank, int world,
                               unsigned* epoch, float* out2, void* stream);
int plk_clip_loss_backward(const float* grad_out, const float* grad_out_emb, float emb_scale,
                           const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype,
                           int64_t bucket_size, int64_t batch_global, const float* logit_scale,
                           void* state, void* workspace, float* dx, float* dy, float* dls,
                           void* stream);
int plk_clip_loss_backward_xgpu(const float* grad_out, const float* grad_out_emb, float emb_scale,
                                const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx, int op_dtype,
                                int64_t bucket_size, int64_t batch_global, const float* logit_scale,
                                void* state, void* workspace, float* dx, float* dy, float* dls,
                                const float* loss_partial, void* const* peer_bufs, int rank,
                                int world, unsigned* epoch, float* out2, void* stream);

/* ------------------------------------------------------------------------------------------
 * N2. SigLIP pairwise-sigmoid loss on the same similarity mainloop.
 *      replaces SigLIPLoss.forward and its autograd backward       reference src/coordination.py:67-95
 *   z_ij = exp(*logit_scale) u_i.v_j + *bias  within the bucket of row i;
 *   loss = (1/B) [ sum_{i != j} softplus(z_ij) + sum_i softplus(-z_ii) ]
 *   state / workspace: plk_clip_loss_state_bytes() / plk_clip_loss_workspace_bytes() (same layout).
 *   forward : normalise(both) -> fused similarity + softplus sum (no row/column statistics) -> loss.
 *   backward: recompute backward with G = sigmoid(z) off the diagonal -> gradient tail (diagonal term
 *             -sigmoid(-z_ii) in fp32) -> dx, dy [B, d] fp32, *dls, *dbias, all scaled by *grad_out.
 *   Building blocks: plk_siglip_loss (loss = sums[0] / B) and plk_siglip_grad_finish_pair
 *   (sums = double[3] from the forward: loss terms, sum_i G_ii S_ii, sum_i G_ii; gs2 = float[2] from the
 *   recompute backward: sum G*S, sum G over the off-diagonal; consumed = reset to 0).
 * ------------------------------------------------------------------------------------------ */
int plk_siglip_loss_forward(const float* x, const float* y, int64_t batch, int64_t d, int64_t ldx,
                            int op_dtype, int64_t bucket_size, const float* logit_scale,
                            const float* bias, void* state, float* loss_out, void* stream);
int plk_siglip_loss_backward(const float* grad_out, const float* x, const float* y, int64_t batch,
                             int64_t d, int64_t ldx, int op_dtype, int64_t bucket_size,
                             const float* logit_scale, const float* bias, void* state,
                             void* workspace, float* dx, float* dy, float* dls, float* dbias,
                             void* stream);
int plk_siglip_loss(const double* sums, int64_t batch, float* loss_out, void* stream);
int plk_siglip_grad_finish_pair(const float* acc_x, const float* acc_y, int parts, const float* x,
                                const float* y, int64_t n, int64_t d, int64_t ldx,
                                const float* inv_den_x, const float* nrm_x, const float* inv_den_y,
                                const float* nrm_y, const float* diag, const float* logit_scale,
                                const float* bias, const float* grad_out, int64_t batch, float* gs2,
                                const double* sums, float* dx, float* dy, float* dls_out,
                                float* dbias_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * N1. Bias-free projection Linear fused with the normalisation that opens the loss:
 *        replaces: nn.Linear(dim_out, dim_embed, bias=False) + F.normalize
 *                  reference src/model.py:29-30,:38-39,:80-83 and src/coordination.py:33-34
 *   emb_i = sum_k feat[i,k] w[j,k]  (16-bit operands, fp32 accumulation);  u = emb / max(||emb||, 1e-12)
 *   feat16 [n, ldf], w16 [d, ldw]   op_dtype (PLK_BF16 / PLK_F16) rows, zero-padded to ceil(f/64)*64 columns
 *   u16    [n, ldu] OUT  normalised operand of plk_infonce_fwd / plk_infonce_grad, ldu = ceil(d/64)*64
 *   emb    [n, d]   OUT  raw fp32 embedding (the `x` / `partner` rows of plk_infonce_grad_finish)
 *   inv_den, nrm [n] OUT as plk_l2norm_fwd.
 * One tcgen05 kernel: accumulator [128 x d] in TMEM (d <= 512), squared norms taken in the epilogue.
 * ------------------------------------------------------------------------------------------ */
int plk_project_normalise(const void* feat16, int64_t ldf, const void* w16, int64_t ldw, int op_dtype,
                          int64_t n, int64_t f, int64_t d, void* u16, int64_t ldu, float* emb,
                          float* inv_den, float* nrm, void* stream);

/* ------------------------------------------------------------------------------------------
 * Staging host-resident batches into HBM under the running step (no reference counterpart: the
 * reference hands each batch to the device serially through Lightning's loop,
 * reference scripts/train_multi.py:78-85).  A stager owns a copy stream and per-slot events:
 *   issue(slot)             host: wait until the slot's PREVIOUS copy has finished (its host buffers may be
 *                           released once this call returns); copy stream: wait until `slot` was released,
 *                           copy x (and y) host->device, mark the slot ready.  Host buffers must be
 *                           page-locked to overlap and must stay alive until the slot's next issue returns.
 *   acquire(slot, release)  consumer stream: mark `release` (>= 0) reusable after everything queued
 *                           so far, then wait until `slot` is ready.
 *   read_async / read_wait  device->host copy of a result on the consumer stream + an event the
 *                           host can wait on alone (reading a loss does not drain the queue).
 * Handles are bound to the device current at creation.  NULL from create => plk_last_error().
 * ------------------------------------------------------------------------------------------ */
void* plk_stager_create(int depth, int read_slots);
void plk_stager_destroy(void* stager);
int plk_stager_issue(void* stager, int slot, void* dst_x, const void* src_x, size_t bytes_x,
                     void* dst_y, const void* src_y, size_t bytes_y);
int plk_stager_acquire(void* stager, int slot, int release_slot, void* consumer_stream);
int plk_stager_read_async(void* stager, int read_slot, void* dst_host, const void* src_dev,
                          size_t bytes, void* consumer_stream);
int plk_stager_read_wait(void* stager, int read_slot);

/* ------------------------------------------------------------------------------------------
 * a12. k-nearest candidates by euclidean distance (ranking key |g|^2 - 2 q.g).
 *      replaces pynndescent NNDescent.query                        reference src/ann.py:15-16
 *   q [nq, ld], g [ng, ld] in op_dtype; g_sqn [ng] = ||g_j||^2 (from plk_l2norm_fwd).
 *   Writes the kc best gallery indices per query (unordered quality: sorted by the
 *   op_dtype key, ascending), global index = gallery_offset + local row.
 *   cand_idx [nq, kc] int32 (-1 pads when ng < kc), cand_key [nq, kc] fp32.
 *   workspace: plk_topk_workspace_bytes() bytes.
 * ------------------------------------------------------------------------------------------ */
size_t plk_topk_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int kc, int op_dtype);
int plk_topk_candidates(const void* q, const void* g, int op_dtype, int64_t ld,
                        const float* g_sqn, int64_t nq, int64_t ng, int64_t d, int kc,
                        int64_t gallery_offset, int32_t* cand_idx, float* cand_key,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Exact re-score of candidates: dist = sqrt(sum (q-g)^2) accumulated in fp64 from the fp32
 * vectors, rounded to fp32, sorted ascending by (dist, index); keeps the k best.
 *   q32 [nq, d], g32 [ng, d] fp32; cand_idx [nq, m] global indices, -1 = empty;
 *   local gallery row = cand_idx - gallery_offset.
 *   out_idx [nq, k] int32, out_dist [nq, k] fp32 (+inf / -1 pads). */
int plk_topk_rescore(const float* q32, const float* g32, int64_t nq, int64_t ng, int64_t d,
                     const int32_t* cand_idx, int m, int64_t gallery_offset, int k,
                     float* scratch /* [nq, m] fp32 */, int32_t* out_idx, float* out_dist,
                     void* stream);

/* Merge per-shard results: cand [nq, m] (idx, dist) -> k best by (dist, idx).  -1 = empty. */
int plk_topk_merge(const int32_t* cand_idx, const float* cand_dist, int64_t nq, int m, int k,
                   int32_t* out_idx, float* out_dist, void* stream);

/* ------------------------------------------------------------------------------------------
 * a13/a14. inverse-distance weighted vote                          reference src/ann.py:19-34
 *   idx/dist [nq, m] neighbour lists (already h-stacked over query modalities),
 *   labels [ng] int64 class of each gallery row.  w = 1/dist (fp32); a row containing a
 *   zero distance votes only with its zero-distance neighbours (weight 1).  Per-class sums
 *   in fp64, ties -> lowest class id.   pred [nq] int64.
 * ------------------------------------------------------------------------------------------ */
int plk_knn_vote(const int32_t* idx, const float* dist, int64_t nq, int m,
                 const int64_t* labels, int64_t ng, int64_t* pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLK_H_ */
