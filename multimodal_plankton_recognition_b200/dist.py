"""Multi-GPU partitioning of the hot path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch) for the exchange steps.  New functionality relative to the reference, which
is single-device; the single-process reference on the concatenated global batch is the oracle.

Loss (row-block sharding).  Rank r owns rows [r*n, (r+1)*n) of BOTH modalities.
    forward : normalise locally -> all-gather u_hat, v_hat (operand dtype)
              fused kernel on S[own rows, all columns] -> complete row sums, PARTIAL column sums
              all-reduce column sums [B];  all-gather row sums [B];  all-reduce scalar loss
    backward: two recompute passes, both complete locally (no gradient reduce-scatter):
              d image[own]   from (u_own, V_all, R_own, C_all)
              d profile[own] from (v_own, U_all, C_own, R_all)
              all-reduce d logit_scale
Retrieval (gallery sharding).  Every rank searches its gallery shard for all queries, the
per-shard (index, exact distance) lists are all-gathered and merged on the device.

All local compute goes through ``ops.*`` / ``ann.*`` (the C ABI); this file only adds collectives.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import ops


def _world(group):
    return dist.get_world_size(group), dist.get_rank(group)


def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    R, _ = _world(group)
    out = torch.empty((R * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def _all_gather_rows_pair(a: torch.Tensor, b: torch.Tensor, group):
    """all-gather two row blocks (two collectives: c10d's coalescing manager gave wrong data intermittently
    with all_gather_into_tensor on this stack, so the two launches are not grouped)."""
    return _all_gather_rows(a, group), _all_gather_rows(b, group)


class XGpuScalars:
    """Peer-mapped symmetric buffer for the fused cross-GPU exchange of the two per-rank scalars
    (loss partial, d logit_scale partial) inside plk_infonce_grad_finish_pair_xgpu -- torch's
    symmetric-memory allocator is used only to allocate and exchange the IPC mappings."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(64, dtype=torch.float32, device=device)   # 8 B x 2 parities + flags at +64 B
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)   # publisher / collector launch counters
        self.out2 = torch.zeros(2, dtype=torch.float32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)       # every rank's zero-fill is visible before anyone signals


def xgpu_supported(image_emb, profile_emb) -> bool:
    """The fused exchange lives in the VECTORISED loss / gradient-tail kernels: d % 128 == 0, d <= 1024 and
    16-byte aligned fp32 rows (plk_infonce_grad_finish_pair_xgpu rejects anything else).  Every rank takes
    the same decision from the same shapes; otherwise the two scalars go through NCCL all-reduces."""
    d = image_emb.shape[1]
    if d % 128 or d > 1024:
        return False
    for t in (image_emb, profile_emb):
        if t.dtype == torch.float32 and t.stride(-1) == 1 and (t.stride(0) % 4 or t.data_ptr() % 16):
            return False
    return True


def sharded_fwd(image_emb, profile_emb, logit_scale, buckets, mode, group, reduce_scalars=True, xgpu=None):
    """Forward of the row-block sharded loss without autograd: -> (global loss [], saved state).
    With reduce_scalars=False the returned loss is this rank's partial sum (the caller all-reduces
    it); in the bucket-aligned case the call then contains no collective at all.  With
    reduce_scalars=True and `xgpu` (XGpuScalars) the bucket-aligned case has none either: the loss
    kernel itself sums the partials over the ranks through NVLink peer memory."""
    R, r = _world(group)
    n, d = image_emb.shape
    B = n * R
    assert B % buckets == 0, "Batch size must be divisible by number of buckets!"
    bs = B // buckets
    off = r * n
    x, y = ops._as_f32_rows(image_emb), ops._as_f32_rows(profile_emb)
    ls = logit_scale.detach().float()
    scal = torch.empty(2, device=x.device, dtype=torch.float32)   # (loss, d logit_scale) partials side by side
    if n % bs == 0:
        # every bucket lives entirely on one rank (block-diagonal logits): no data-path exchange, the
        # local problem is complete -- the single-GPU composite step with the global 1/(2B); only the
        # two scalars are reduced.
        if reduce_scalars and xgpu is not None:
            loss, st = ops.clip_loss_forward_state(x, y, ls, bs, mode, batch_global=B, xgpu=xgpu, partial_out=scal[0:1])
        else:
            loss, st = ops.clip_loss_forward_state(x, y, ls, bs, mode, batch_global=B, loss_out=scal[0])
            if reduce_scalars:
                dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        return loss, (x, y, ls, st, scal, (True, n, d, B, bs, off, mode, group))
    # General case: one exchange each way.  Four collectives per step instead of seven:
    #   the two all-gathers of u_hat, v_hat; ONE all-reduce of [3, B] statistics -- partial
    #   column sums, and the row sums / diagonal logits scattered at the owned offsets (zeros elsewhere), which
    #   turns their all-gathers into the same sum; the scalar loss needs no collective at all: every rank
    #   evaluates it over the B global rows from the reduced statistics (identical bits everywhere).
    st4 = torch.empty((4, n), device=x.device, dtype=torch.float32)
    stats = torch.empty((3, B), device=x.device, dtype=torch.float32)      # cs | rs | diag, global row order
    cs_all, rs_all, dg_all = stats.unbind(0)
    rs, dg = rs_all[off:off + n], dg_all[off:off + n]
    u, v = ops.l2norm_pair(x, y, mode, st4, stats)            # also zero-fills `stats`
    # the forward needs the gathered v_hat only; u_hat (the streamed operand of the backward's second
    # direction) is gathered asynchronously under the forward kernel and waited for in sharded_bwd
    v_all = _all_gather_rows(v, group)
    u_all = torch.empty((B,) + tuple(u.shape[1:]), device=u.device, dtype=u.dtype)
    u_work = dist.all_gather_into_tensor(u_all, u, group=group, async_op=True)
    idx, nx, idy, ny = st4.unbind(0)
    ops.infonce_fwd_local(u, v_all, mode, d, off, bs, ls, rs, cs_all, dg, sums_zeroed=True)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    part, aux = ops.infonce_loss_local(rs, cs_all[off:off + n], dg, ls, B, scal[0])    # owned rows: partial + sum of own diagonal
    if reduce_scalars:
        loss, _ = ops.infonce_loss_local(rs_all, cs_all, dg_all, ls, B)                  # all rows: the global loss
    else:
        loss = part
    state = (x, y, ls, u, v, u_all, v_all, idx, nx, idy, ny, rs_all, cs_all, dg, aux, scal, u_work,
             (False, n, d, B, bs, off, mode, group))
    return loss, state


def sharded_bwd(state, grad_out, grad_scale="ddp", out_dtypes=(torch.float32, torch.float32),
                reduce_scalars=True, xgpu=None):
    """Backward of `sharded_fwd`: -> (d image_emb [n,d], d profile_emb [n,d], d logit_scale []).
    With reduce_scalars=False d logit_scale is this rank's partial (the caller all-reduces it) --
    unless `xgpu` (XGpuScalars) is given: then the gradient-tail kernel itself sums (loss partial,
    d logit_scale partial) over the ranks through NVLink peer memory and xgpu.out2 holds the
    global (loss, d logit_scale)."""
    meta = state[-1]
    aligned, n, d, B, bs, off, mode, group = meta
    R, _ = _world(group)
    go = grad_out.detach().float().reshape(1).contiguous()
    # grad_scale == "ddp": DistributedDataParallel AVERAGES parameter gradients over ranks, while
    # each rank holds the exact d(global loss)/d(local rows); pre-multiplying by the world size
    # makes the averaged encoder gradients equal the true global-batch gradients.
    scale = float(R) if grad_scale == "ddp" else 1.0
    if aligned:
        x, y, ls, st, scal = state[:-1]
        in_kernel = d % 128 == 0 and d <= 1024      # the vectorised gradient tail applies the factor itself
        dx, dy, dls = ops.clip_loss_backward_state(go, x, y, ls, st, bs, mode, batch_global=B,
                                                   go_emb=None if in_kernel or scale == 1.0 else go * scale,
                                                   emb_scale=scale if in_kernel else 1.0,
                                                   dls_out=scal[1] if not reduce_scalars else None, xgpu=xgpu,
                                                   loss_partial=scal[0:1] if xgpu is not None else None)
    else:
        go_emb = go * scale if scale != 1.0 else go
        x, y, ls, u, v, u_all, v_all, idx, nx, idy, ny, rs_all, cs_all, dg, aux, scal, u_work = state[:-1]
        if u_work is not None:
            u_work.wait()        # the asynchronous all-gather of u_hat issued in sharded_fwd
        rs_own, cs_own = rs_all[off:off + n], cs_all[off:off + n]
        gs = aux[1:]
        acc_x, acc_y = ops.infonce_grad_pair_local(u, v_all, v, u_all, mode, d, off, bs, ls, rs_own, cs_all,
                                                   cs_own, rs_all, gs)
        dls_work = None
        if reduce_scalars and xgpu is None:
            # d logit_scale is final once the recompute kernel has added sum G*S: reduce it over the ranks
            # under the gradient tail instead of after it
            dls_early = ops.infonce_dls(gs, aux[0:1], go, B)
            dls_work = dist.all_reduce(dls_early, op=dist.ReduceOp.SUM, group=group, async_op=True)
        dx, dy, dls = ops.infonce_grad_finish_pair(acc_x, acc_y, x, y, (idx, nx), (idy, ny), dg, rs_own, cs_own, ls,
                                                   go_emb, go, B, gs, aux[0:1],
                                                   scal[1] if not reduce_scalars else None,
                                                   xgpu=xgpu, loss_partial=scal[0:1] if xgpu is not None else None)
        if dls_work is not None:
            dls_work.wait()
            return dx.to(out_dtypes[0]), dy.to(out_dtypes[1]), dls_early
    dx, dy = dx.to(out_dtypes[0]), dy.to(out_dtypes[1])
    if reduce_scalars:
        if xgpu is not None and aligned:
            dls = xgpu.out2[1].clone()      # summed over the ranks inside the gradient-tail kernel
        else:
            dist.all_reduce(dls, op=dist.ReduceOp.SUM, group=group)   # identical on every rank afterwards
    return dx, dy, dls


class _ShardedClipLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, image_emb, profile_emb, logit_scale, buckets, mode, group, grad_scale, xgpu):
        if xgpu is not None and not xgpu_supported(image_emb, profile_emb):
            xgpu = None          # same decision in the backward: it reads it from ctx.meta
        loss, state = sharded_fwd(image_emb, profile_emb, logit_scale, buckets, mode, group, xgpu=xgpu)
        ctx.plk_state = state    # intermediates only (detached fp32 rows, opaque buffers): no graph edges
        ctx.meta = (grad_scale, image_emb.dtype, profile_emb.dtype, logit_scale.dtype, xgpu)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        grad_scale, dt_x, dt_y, dt_ls, xgpu = ctx.meta
        dx, dy, dls = sharded_bwd(ctx.plk_state, g, grad_scale, (dt_x, dt_y), xgpu=xgpu)
        return dx, dy, dls.to(dt_ls), None, None, None, None, None


def sharded_clip_loss(image_emb, profile_emb, logit_scale, buckets: int = 1, mode: int = ops.PLK_BF16,
                      group=None, grad_scale: str = "ddp", xgpu=None) -> torch.Tensor:
    """Global-batch symmetric InfoNCE from per-rank row blocks (every rank passes n rows; the
    global batch is the rank-ordered concatenation).  Returns the global loss on every rank.
    `xgpu` (XGpuScalars of the same group): in the bucket-aligned case the two scalar sums run inside
    the kernels over NVLink peer memory and the step contains no collective launch."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("sharded_clip_loss needs an initialised torch.distributed process group")
    if grad_scale not in ("ddp", "none"):
        raise ValueError("grad_scale must be 'ddp' or 'none'")
    return _ShardedClipLoss.apply(image_emb, profile_emb, logit_scale, int(buckets), int(mode), group, grad_scale,
                                  xgpu)


def merge_shard_results(idx: torch.Tensor, dst: torch.Tensor, k: int, group=None):
    """All-gather the per-shard (global index, exact distance) lists [nq, k] and keep the k best per
    query by (distance, index) on the device (plk_topk_merge)."""
    from . import ann
    R, _ = _world(group)
    nq = idx.shape[0]
    all_i = torch.empty((R * nq, k), device=idx.device, dtype=torch.int32)
    all_d = torch.empty((R * nq, k), device=idx.device, dtype=torch.float32)
    dist.all_gather_into_tensor(all_i, idx.contiguous(), group=group)
    dist.all_gather_into_tensor(all_d, dst.contiguous(), group=group)
    cand_i = all_i.view(R, nq, k).permute(1, 0, 2).reshape(nq, R * k).contiguous()
    cand_d = all_d.view(R, nq, k).permute(1, 0, 2).reshape(nq, R * k).contiguous()
    return ann.topk_merge_device(cand_i, cand_d, k)


class ShardedANNClassifier:
    """Gallery-sharded counterpart of ``ANNClassifier``: each rank passes ITS shard (X, y); queries
    are replicated.  Global gallery index = rank-ordered concatenation of the shards."""

    def __init__(self, X, y, group=None, **nndescent_args):
        from . import ann
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ShardedANNClassifier needs an initialised torch.distributed process group")
        self.group = group
        R, r = _world(group)
        precision = nndescent_args.pop("plk_precision", "bf16")
        device = nndescent_args.pop("plk_device", None)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._exchange_layout(len(X), dev)
        self.index = ann.GpuExactIndex(X, precision=precision, device=dev, gallery_offset=self.offset)
        self._exchange_labels(y, dev)

    @classmethod
    def from_device(cls, g32: torch.Tensor, y, group=None, precision: str = "bf16"):
        """Same over a gallery shard that already lives in HBM (fp32 [n, d] on this rank's GPU)."""
        from . import ann
        self = cls.__new__(cls)
        self.group = group
        self._exchange_layout(g32.shape[0], g32.device)
        self.index = ann.GpuExactIndex.from_device(g32, precision=precision, gallery_offset=self.offset)
        self._exchange_labels(y, g32.device)
        return self

    def _exchange_layout(self, n_local, dev):
        R, r = _world(self.group)
        counts = torch.zeros(R, dtype=torch.int64, device=dev)
        counts[r] = n_local
        dist.all_reduce(counts, group=self.group)
        self.counts = counts.cpu().tolist()
        self.offset = int(sum(self.counts[:r]))
        self.total = int(sum(self.counts))

    def _exchange_labels(self, y, dev):
        y = y.detach().cpu().numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
        labels = torch.zeros(self.total, dtype=torch.int64, device=dev)
        labels[self.offset:self.offset + len(y)] = torch.from_numpy(y.astype(np.int64)).to(dev)
        dist.all_reduce(labels, group=self.group)
        self._labels_dev = labels
        self.y_ = labels.cpu().numpy()

    def search_device(self, q32: torch.Tensor, k: int):
        k_loc = min(k, self.index.n)
        idx, dst = self.index.search_device(q32, k_loc)
        if k_loc < k:  # pad short shards with empty slots
            pad_i = torch.full((idx.shape[0], k - k_loc), -1, device=idx.device, dtype=torch.int32)
            pad_d = torch.full((idx.shape[0], k - k_loc), float("inf"), device=idx.device)
            idx, dst = torch.cat((idx, pad_i), 1).contiguous(), torch.cat((dst, pad_d), 1).contiguous()
        return merge_shard_results(idx, dst, k, self.group)

    def kneighbors(self, *X, **query_args):
        k = min(int(query_args.get("k", 10)), self.total)
        out = []
        for x in X:
            q32 = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.index.device)
            i, d_ = self.search_device(q32, k)
            out.append((i.cpu().numpy(), d_.cpu().numpy()))
        return tuple(out)

    def predict(self, *X, **query_args):
        from . import ann
        k = min(int(query_args.get("k", 10)), self.total)
        lists = [self.search_device(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.index.device), k)
                 for x in X]
        idx = torch.cat([p[0] for p in lists], dim=1).contiguous()
        dst = torch.cat([p[1] for p in lists], dim=1).contiguous()
        return ann.knn_vote_device(idx, dst, self._labels_dev).cpu().numpy().astype(int).ravel()
