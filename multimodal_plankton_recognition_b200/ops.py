"""Torch-level plumbing over the C ABI (``include/plk.h``): device buffers, the current CUDA
stream, and the ``plk::clip_loss_fwd`` / ``plk::clip_loss_bwd`` custom ops with autograd.

PyTorch is used here for device memory, streams and autograd wiring only; every arithmetic
step of the hot path runs in ``libplk.so``.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import PLK_BF16, PLK_F16, PLK_F32

MODES = {"fp32": PLK_F32, "bf16": PLK_BF16, "fp16": PLK_F16}
OP_TORCH_DTYPE = {PLK_F32: torch.float32, PLK_BF16: torch.bfloat16, PLK_F16: torch.float16}
_DT = {torch.float32: PLK_F32, torch.bfloat16: PLK_BF16, torch.float16: PLK_F16}


def _stream(t: torch.Tensor) -> int:
    idx = t.device.index
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if idx is None else idx)


def _require_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("multimodal_plankton_recognition_b200 runs on CUDA (sm_100a) only; "
                               "there is no CPU fallback -- move the tensors to a B200")


def padded_width(d: int, mode: int) -> int:
    return d if mode == PLK_F32 else (d + 63) // 64 * 64


def l2norm(x: torch.Tensor, mode: int, normalise: bool = True, inv_den=None, nrm=None, want_sqn=False):
    """-> (u [n, ld] operand dtype, inv_den [n], nrm [n], sqn [n] or None)   (plk_l2norm_fwd)
    `inv_den` / `nrm` may be caller-provided fp32 [n] rows (e.g. slices of a stats buffer)."""
    lib = _lib.load()
    n, d = x.shape
    ld = padded_width(d, mode)
    u = torch.empty((n, ld), device=x.device, dtype=OP_TORCH_DTYPE[mode])
    if inv_den is None:
        inv_den = torch.empty(n, device=x.device, dtype=torch.float32)
    if nrm is None:
        nrm = torch.empty(n, device=x.device, dtype=torch.float32)
    sqn = torch.empty(n, device=x.device, dtype=torch.float32) if want_sqn else None
    with torch.cuda.device(x.device):
        lib.check(lib.plk_l2norm_fwd(x.data_ptr(), _DT[x.dtype], n, d, x.stride(0), u.data_ptr(), mode, ld,
                                     inv_den.data_ptr(), nrm.data_ptr(), sqn.data_ptr() if want_sqn else None,
                                     1 if normalise else 0, _stream(x)), "plk_l2norm_fwd")
    return u, inv_den, nrm, sqn


def l2norm_pair(x, y, mode, stats, zero_a=None, zero_b=None):
    """Normalise both modalities in one launch; writes (1/den, |.|) into stats rows 0..3 and
    zero-fills the two optional fp32 vectors.  -> (u, v)   (plk_l2norm_pair_fwd)"""
    lib = _lib.load()
    n, d = x.shape
    ld = padded_width(d, mode)
    u = torch.empty((n, ld), device=x.device, dtype=OP_TORCH_DTYPE[mode])   # separate storages: custom-op
    v = torch.empty((n, ld), device=x.device, dtype=OP_TORCH_DTYPE[mode])   # outputs must not alias each other
    with torch.cuda.device(x.device):
        lib.check(lib.plk_l2norm_pair_fwd(x.data_ptr(), y.data_ptr(), n, d, x.stride(0), u.data_ptr(),
                                          v.data_ptr(), mode, ld, stats[0].data_ptr(), stats[1].data_ptr(),
                                          stats[2].data_ptr(), stats[3].data_ptr(),
                                          zero_a.data_ptr() if zero_a is not None else None,
                                          zero_a.numel() if zero_a is not None else 0,
                                          zero_b.data_ptr() if zero_b is not None else None,
                                          zero_b.numel() if zero_b is not None else 0, _stream(x)),
                  "plk_l2norm_pair_fwd")
    return u, v


def infonce_fwd_local(u, v, mode, d, row_offset, bucket_size, logit_scale, rs=None, cs=None, dg=None,
                      sums_zeroed=False):
    """Fused similarity + sum-exp for the owned rows `u` against all rows `v`.
    -> (row_sumexp [n_rows], col_sumexp [n_cols] (partial over owned rows), diag [n_rows])"""
    lib = _lib.load()
    n_rows, n_cols = u.shape[0], v.shape[0]
    rs = torch.empty(n_rows, device=u.device, dtype=torch.float32) if rs is None else rs
    cs = torch.empty(n_cols, device=u.device, dtype=torch.float32) if cs is None else cs
    dg = torch.empty(n_rows, device=u.device, dtype=torch.float32) if dg is None else dg
    with torch.cuda.device(u.device):
        lib.check(lib.plk_infonce_fwd(u.data_ptr(), v.data_ptr(), mode, u.stride(0), n_rows, row_offset, n_cols,
                                      d, bucket_size, logit_scale.data_ptr(), rs.data_ptr(), cs.data_ptr(),
                                      dg.data_ptr(), 1 if sums_zeroed else 0, _stream(u)), "plk_infonce_fwd")
    return rs, cs, dg


def infonce_loss_local(rs, cs_own, dg, logit_scale, batch_global, loss_out=None):
    """-> (loss partial over the owned rows [], aux [2] = (sum of the owned diagonal logits, 0.0)).
    aux[1] is the zero-initialised accumulator the backward adds sum G*S into (`gs`)."""
    lib = _lib.load()
    loss = torch.empty((), device=rs.device, dtype=torch.float32) if loss_out is None else loss_out
    aux = torch.empty(2, device=rs.device, dtype=torch.float32)
    with torch.cuda.device(rs.device):
        lib.check(lib.plk_infonce_loss(rs.data_ptr(), cs_own.data_ptr(), dg.data_ptr(), logit_scale.data_ptr(),
                                       rs.shape[0], batch_global, loss.data_ptr(), aux.data_ptr(),
                                       aux[1:].data_ptr(), _stream(rs)), "plk_infonce_loss")
    return loss, aux


def infonce_grad_local(a, b, mode, d, row_offset, bucket_size, logit_scale, rs, cs, gs=None):
    """One direction of the recompute backward -> acc [parts, n_rows, d].
    `gs` (optional fp32 [1], already zeroed or holding a partial sum) is ADDED to."""
    lib = _lib.load()
    n_rows, n_cols = a.shape[0], b.shape[0]
    parts = lib.plk_infonce_grad_parts(mode, n_rows, n_cols, d, bucket_size)
    acc = torch.empty((parts, n_rows, d), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        lib.check(lib.plk_infonce_grad(a.data_ptr(), b.data_ptr(), mode, a.stride(0), n_rows, row_offset, n_cols,
                                       d, bucket_size, logit_scale.data_ptr(), rs.data_ptr(), cs.data_ptr(),
                                       acc.data_ptr(), gs.data_ptr() if gs is not None else None, _stream(a)),
                  "plk_infonce_grad")
    return acc


def infonce_grad_pair_local(a0, b0, a1, b1, mode, d, row_offset, bucket_size, logit_scale, rs0, cs0, rs1, cs1,
                            gs=None):
    """Both directions of the recompute backward in one launch -> (acc0, acc1), each
    [parts, n_rows, d]; direction 0 adds sum G*S into `gs`."""
    lib = _lib.load()
    n_rows, n_cols = a0.shape[0], b0.shape[0]
    parts = lib.plk_infonce_grad_pair_parts(mode, n_rows, n_cols, d, bucket_size)
    acc = torch.empty((2, parts, n_rows, d), device=a0.device, dtype=torch.float32)
    with torch.cuda.device(a0.device):
        lib.check(lib.plk_infonce_grad_pair(a0.data_ptr(), b0.data_ptr(), a1.data_ptr(), b1.data_ptr(), mode,
                                            a0.stride(0), n_rows, row_offset, n_cols, d, bucket_size,
                                            logit_scale.data_ptr(), rs0.data_ptr(), cs0.data_ptr(),
                                            rs1.data_ptr(), cs1.data_ptr(), acc[0].data_ptr(), acc[1].data_ptr(),
                                            gs.data_ptr() if gs is not None else None, _stream(a0)),
                  "plk_infonce_grad_pair")
    return acc[0], acc[1]


def infonce_grad_finish(acc, x, partner, inv_den_x, nrm_x, inv_den_p, dg, rs_own, cs_own, logit_scale, grad_out,
                        batch_global, out_dtype):
    lib = _lib.load()
    n, d = x.shape
    dx = torch.empty((n, d), device=x.device, dtype=out_dtype)
    with torch.cuda.device(x.device):
        lib.check(lib.plk_infonce_grad_finish(acc.data_ptr(), acc.shape[0], x.data_ptr(), partner.data_ptr(),
                                              PLK_F32, n, d, x.stride(0), inv_den_x.data_ptr(),
                                              nrm_x.data_ptr(), inv_den_p.data_ptr(), dg.data_ptr(),
                                              rs_own.data_ptr(), cs_own.data_ptr(), logit_scale.data_ptr(),
                                              grad_out.data_ptr(), batch_global, dx.data_ptr(),
                                              _DT[out_dtype], _stream(x)), "plk_infonce_grad_finish")
    return dx


def infonce_grad_finish_pair(acc_x, acc_y, x, y, stats_x, stats_y, dg, rs_own, cs_own, logit_scale, go_emb, go,
                             batch_global, gs, diag_sum, dls_out=None, xgpu=None, loss_partial=None):
    """Both gradient tails + d logit_scale in one launch (fp32 rows).  stats_x / stats_y are
    (1/den, |.|) pairs.  `gs` is consumed (reset to 0).  -> (dx, dy, dls)
    With `xgpu` (a dist.XGpuScalars) the kernel also sums (loss_partial, dls) over the ranks through
    peer memory; the global pair lands in xgpu.out2."""
    lib = _lib.load()
    n, d = x.shape
    dx = torch.empty((n, d), device=x.device, dtype=torch.float32)
    dy = torch.empty((n, d), device=x.device, dtype=torch.float32)
    dls = torch.empty((), device=x.device, dtype=torch.float32) if dls_out is None else dls_out
    common = (acc_x.data_ptr(), acc_y.data_ptr(), acc_x.shape[0], x.data_ptr(), y.data_ptr(), n, d, x.stride(0),
              stats_x[0].data_ptr(), stats_x[1].data_ptr(), stats_y[0].data_ptr(), stats_y[1].data_ptr(),
              dg.data_ptr(), rs_own.data_ptr(), cs_own.data_ptr(), logit_scale.data_ptr(), go_emb.data_ptr(),
              go.data_ptr(), batch_global, gs.data_ptr(), diag_sum.data_ptr(), dx.data_ptr(), dy.data_ptr(),
              dls.data_ptr())
    with torch.cuda.device(x.device):
        if xgpu is None:
            lib.check(lib.plk_infonce_grad_finish_pair(*common, _stream(x)), "plk_infonce_grad_finish_pair")
        else:
            lib.check(lib.plk_infonce_grad_finish_pair_xgpu(*common, loss_partial.data_ptr(), xgpu.peer_ptrs_dev,
                                                            xgpu.rank, xgpu.world, xgpu.epoch.data_ptr(),
                                                            xgpu.out2.data_ptr(), _stream(x)),
                      "plk_infonce_grad_finish_pair_xgpu")
    return dx, dy, dls


def infonce_dls(gs, diag_sum, grad_out, batch_global, out=None):
    lib = _lib.load()
    out = torch.empty((), device=gs.device, dtype=torch.float32) if out is None else out
    with torch.cuda.device(gs.device):
        lib.check(lib.plk_infonce_dls(gs.data_ptr(), diag_sum.data_ptr(), grad_out.data_ptr(), batch_global,
                                      out.data_ptr(), _stream(gs)), "plk_infonce_dls")
    return out


def _as_f32_rows(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.stride(-1) == 1 else t.contiguous()


# ---------------------------------------------------------------------------------------------
# single-GPU custom ops (the object bound to MultiModel.loss calls these)
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("plk::clip_loss_fwd", mutates_args=())
def clip_loss_fwd(image_emb: torch.Tensor, profile_emb: torch.Tensor, logit_scale: torch.Tensor,
                  buckets: int, mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                    torch.Tensor]:
    """-> (loss [], u [B,ld], v [B,ld], stats [7,B] = (1/den_x, |x|, 1/den_y, |y|, row sum-exp,
    col sum-exp, diagonal logit), aux [2] = (sum of diagonal logits, zeroed `gs` accumulator))"""
    _require_cuda(image_emb, profile_emb, logit_scale)
    B, d = image_emb.shape
    bs = B // buckets
    x, y = _as_f32_rows(image_emb), _as_f32_rows(profile_emb)
    ls = logit_scale.detach().float()
    stats = torch.empty((7, B), device=x.device, dtype=torch.float32)
    u, v = l2norm_pair(x, y, mode, stats, stats[4], stats[5])
    infonce_fwd_local(u, v, mode, d, 0, bs, ls, stats[4], stats[5], stats[6], sums_zeroed=True)
    loss, aux = infonce_loss_local(stats[4], stats[5], stats[6], ls, B)
    return loss, u, v, stats, aux


@clip_loss_fwd.register_fake
def _(image_emb, profile_emb, logit_scale, buckets, mode):
    B, d = image_emb.shape
    ld = padded_width(d, mode)
    odt = OP_TORCH_DTYPE[mode]
    f = image_emb.new_empty
    return (f((), dtype=torch.float32), f((B, ld), dtype=odt), f((B, ld), dtype=odt),
            f((7, B), dtype=torch.float32), f((2,), dtype=torch.float32))


@torch.library.custom_op("plk::clip_loss_bwd", mutates_args=())
def clip_loss_bwd(grad_out: torch.Tensor, image_emb: torch.Tensor, profile_emb: torch.Tensor,
                  logit_scale: torch.Tensor, u: torch.Tensor, v: torch.Tensor, stats: torch.Tensor,
                  aux: torch.Tensor, buckets: int, mode: int) -> tuple[torch.Tensor, torch.Tensor,
                                                                           torch.Tensor]:
    B, d = image_emb.shape
    bs = B // buckets
    x, y = _as_f32_rows(image_emb), _as_f32_rows(profile_emb)
    ls = logit_scale.detach().float()
    go = grad_out.detach().float().reshape(1).contiguous()
    idx, nx, idy, ny, rs, cs, dg = stats.unbind(0)
    # the kernels accumulate sum G*S into `gs` and reset it: work on a private copy so that this op mutates
    # none of its (saved) inputs -- what `mutates_args=()` promises to the tracer and to autograd's
    # version counters (a second backward over the same graph sees the forward's zero again)
    gs = aux[1:].clone()
    acc_x, acc_y = infonce_grad_pair_local(u, v, v, u, mode, d, 0, bs, ls, rs, cs, cs, rs, gs)
    dx, dy, dls = infonce_grad_finish_pair(acc_x, acc_y, x, y, stats[0:2], stats[2:4], dg, rs, cs, ls, go, go, B,
                                           gs, aux[0:1])
    return dx.to(image_emb.dtype), dy.to(profile_emb.dtype), dls.to(logit_scale.dtype)


@clip_loss_bwd.register_fake
def _(grad_out, image_emb, profile_emb, logit_scale, u, v, stats, aux, buckets, mode):
    return torch.empty_like(image_emb), torch.empty_like(profile_emb), torch.empty_like(logit_scale)


def _setup_ctx(ctx, inputs, output):
    image_emb, profile_emb, logit_scale, buckets, mode = inputs
    _, u, v, stats, aux = output
    ctx.save_for_backward(image_emb, profile_emb, logit_scale, u, v, stats, aux)
    ctx.buckets, ctx.mode = buckets, mode
    ctx.set_materialize_grads(False)   # no zero-filled gradients for the auxiliary outputs


def _backward(ctx, g_loss, *_unused):
    if g_loss is None:
        return None, None, None, None, None
    image_emb, profile_emb, logit_scale, u, v, stats, aux = ctx.saved_tensors
    dx, dy, dls = clip_loss_bwd(g_loss, image_emb, profile_emb, logit_scale, u, v, stats, aux,
                                ctx.buckets, ctx.mode)
    return dx, dy, dls, None, None


clip_loss_fwd.register_autograd(_backward, setup_context=_setup_ctx)


# ---------------------------------------------------------------------------------------------
# eager path: two native calls per step (plk_clip_loss_forward / plk_clip_loss_backward) behind a
# plain autograd.Function.  Same kernels as the custom ops above; the dispatcher / fake-tensor
# machinery of `torch.library` costs ~300 us of host time per step, 4x the GPU time of the step.
# ---------------------------------------------------------------------------------------------
_SIZES: dict = {}


def _clip_sizes(lib, mode: int, B: int, d: int, bs: int):
    key = (mode, B, d, bs)
    r = _SIZES.get(key)
    if r is None:
        r = (lib.plk_clip_loss_state_bytes(mode, B, d), lib.plk_clip_loss_workspace_bytes(mode, B, d, bs))
        _SIZES[key] = r
    return r


def _f32_rows(t: torch.Tensor) -> torch.Tensor:
    if t.dtype is not torch.float32:
        t = t.float()
    return t if t.stride(1) == 1 else t.contiguous()


def clip_loss_forward_state(x: torch.Tensor, y: torch.Tensor, ls: torch.Tensor, bs: int, mode: int,
                            batch_global: int | None = None, loss_out: torch.Tensor | None = None, xgpu=None,
                            partial_out: torch.Tensor | None = None):
    """x, y: fp32 [B, d] rows (unit inner stride, equal row stride); ls: fp32 scalar on the device.
    -> (loss [], state) where `state` is the opaque buffer plk_clip_loss_backward consumes.
    `batch_global` > B: the rows are one rank's share of a bucket-aligned global batch and `loss`
    is that rank's partial sum -- unless `xgpu` (dist.XGpuScalars) is given: then the loss kernel sums
    the partials over the ranks through peer memory, `loss` is the global loss and the partial goes
    to `partial_out`."""
    lib = _lib.load()
    B, d = x.shape
    dev = x.device
    state_bytes, _ = _clip_sizes(lib, mode, B, d, bs)
    state = torch.empty(state_bytes, device=dev, dtype=torch.uint8)
    loss = torch.empty((), device=dev, dtype=torch.float32) if loss_out is None else loss_out
    if y.stride(0) != x.stride(0):
        x, y = x.contiguous(), y.contiguous()
    idx = dev.index
    if torch.cuda.current_device() != idx:
        with torch.cuda.device(idx):
            return clip_loss_forward_state(x, y, ls, bs, mode, batch_global, loss, xgpu, partial_out)
    stream = torch._C._cuda_getCurrentRawStream(idx)
    common = (x.data_ptr(), y.data_ptr(), B, d, x.stride(0), mode, bs, B if batch_global is None else batch_global,
              ls.data_ptr(), state.data_ptr(), loss.data_ptr())
    if xgpu is None:
        lib.check(lib.plk_clip_loss_forward(*common, stream), "plk_clip_loss_forward")
    else:
        lib.check(lib.plk_clip_loss_forward_xgpu(*common, partial_out.data_ptr(), xgpu.peer_ptrs_dev, xgpu.rank,
                                                 xgpu.world, xgpu.epoch.data_ptr(), xgpu.out2.data_ptr(), stream),
                  "plk_clip_loss_forward_xgpu")
    return loss, state


def clip_loss_backward_state(go: torch.Tensor, x: torch.Tensor, y: torch.Tensor, ls: torch.Tensor, state, bs: int,
                             mode: int, batch_global: int | None = None, go_emb: torch.Tensor | None = None,
                             dls_out: torch.Tensor | None = None, xgpu=None, loss_partial=None,
                             emb_scale: float = 1.0):
    """-> (dx, dy, dls) fp32: dx, dy scaled by the scalar `go_emb` (default `go`) times the host float
    `emb_scale` (needs d % 128 == 0 when != 1), dls by `go` (fp32, on the device).  With `xgpu` (dist.XGpuScalars) the gradient-tail kernel also sums (loss_partial, dls)
    over the ranks through peer memory; the global pair lands in xgpu.out2."""
    lib = _lib.load()
    B, d = x.shape
    dev = x.device
    if y.stride(0) != x.stride(0):
        x, y = x.contiguous(), y.contiguous()
    idx = dev.index
    if torch.cuda.current_device() != idx:
        with torch.cuda.device(idx):
            return clip_loss_backward_state(go, x, y, ls, state, bs, mode, batch_global, go_emb, dls_out, xgpu,
                                            loss_partial, emb_scale)
    _, ws_bytes = _clip_sizes(lib, mode, B, d, bs)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    dx = torch.empty((B, d), device=dev, dtype=torch.float32)
    dy = torch.empty((B, d), device=dev, dtype=torch.float32)
    dls = torch.empty((), device=dev, dtype=torch.float32) if dls_out is None else dls_out
    common = (go.data_ptr(), (go if go_emb is None else go_emb).data_ptr(), float(emb_scale), x.data_ptr(),
              y.data_ptr(), B, d,
              x.stride(0), mode, bs, B if batch_global is None else batch_global, ls.data_ptr(), state.data_ptr(),
              ws.data_ptr(), dx.data_ptr(), dy.data_ptr(), dls.data_ptr())
    stream = torch._C._cuda_getCurrentRawStream(idx)
    if xgpu is None:
        lib.check(lib.plk_clip_loss_backward(*common, stream), "plk_clip_loss_backward")
    else:
        lib.check(lib.plk_clip_loss_backward_xgpu(*common, loss_partial.data_ptr(), xgpu.peer_ptrs_dev, xgpu.rank,
                                                  xgpu.world, xgpu.epoch.data_ptr(), xgpu.out2.data_ptr(), stream),
                  "plk_clip_loss_backward_xgpu")
    return dx, dy, dls


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_emb, profile_emb, logit_scale, buckets, mode):
        _require_cuda(image_emb, profile_emb, logit_scale)
        bs = image_emb.shape[0] // buckets
        x, y = _f32_rows(image_emb), _f32_rows(profile_emb)
        ls = logit_scale if logit_scale.dtype is torch.float32 else logit_scale.float()
        loss, state = clip_loss_forward_state(x, y, ls, bs, mode)
        ctx.save_for_backward(x, y, ls, state)
        ctx.meta = (bs, mode, image_emb.dtype, profile_emb.dtype, logit_scale.dtype)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss):
        x, y, ls, state = ctx.saved_tensors
        bs, mode, dtx, dty, dtl = ctx.meta
        go = g_loss if g_loss.dtype is torch.float32 else g_loss.float()
        dx, dy, dls = clip_loss_backward_state(go, x, y, ls, state, bs, mode)
        if dtx is not torch.float32:
            dx = dx.to(dtx)
        if dty is not torch.float32:
            dy = dy.to(dty)
        if dtl is not torch.float32:
            dls = dls.to(dtl)
        return dx, dy, dls, None, None


# ---------------------------------------------------------------------------------------------
# SigLIP: same state layout, composite calls plk_siglip_loss_forward/backward
# ---------------------------------------------------------------------------------------------
def siglip_loss_forward_state(x: torch.Tensor, y: torch.Tensor, ls: torch.Tensor, bias: torch.Tensor, bs: int,
                              mode: int):
    """x, y: fp32 [B, d] rows; ls, bias: fp32 scalars on the device.  -> (loss [], state)"""
    lib = _lib.load()
    B, d = x.shape
    dev = x.device
    state_bytes, _ = _clip_sizes(lib, mode, B, d, bs)
    state = torch.empty(state_bytes, device=dev, dtype=torch.uint8)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    if y.stride(0) != x.stride(0):
        x, y = x.contiguous(), y.contiguous()
    with torch.cuda.device(dev):
        lib.check(lib.plk_siglip_loss_forward(x.data_ptr(), y.data_ptr(), B, d, x.stride(0), mode, bs, ls.data_ptr(),
                                              bias.data_ptr(), state.data_ptr(), loss.data_ptr(), _stream(x)),
                  "plk_siglip_loss_forward")
    return loss, state


def siglip_loss_backward_state(go: torch.Tensor, x: torch.Tensor, y: torch.Tensor, ls: torch.Tensor,
                               bias: torch.Tensor, state, bs: int, mode: int):
    """-> (dx, dy, dls, dbias) fp32, scaled by the scalar `go` (fp32, on the device)."""
    lib = _lib.load()
    B, d = x.shape
    dev = x.device
    if y.stride(0) != x.stride(0):
        x, y = x.contiguous(), y.contiguous()
    _, ws_bytes = _clip_sizes(lib, mode, B, d, bs)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    dx = torch.empty((B, d), device=dev, dtype=torch.float32)
    dy = torch.empty((B, d), device=dev, dtype=torch.float32)
    dls = torch.empty((), device=dev, dtype=torch.float32)
    dbias = torch.empty((), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        lib.check(lib.plk_siglip_loss_backward(go.data_ptr(), x.data_ptr(), y.data_ptr(), B, d, x.stride(0), mode, bs,
                                               ls.data_ptr(), bias.data_ptr(), state.data_ptr(), ws.data_ptr(),
                                               dx.data_ptr(), dy.data_ptr(), dls.data_ptr(), dbias.data_ptr(),
                                               _stream(x)), "plk_siglip_loss_backward")
    return dx, dy, dls, dbias


class _SigLipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_emb, profile_emb, logit_scale, bias, buckets, mode):
        _require_cuda(image_emb, profile_emb, logit_scale, bias)
        bs = image_emb.shape[0] // buckets
        x, y = _f32_rows(image_emb), _f32_rows(profile_emb)
        ls = logit_scale if logit_scale.dtype is torch.float32 else logit_scale.float()
        b = bias if bias.dtype is torch.float32 else bias.float()
        loss, state = siglip_loss_forward_state(x, y, ls, b, bs, mode)
        ctx.save_for_backward(x, y, ls, b, state)
        ctx.meta = (bs, mode, image_emb.dtype, profile_emb.dtype, logit_scale.dtype, bias.dtype)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss):
        x, y, ls, b, state = ctx.saved_tensors
        bs, mode, dtx, dty, dtl, dtb = ctx.meta
        go = g_loss if g_loss.dtype is torch.float32 else g_loss.float()
        dx, dy, dls, dbias = siglip_loss_backward_state(go, x, y, ls, b, state, bs, mode)
        return dx.to(dtx), dy.to(dty), dls.to(dtl), dbias.to(dtb), None, None


def siglip_loss(image_emb, profile_emb, logit_scale, bias, buckets: int = 1, mode: int = PLK_BF16) -> torch.Tensor:
    """Pairwise-sigmoid loss of reference src/coordination.py:67-95 on the fused CUDA path."""
    return _SigLipLossFn.apply(image_emb, profile_emb, logit_scale, bias, int(buckets), int(mode))


# ---------------------------------------------------------------------------------------------
# N1: projection Linears fused with the normalisation that opens the loss
# ---------------------------------------------------------------------------------------------
def _pad64_cast(t: torch.Tensor, dtype) -> torch.Tensor:
    """[r, f] -> 16-bit [r, ceil(f/64)*64], zero padded (one cast pass; the TMA boxes are 64 elements wide)."""
    r, f = t.shape
    fp = (f + 63) // 64 * 64
    if fp == f:
        return t.detach().to(dtype).contiguous()
    out = torch.zeros((r, fp), device=t.device, dtype=dtype)
    out[:, :f] = t.detach()
    return out


def project_normalise(feat: torch.Tensor, weight: torch.Tensor, mode: int):
    """emb = feat @ weight.T (nn.Linear without bias) and u = emb / max(||emb||, eps) in one tcgen05 kernel
    (plk_project_normalise).  -> (u [n, ld] operand dtype, emb [n, d] fp32, inv_den [n], nrm [n]).
    fp32 mode: the projection is a plain fp32 matmul followed by plk_l2norm_fwd."""
    lib = _lib.load()
    n, f = feat.shape
    d = weight.shape[0]
    if mode == PLK_F32:
        emb = torch.matmul(feat.detach().float(), weight.detach().float().t()).contiguous()
        u, inv_den, nrm, _ = l2norm(emb, mode)
        return u, emb, inv_den, nrm
    odt = OP_TORCH_DTYPE[mode]
    x16, w16 = _pad64_cast(feat, odt), _pad64_cast(weight, odt)
    return _project_normalise16(x16, w16, n, f, d, mode)


def _project_normalise16(x16: torch.Tensor, w16: torch.Tensor, n: int, f: int, d: int, mode: int):
    """plk_project_normalise on operands that are already 16-bit and zero-padded to a multiple of 64 columns."""
    lib = _lib.load()
    odt = OP_TORCH_DTYPE[mode]
    feat = x16
    ld = padded_width(d, mode)
    u = torch.empty((n, ld), device=feat.device, dtype=odt)
    emb = torch.empty((n, d), device=feat.device, dtype=torch.float32)
    inv_den = torch.empty(n, device=feat.device, dtype=torch.float32)
    nrm = torch.empty(n, device=feat.device, dtype=torch.float32)
    with torch.cuda.device(feat.device):
        lib.check(lib.plk_project_normalise(x16.data_ptr(), x16.stride(0), w16.data_ptr(), w16.stride(0), mode, n, f, d,
                                            u.data_ptr(), ld, emb.data_ptr(), inv_den.data_ptr(), nrm.data_ptr(),
                                            _stream(feat)), "plk_project_normalise")
    return u, emb, inv_den, nrm


class _ProjectedClipLossFn(torch.autograd.Function):
    """loss(image_feat @ Wi^T, profile_feat @ Wp^T): fused projection + normalisation forward, the loss kernels
    on the resulting operands, and in the backward dW = d_emb^T feat, d_feat = d_emb W from the embedding
    gradients the gradient tail produces (two library GEMMs per modality: off the similarity path)."""

    @staticmethod
    def forward(ctx, image_feat, profile_feat, w_i, w_p, logit_scale, buckets, mode):
        _require_cuda(image_feat, profile_feat, w_i, w_p, logit_scale)
        B = image_feat.shape[0]
        d = w_i.shape[0]
        bs = B // buckets
        ls = logit_scale.detach().float()
        if mode == PLK_F32:
            u, x, idx, nx = project_normalise(image_feat, w_i, mode)
            v, y, idy, ny = project_normalise(profile_feat, w_p, mode)
            ops16 = (image_feat, profile_feat, w_i, w_p)
        else:   # the 16-bit copies of features and weights feed the forward kernel AND the backward GEMMs
            odt = OP_TORCH_DTYPE[mode]
            ops16 = tuple(_pad64_cast(t, odt) for t in (image_feat, profile_feat, w_i, w_p))
            u, x, idx, nx = _project_normalise16(ops16[0], ops16[2], B, image_feat.shape[1], d, mode)
            v, y, idy, ny = _project_normalise16(ops16[1], ops16[3], B, profile_feat.shape[1], d, mode)
        sums = torch.zeros((2, B), device=x.device, dtype=torch.float32)
        dg = torch.empty(B, device=x.device, dtype=torch.float32)
        infonce_fwd_local(u, v, mode, d, 0, bs, ls, sums[0], sums[1], dg, sums_zeroed=True)
        loss, aux = infonce_loss_local(sums[0], sums[1], dg, ls, B)
        ctx.save_for_backward(*ops16, ls, u, v, x, y, idx, nx, idy, ny, sums, dg, aux)
        ctx.meta = (bs, mode, logit_scale.dtype, image_feat.shape[1], profile_feat.shape[1], image_feat.dtype,
                    profile_feat.dtype, w_i.dtype, w_p.dtype)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss):
        f16_i, f16_p, w16_i, w16_p, ls, u, v, x, y, idx, nx, idy, ny, sums, dg, aux = ctx.saved_tensors
        bs, mode, dtl, f_i, f_p, dt_fi, dt_fp, dt_wi, dt_wp = ctx.meta
        B, d = x.shape
        go = g_loss.detach().float().reshape(1).contiguous()
        gs = aux[1:].clone()
        acc_x, acc_y = infonce_grad_pair_local(u, v, v, u, mode, d, 0, bs, ls, sums[0], sums[1], sums[1], sums[0], gs)
        dx, dy, dls = infonce_grad_finish_pair(acc_x, acc_y, x, y, (idx, nx), (idy, ny), dg, sums[0], sums[1], ls, go,
                                               go, B, gs, aux[0:1])
        out = []
        for demb, feat, w, f, dtf, dtw in ((dx, f16_i, w16_i, f_i, dt_fi, dt_wi), (dy, f16_p, w16_p, f_p, dt_fp, dt_wp)):
            d16 = demb if mode == PLK_F32 else demb.to(OP_TORCH_DTYPE[mode])
            out.append((_mm_mode(d16, w[:, :f], mode).to(dtf),                              # d feat  [B, f]
                        _mm_mode(d16.t(), feat[:, :f], mode).to(dtw)))                      # d W     [d, f]
        return out[0][0], out[1][0], out[0][1], out[1][1], dls.to(dtl), None, None


def _mm_mode(a: torch.Tensor, b: torch.Tensor, mode: int) -> torch.Tensor:
    """a @ b for the projection backward (library GEMM, off the similarity path).  fp32 mode: fp32 operands.
    16-bit modes: operands rounded to the mode's 16-bit type, fp32 accumulation AND fp32 result
    (`aten::mm.dtype`) -- what the reference's '16-mixed' autocast does to `nn.Linear`'s backward, minus its
    rounding of the result.  (An fp32 GEMM here cost more than the whole loss step: 90 us at B = 4096, f = 1280.)"""
    if mode == PLK_F32:
        return torch.matmul(a.detach().float(), b.detach().float())
    odt = OP_TORCH_DTYPE[mode]
    a16, b16 = a.detach().to(odt), b.detach().to(odt)    # no-ops for operands that are already 16-bit
    try:
        return torch.mm(a16, b16, out_dtype=torch.float32)
    except (TypeError, NotImplementedError, RuntimeError):
        return torch.mm(a16, b16).float()


def clip_loss_projected(image_feat, profile_feat, image_weight, profile_weight, logit_scale, buckets: int = 1,
                        mode: int = PLK_BF16) -> torch.Tensor:
    """Symmetric InfoNCE of the PROJECTED features: reference src/model.py:80-83 (the two bias-free
    `nn.Linear`) followed by src/coordination.py:26-47, with projection + normalisation in one kernel."""
    return _ProjectedClipLossFn.apply(image_feat, profile_feat, image_weight, profile_weight, logit_scale,
                                      int(buckets), int(mode))


def clip_loss(image_emb, profile_emb, logit_scale, buckets: int = 1, mode: int = PLK_BF16) -> torch.Tensor:
    """Symmetric InfoNCE of reference src/coordination.py:26-47 on the fused CUDA path.  Under
    torch.compile the registered custom ops are used (traceable); in eager mode the lean path."""
    if torch.compiler.is_compiling():
        return clip_loss_fwd(image_emb, profile_emb, logit_scale, int(buckets), int(mode))[0]
    return _ClipLossFn.apply(image_emb, profile_emb, logit_scale, int(buckets), int(mode))
