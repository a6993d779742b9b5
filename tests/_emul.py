"""CPU stand-ins for the libplk entry points, used ONLY by the gloo (world_size 2) tests of the
host-side sharding logic in multimodal_plankton_recognition_b200/dist.py.  Each function follows
the contract stated in include/plk.h for the entry point it replaces (fp64 arithmetic)."""
import numpy as np
import torch

from oracle import ann as oann


def _mask(n_rows, n_cols, off, bs):
    gi = torch.arange(n_rows)[:, None] + off
    return (gi // bs) == (torch.arange(n_cols)[None, :] // bs)


def l2norm(x, mode, normalise=True, inv_den=None, nrm=None, want_sqn=False):
    n = x.double().norm(dim=1)
    den = n.clamp_min(1e-12)
    u = (x.double() / den[:, None]).float() if normalise else x.clone()
    inv = (1.0 / den).float()
    if inv_den is not None:
        inv_den.copy_(inv)
    if nrm is not None:
        nrm.copy_(n.float())
    return u, inv if inv_den is None else inv_den, n.float() if nrm is None else nrm, \
        (u.double() ** 2).sum(1).float() if want_sqn else None


def l2norm_pair(x, y, mode, stats, zero_a=None, zero_b=None):
    u, _, _, _ = l2norm(x, mode, True, stats[0], stats[1])
    v, _, _, _ = l2norm(y, mode, True, stats[2], stats[3])
    for z in (zero_a, zero_b):
        if z is not None:
            z.zero_()
    return u, v


def _E(a, b, off, bs, ls):
    s = float(torch.exp(ls.double()))
    S = s * (a.double() @ b.double().T)
    m = _mask(a.shape[0], b.shape[0], off, bs)
    return torch.where(m, torch.exp(S - s + 64.0), torch.zeros_like(S)), S, s   # kShiftK, csrc/common.cuh


def infonce_fwd_local(u, v, mode, d, row_offset, bucket_size, ls, rs=None, cs=None, dg=None, sums_zeroed=False):
    E, S, _ = _E(u, v, row_offset, bucket_size, ls)
    n = u.shape[0]
    out = (E.sum(1).float(), E.sum(0).float(), S[torch.arange(n), torch.arange(n) + row_offset].float())
    for dst, src in zip((rs, cs, dg), out):
        if dst is not None:
            dst.copy_(src)
    return tuple(o if dst is None else dst for dst, o in zip((rs, cs, dg), out))


def infonce_loss_local(rs, cs_own, dg, ls, batch_global, loss_out=None):
    s = float(torch.exp(ls.double()))
    loss = ((2 * (s - 64.0) + rs.double().log() + cs_own.double().log() - 2 * dg.double()).sum() / (2 * batch_global)).float()
    if loss_out is not None:
        loss_out.copy_(loss)
        loss = loss_out
    return loss, torch.tensor([float(dg.double().sum()), 0.0])


def infonce_grad_pair_local(a0, b0, a1, b1, mode, d, off, bs, ls, rs0, cs0, rs1, cs1, gs=None):
    out = []
    for k, (a, b, rs, cs) in enumerate(((a0, b0, rs0, cs0), (a1, b1, rs1, cs1))):
        E, S, _ = _E(a, b, off, bs, ls)
        G = E * (1.0 / rs.double()[:, None] + 1.0 / cs.double()[None, :])
        if k == 0 and gs is not None:
            gs += float((G * S).sum())
        n = a.shape[0]
        G[torch.arange(n), torch.arange(n) + off] = 0          # j == i term is grad_finish's job
        out.append((G @ b.double()).float()[None])
    return out[0], out[1]


def infonce_grad_finish(acc, x, partner, inv_den_x, nrm_x, inv_den_p, dg, rs_own, cs_own, ls, go, batch_global,
                        out_dtype):
    s = float(torch.exp(ls.double()))
    coef = float(go) * s / (2 * batch_global)
    dterm = torch.exp(dg.double() - s + 64.0) * (1 / rs_own.double() + 1 / cs_own.double()) - 2
    p = partner.double() * inv_den_p.double()[:, None]
    dU = coef * (acc.double().sum(0) + dterm[:, None] * p)
    u = x.double() * inv_den_x.double()[:, None]
    dot = (u * dU).sum(1, keepdim=True)
    dot = torch.where((nrm_x > 1e-12)[:, None], dot, torch.zeros_like(dot))
    return ((dU - u * dot) * inv_den_x.double()[:, None]).to(out_dtype)


def infonce_grad_finish_pair(acc_x, acc_y, x, y, stats_x, stats_y, dg, rs_own, cs_own, ls, go_emb, go, batch_global,
                             gs, diag_sum, dls_out=None, xgpu=None, loss_partial=None):
    assert xgpu is None, "the fused cross-GPU exchange is CUDA-only"
    dx = infonce_grad_finish(acc_x, x, y, stats_x[0], stats_x[1], stats_y[0], dg, rs_own, cs_own, ls, go_emb,
                             batch_global, torch.float32)
    dy = infonce_grad_finish(acc_y, y, x, stats_y[0], stats_y[1], stats_x[0], dg, rs_own, cs_own, ls, go_emb,
                             batch_global, torch.float32)
    dls = infonce_dls(gs, diag_sum, go, batch_global, dls_out)
    gs.zero_()
    return dx, dy, dls


def infonce_dls(gs, diag_sum, go, batch_global, out=None):
    val = (go.double() / (2 * batch_global) * (gs.double() - 2 * diag_sum.double())).float().reshape(())
    if out is not None:
        out.copy_(val)
        return out
    return val


def clip_loss_forward_state(x, y, ls, bs, mode, batch_global=None, loss_out=None, xgpu=None, partial_out=None):
    """plk_clip_loss_forward: the complete local problem with the global 1/(2B)."""
    assert xgpu is None, "the fused cross-GPU exchange is CUDA-only"
    n, d = x.shape
    stats = torch.empty((4, n))
    u, v = l2norm_pair(x, y, mode, stats)
    rs, cs, dg = infonce_fwd_local(u, v, mode, d, 0, bs, ls)
    loss, aux = infonce_loss_local(rs, cs, dg, ls, n if batch_global is None else batch_global, loss_out)
    return loss, (u, v, stats, rs, cs, dg, aux)


def clip_loss_backward_state(go, x, y, ls, state, bs, mode, batch_global=None, go_emb=None, dls_out=None, xgpu=None,
                             loss_partial=None, emb_scale=1.0):
    """plk_clip_loss_backward."""
    u, v, stats, rs, cs, dg, aux = state
    n, d = x.shape
    Bg = n if batch_global is None else batch_global
    gs = aux[1:]
    acc_x, acc_y = infonce_grad_pair_local(u, v, v, u, mode, d, 0, bs, ls, rs, cs, cs, rs, gs)
    return infonce_grad_finish_pair(acc_x, acc_y, x, y, stats[0:2], stats[2:4], dg, rs, cs, ls,
                                    (go if go_emb is None else go_emb) * emb_scale, go, Bg, gs, aux[0:1], dls_out, xgpu,
                                    loss_partial)


class CpuExactIndex:
    """Stand-in for ann.GpuExactIndex over the oracle's exact index (global indices via offset)."""

    def __init__(self, X, precision="bf16", device=None, gallery_offset=0, slack=6):
        self.idx = oann.ExactIndex(np.asarray(X))
        self.n, self.d = np.asarray(X).shape
        self.device = torch.device("cpu")
        self.gallery_offset = gallery_offset

    def search_device(self, q32, k):
        i, d_ = self.idx.query(q32.numpy(), k=k)
        return torch.from_numpy(i + self.gallery_offset).int(), torch.from_numpy(d_)


def topk_merge_device(cand_i, cand_d, k):
    nq, m = cand_i.shape
    oi = torch.full((nq, k), -1, dtype=torch.int32)
    od = torch.full((nq, k), float("inf"))
    for q in range(nq):
        pairs = sorted((float(cand_d[q, t]), int(cand_i[q, t])) for t in range(m) if int(cand_i[q, t]) >= 0)[:k]
        for t, (dd, ii) in enumerate(pairs):
            oi[q, t], od[q, t] = ii, dd
    return oi, od


def knn_vote_device(idx, dist, labels):
    w = oann.inverse_distance_weights(dist.numpy())
    return torch.from_numpy(oann.weighted_vote(labels.numpy()[idx.numpy()], w))
