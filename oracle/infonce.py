"""Oracle: symmetric CLIP-style InfoNCE coordination loss, forward + backward.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Two restatements:

``clip_loss_closed_form``  numpy, any float dtype (fp64 by default).  Follows
    the arithmetic of ``CLIPLoss.forward`` (reference src/coordination.py:26-47)
    and of the autograd graph it builds, written out in closed form so that the
    gradients w.r.t. the raw embeddings and ``logit_scale`` are explicit.
``clip_loss_materialised`` torch-CPU, fp32: the same op sequence the reference
    executes (normalise, bucket view, batched matmul, two cross-entropies per
    bucket, autograd backward) -- this is the "port" that ``bench.py`` times as
    the CPU baseline, because ``/root/reference`` does not travel to the GPU box.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12  # F.normalize default eps, reference src/coordination.py:33-34


def _normalise(x: np.ndarray):
    """x / max(||x||_2, eps) per row -- reference src/coordination.py:33-34."""
    nrm = np.sqrt((x * x).sum(axis=1, keepdims=True))
    den = np.maximum(nrm, EPS)
    return x / den, nrm, den


def _normalise_bwd(g: np.ndarray, x: np.ndarray, nrm: np.ndarray, den: np.ndarray):
    """Backward of ``x / clamp_min(||x||, eps)``.

    Above the clamp: dx = (g - u (u.g)) / ||x||.  Below it the denominator is the
    constant eps, so dx = g / eps (no projection term).
    """
    u = x / den
    proj = (u * g).sum(axis=1, keepdims=True)
    above = nrm > EPS
    return np.where(above, (g - u * proj) / den, g / den)


def clip_loss_closed_form(image_emb, profile_emb, logit_scale=1.0, buckets=1,
                          dtype=np.float64, grad_out=1.0):
    """Loss and gradients of the reference ``CLIPLoss``.

    reference src/coordination.py:29-31  -> divisibility assert, bucket size
    reference src/coordination.py:33-34  -> L2 normalisation
    reference src/coordination.py:36-38  -> [buckets, bs, d] view, U V^T * exp(ls)
    reference src/coordination.py:40-45  -> CE over rows and over columns with the
                                            diagonal as target, mean over buckets, /2

    Returns dict(loss, d_image, d_profile, d_logit_scale, row_lse, col_lse, diag).
    """
    x = np.asarray(image_emb, dtype=dtype)
    y = np.asarray(profile_emb, dtype=dtype)
    B, d = x.shape
    assert B % buckets == 0, "Batch size must be divisible by number of buckets!"
    bs = B // buckets
    s = dtype(np.exp(dtype(logit_scale)))

    u, nx, dx_den = _normalise(x)
    v, ny, dy_den = _normalise(y)

    loss = dtype(0)
    dU = np.zeros_like(u)
    dV = np.zeros_like(v)
    dls = dtype(0)
    row_lse = np.zeros(B, dtype=dtype)
    col_lse = np.zeros(B, dtype=dtype)
    diag = np.zeros(B, dtype=dtype)
    for b in range(buckets):
        sl = slice(b * bs, (b + 1) * bs)
        S = (u[sl] @ v[sl].T) * s                      # logits of this bucket
        m_r = S.max(axis=1, keepdims=True)
        m_c = S.max(axis=0, keepdims=True)
        Er = np.exp(S - m_r)
        Ec = np.exp(S - m_c)
        lr = np.log(Er.sum(axis=1)) + m_r[:, 0]        # row log-sum-exp
        lc = np.log(Ec.sum(axis=0)) + m_c[0, :]        # column log-sum-exp
        dg = np.diagonal(S)
        row_lse[sl], col_lse[sl], diag[sl] = lr, lc, dg
        # (CE_rows + CE_cols)/2 averaged over buckets == sum / (2 B)
        loss += (lr.sum() + lc.sum() - 2 * dg.sum())
        # d loss / d S  (softmax over rows + softmax over columns - 2 I) / (2B)
        G = Er / Er.sum(axis=1, keepdims=True) + Ec / Ec.sum(axis=0, keepdims=True)
        G[np.arange(bs), np.arange(bs)] -= 2
        G = G * (grad_out / (2 * B))
        dU[sl] = s * (G @ v[sl])
        dV[sl] = s * (G.T @ u[sl])
        dls += (G * S).sum()                           # d/d ls of S = S
    loss = loss / (2 * B)
    return dict(
        loss=loss,
        d_image=_normalise_bwd(dU, x, nx, dx_den),
        d_profile=_normalise_bwd(dV, y, ny, dy_den),
        d_logit_scale=dls,
        row_lse=row_lse, col_lse=col_lse, diag=diag,
    )


def clip_loss_materialised(image_emb, profile_emb, logit_scale, buckets=1):
    """torch-CPU port executing the reference's op sequence (B x bs logits held
    in memory, python loop over buckets).  Inputs are torch tensors; returns the
    0-dim loss tensor with an autograd graph, exactly like the reference module.
    Used by bench.py as the timed CPU baseline ("port").

    reference src/coordination.py:29-45.
    """
    import torch
    import torch.nn.functional as F

    n = image_emb.size(0)
    assert n % buckets == 0, "Batch size must be divisible by number of buckets!"
    bs = n // buckets
    u = F.normalize(image_emb).view(buckets, bs, -1)
    v = F.normalize(profile_emb).view(buckets, bs, -1)
    logits = torch.bmm(u, v.transpose(1, 2)) * logit_scale.exp()
    target = torch.arange(bs, device=logits.device)
    fwd = torch.stack([F.cross_entropy(lg, target) for lg in logits]).mean()
    rev = torch.stack([F.cross_entropy(lg.T, target) for lg in logits]).mean()
    return (fwd + rev) / 2


def clip_loss_grads_rounded_operands(image_emb, profile_emb, logit_scale=1.0, buckets=1, op_fmt="bf16",
                                     g_fmt="bf16"):
    """Floor of a 16-bit tensor-core evaluation of the same gradients: the closed form above with (a) the
    normalised operands of the similarity GEMM and (b) the softmax weights G (diagonal excluded) rounded to
    `op_fmt` / `g_fmt` ("bf16", "fp16" or "f64" = not rounded); sums, exponentials, the diagonal term and
    the normalisation backward stay in fp64.  What it returns is the error a kernel with that operand
    format cannot go below -- tests use it to show that the bf16-mode deviation is the format's, not the
    kernel's (tests/test_oracle_golden.py, tools/bf16_error_budget.py).  -> (d_image, d_profile)"""
    import torch

    def rnd(t, fmt):
        if fmt == "f64":
            return t
        return t.float().to({"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]).double()

    x, y = torch.tensor(np.asarray(image_emb)).double(), torch.tensor(np.asarray(profile_emb)).double()
    B = x.shape[0]
    bs = B // buckets
    nx, ny = x.norm(dim=1).clamp_min(EPS), y.norm(dim=1).clamp_min(EPS)
    u, v = x / nx[:, None], y / ny[:, None]
    ub, vb = rnd(u, op_fmt), rnd(v, op_fmt)
    s = float(np.exp(logit_scale))
    S = s * (ub @ vb.T)
    mask = (torch.arange(B)[:, None] // bs) == (torch.arange(B)[None, :] // bs)
    S = torch.where(mask, S, torch.full_like(S, -float("inf")))
    G = torch.softmax(S, dim=1) + torch.softmax(S, dim=0)
    Gd = torch.diagonal(G).clone()
    G.fill_diagonal_(0)
    Gb = rnd(G, g_fmt)
    coef = s / (2 * B)
    dU = coef * (Gb @ vb + (Gd - 2)[:, None] * v)
    dV = coef * (Gb.T @ ub + (Gd - 2)[:, None] * u)
    dx = (dU - u * (u * dU).sum(1, keepdim=True)) / nx[:, None]
    dy = (dV - v * (v * dV).sum(1, keepdim=True)) / ny[:, None]
    return dx.numpy(), dy.numpy()

