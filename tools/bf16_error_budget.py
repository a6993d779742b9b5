"""CPU emulation of the 16-bit tensor-core loss path: which rounding dominates the gradient error?

Rounds (a) the normalised operands u, v and (b) the recomputed softmax weights G to bf16 / fp16
independently and reports the two gradient error metrics of tests/test_gpu_loss.py against the fp64
closed form.  Everything else (sums, exp, tail) is fp64 here, so the numbers are the floor the
kernels can reach with that operand format.  No GPU needed.
"""
import sys, os, itertools
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import infonce as oinf


def rnd(t, fmt):
    if fmt == "f64":
        return t
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    return t.float().to(dt).double()


def emulate(img, pro, ls, buckets, op_fmt, g_fmt, v2_fmt=None):
    x, y = torch.tensor(img).double(), torch.tensor(pro).double()
    B, d = x.shape
    bs = B // buckets
    nx, ny = x.norm(dim=1).clamp_min(1e-12), y.norm(dim=1).clamp_min(1e-12)
    u, v = x / nx[:, None], y / ny[:, None]
    ub, vb = rnd(u, op_fmt), rnd(v, op_fmt)
    s = float(np.exp(ls))
    S = s * (ub @ vb.T)
    mask = (torch.arange(B)[:, None] // bs) == (torch.arange(B)[None, :] // bs)
    E = torch.where(mask, torch.exp(S - s + 64.0), torch.zeros_like(S))   # kShiftK (csrc/common.cuh)
    R, C = E.sum(1), E.sum(0)
    G = E * (1 / R[:, None] + 1 / C[None, :])
    dg = torch.diagonal(S).clone()
    Gd = torch.diagonal(G).clone()
    G.fill_diagonal_(0)
    Gb = rnd(G, g_fmt)
    v2 = vb if v2_fmt is None else rnd(v, v2_fmt)
    u2 = ub if v2_fmt is None else rnd(u, v2_fmt)
    coef = s / (2 * B)
    dU = coef * (Gb @ v2 + (Gd - 2)[:, None] * v)
    dV = coef * (Gb.T @ u2 + (Gd - 2)[:, None] * u)
    dx = (dU - u * (u * dU).sum(1, keepdim=True)) / nx[:, None]
    dy = (dV - v * (v * dV).sum(1, keepdim=True)) / ny[:, None]
    return dx.numpy(), dy.numpy()


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def case(B, d, buckets, ls, seed=None):
    r = np.random.default_rng(B + d if seed is None else seed)
    cent = r.standard_normal((27, d))
    lab = r.integers(0, 27, B)
    z = r.standard_normal((B, d))
    img = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    pro = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    return img, pro


if __name__ == "__main__":
    for (B, d, bk, ls) in [(1024, 384, 8, 2.659), (2048, 256, 1, 2.659), (1024, 256, 1, 1.0), (2048, 256, 1, 3.7)]:
        img, pro = case(B, d, bk, ls)
        ref = oinf.clip_loss_closed_form(img, pro, ls, bk)
        print(f"B={B} d={d} buckets={bk} ls={ls}")
        for op_fmt, g_fmt in itertools.product(("f64", "bf16", "fp16"), ("f64", "bf16", "fp16")):
            dx, dy = emulate(img, pro, ls, bk, op_fmt, g_fmt)
            print(f"  operands {op_fmt:5s} G {g_fmt:5s}: max {rel(dx, ref['d_image']):.2e} {rel(dy, ref['d_profile']):.2e}"
                  f"  l2 {rel_l2(dx, ref['d_image']):.2e} {rel_l2(dy, ref['d_profile']):.2e}")
