"""Soak of the paired-CTA backward (infonce_grad_tc5) and the forward: many shapes / seeds back to back, each
checked against an fp64 evaluation of the same 16-bit operands on the GPU (max-abs / max-abs), plus bitwise
run-to-run determinism of the partial slabs.  A lost or early barrier hand-off between the two CTAs of a pair
shows up here as a wrong slab long before it shows up in a tolerance test.
usage: python tools/soak_pairs.py [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 120
mode = ops.MODES["bf16"]
g = torch.Generator().manual_seed(7)
worst = 0.0
for it in range(iters):
    d = (128, 256)[int(torch.randint(0, 2, (1,), generator=g))]
    B = int(torch.randint(2, 40, (1,), generator=g)) * 128 - int(torch.randint(0, 2, (1,), generator=g)) * 37
    img, pro, _ = synth.pairs(B, d, 100 + it, "cuda")
    ls = torch.full((), 0.5 + 0.02 * it, device="cuda")
    u, *_ = ops.l2norm(img, mode)
    v, *_ = ops.l2norm(pro, mode)
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, B, ls)
    gs = torch.zeros(1, device="cuda")
    ax, ay = ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, B, ls, rs, cs, cs, rs, gs)
    ax2, ay2 = ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, B, ls, rs, cs, cs, rs, None)
    torch.cuda.synchronize()
    assert torch.equal(ax, ax2) and torch.equal(ay, ay2), f"non-deterministic slabs at it={it} B={B} d={d}"
    uf, vf = u.double()[:, :d], v.double()[:, :d]
    s = float(ls.exp())
    S = s * (uf @ vf.T)
    E = torch.exp(S - s + 64.0)
    rs_ref, cs_ref = E.sum(1), E.sum(0)
    G = E / rs.double()[:, None] + E / cs.double()[None, :]
    G.fill_diagonal_(0)
    wx, wy = G @ vf, G.T @ uf
    err = max(float((ax.double().sum(0) - wx).abs().max() / wx.abs().max()),
              float((ay.double().sum(0) - wy).abs().max() / wy.abs().max()),
              float(((rs.double() - rs_ref).abs() / rs_ref).max()), float(((cs.double() - cs_ref).abs() / cs_ref).max()))
    worst = max(worst, err)
    assert err < 2e-3, f"it={it} B={B} d={d} err={err}"
print(f"soak ok: {iters} shapes, worst relative error {worst:.2e}")
