"""A/B timing of single hot kernels at the bench shapes: one launch captured in a CUDA graph, L2 flushed before
every timed replay, CUDA events (the method bench.py uses for `roofline.kernel_ms`).  Environment switches
such as PLK_GRAD_TC2=1 are read once per process, so run it once per variant:
    python tools/kernel_ab.py            ;  PLK_GRAD_TC2=1 python tools/kernel_ab.py
Also checks the pair backward against a plain torch evaluation of the same bf16 operands (max-abs / max-abs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth

mode = ops.MODES[os.environ.get("PLK_AB_PRECISION", "bf16")]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def graphed(fn):
    fn()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    g.keep = keep
    return g.replay


def timed(fn, n=50):
    for _ in range(3):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    # CUDA event stamps tick every 2.048 us on these boxes: medians land on multiples of it, the MEAN resolves finer
    return sum(ts) / len(ts) * 1e3, ts[0] * 1e3      # mean, best (us)


shapes = [(4096, 256, 1), (4096, 128, 1), (8192, 256, 1), (16384, 256, 1), (4096, 256, 8), (4096, 512, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
print(f"variant: PLK_GRAD_TC2={os.environ.get('PLK_GRAD_TC2', '0')} PLK_FWD_OLD={os.environ.get('PLK_FWD_OLD', '0')}")
for B, d, buckets in shapes:
    bs = B // buckets
    img, pro, _ = synth.pairs(B, d, 1, "cuda")
    ls = torch.ones((), device="cuda")
    u, *_ = ops.l2norm(img, mode)
    v, *_ = ops.l2norm(pro, mode)
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, bs, ls)
    torch.cuda.synchronize()
    f_med, f_best = timed(graphed(lambda: ops.infonce_fwd_local(u, v, mode, d, 0, bs, ls, rs, cs, dg)))
    # the timed forward accumulated into rs / cs again: recompute clean statistics
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, bs, ls)
    gs = torch.zeros(1, device="cuda")
    acc_x, acc_y = ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, bs, ls, rs, cs, cs, rs, gs)
    torch.cuda.synchronize()
    # reference from the same 16-bit operands (fp64 on the GPU; the j == i term is left to the gradient tail)
    err = float("nan")
    if B <= 8192:
        uf, vf = u.double()[:, :d], v.double()[:, :d]
        s = float(ls.exp())
        S = s * (uf @ vf.T)
        m = (torch.arange(B, device="cuda")[:, None] // bs) == (torch.arange(B, device="cuda")[None, :] // bs)
        E = torch.where(m, torch.exp(S - s + 64.0), torch.zeros_like(S))
        G = E * (1.0 / rs.double())[:, None] + E * (1.0 / cs.double())[None, :]
        gs_want = float((G * S).sum())
        G.fill_diagonal_(0)
        wx, wy = G @ vf, G.T @ uf
        err = max(float((acc_x.double().sum(0) - wx).abs().max() / wx.abs().max()),
                  float((acc_y.double().sum(0) - wy).abs().max() / wy.abs().max()))
        gerr = abs(float(gs) - gs_want) / abs(gs_want)
        del S, E, G, m
    else:
        gerr = float("nan")
    b_med, b_best = timed(graphed(lambda: ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, bs, ls, rs, cs, cs, rs, None)))
    fl = 2.0 * B * bs * d
    print(f"B={B} d={d} buckets={buckets}: fwd {f_med:7.1f} us (best {f_best:7.1f}) {fl / f_med / 1e6:7.1f} TFLOP/s | "
          f"bwd pair {b_med:7.1f} us (best {b_best:7.1f}) credited {2 * fl / b_med / 1e6:7.1f} TFLOP/s parts={acc_x.shape[0]} "
          f"| acc err {err:.2e} gs err {gerr:.2e}", flush=True)
