"""GPU bring-up probe: runs each libplk kernel family on small shapes against plain torch math and
prints error statistics.  Each case runs in its own subprocess under a timeout so that a hang or a
trap in one kernel does not hide the others.   usage: python tools/probe_tc.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _mk(B, d, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, d, device="cuda", generator=g)
    y = x + 0.7 * torch.randn(B, d, device="cuda", generator=g)
    return x, y


def _fwd_case(B, d, buckets, mode_name):
    import torch
    from multimodal_plankton_recognition_b200 import ops
    mode = ops.MODES[mode_name]
    x, y = _mk(B, d)
    ls = torch.tensor(1.0, device="cuda")
    u, idx, nx, _ = ops.l2norm(x, mode)
    v, idy, ny, _ = ops.l2norm(y, mode)
    bs = B // buckets
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, bs, ls)
    torch.cuda.synchronize()
    uf, vf = u.float()[:, :d], v.float()[:, :d]
    s = float(ls.exp())
    S = (uf.double() @ vf.double().T) * s
    mask = (torch.arange(B, device="cuda")[:, None] // bs) == (torch.arange(B, device="cuda")[None, :] // bs)
    E = torch.where(mask, torch.exp(S - s + 64.0), torch.zeros_like(S))   # kShiftK (csrc/common.cuh)
    err = lambda a, b: float(((a.double() - b).abs() / b.abs().clamp_min(1e-30)).max())
    print(f"fwd[{mode_name}] B={B} d={d} bk={buckets}: rs {err(rs, E.sum(1)):.2e} cs {err(cs, E.sum(0)):.2e} "
          f"diag {float((dg.double() - S.diagonal()).abs().max()):.2e}", flush=True)
    # backward pieces
    gs = torch.zeros(1, device="cuda")
    acc = ops.infonce_grad_local(u, v, mode, d, 0, bs, ls, rs, cs, gs)
    torch.cuda.synchronize()
    G = E * (1.0 / E.sum(1))[:, None] + E * (1.0 / E.sum(0))[None, :]
    want = (G - torch.diag(G.diagonal())) @ vf.double()      # the j == i term is left to grad_finish
    got = acc.double().sum(0)
    print(f"grad[{mode_name}] parts={acc.shape[0]}: acc {float((got - want).abs().max() / want.abs().max()):.2e} "
          f"gs {abs(float(gs) - float((G * S).sum())) / abs(float((G * S).sum())):.2e}", flush=True)


@case
def simt_small():
    _fwd_case(256, 128, 2, "fp32")


@case
def tc_1tile():
    _fwd_case(128, 64, 1, "bf16")


@case
def tc_small():
    _fwd_case(256, 256, 1, "bf16")


@case
def tc_c2():
    _fwd_case(4096, 256, 1, "bf16")


@case
def tc_d512():
    _fwd_case(1024, 512, 4, "bf16")


@case
def tc_ragged():
    _fwd_case(1000, 200, 5, "bf16")


@case
def topk():
    import torch
    from multimodal_plankton_recognition_b200.ann import GpuExactIndex
    for prec in ("fp32", "bf16"):
        g = torch.Generator(device="cuda").manual_seed(1)
        gal = torch.nn.functional.normalize(torch.randn(5000, 256, device="cuda", generator=g))
        q = torch.nn.functional.normalize(torch.randn(300, 256, device="cuda", generator=g))
        index = GpuExactIndex(gal.cpu().numpy(), precision=prec)
        idx, dist = index.search_device(q, 10)
        torch.cuda.synchronize()
        full = torch.cdist(q.double(), gal.double()).float()
        best = full.topk(10, dim=1, largest=False)
        print(f"topk[{prec}]: idx match {float((best.indices.int() == idx).float().mean()):.4f} "
              f"dist err {float((best.values - dist).abs().max()):.2e}", flush=True)


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and names[0] in CASES and os.environ.get("PLK_PROBE_CHILD"):
        CASES[names[0]]()
        return
    rc = 0
    for n in names:
        env = dict(os.environ, PLK_PROBE_CHILD="1")
        try:
            r = subprocess.run([sys.executable, __file__, n], env=env, capture_output=True, text=True, timeout=180)
            tail = (r.stdout + r.stderr).strip().splitlines()[-12:]
            print(f"=== {n}: exit {r.returncode}")
            print("\n".join(tail), flush=True)
            rc |= r.returncode != 0
        except subprocess.TimeoutExpired as e:
            print(f"=== {n}: TIMEOUT\n{(e.stdout or b'').decode()[-2000:]}", flush=True)
            rc = 1
    sys.exit(rc)


if __name__ == "__main__":
    main()
