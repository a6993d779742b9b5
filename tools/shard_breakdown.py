"""Per-kernel device time of ONE rank's share of the row-sharded loss step, measured on a single GPU (the
collectives are left out): what the kernels cost at n owned rows x B global columns, to separate kernel time
from exchange time in the multi-GPU numbers.   usage: python tools/shard_breakdown.py [B=32768] [d=512] [R=8]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
R = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n, off = B // R, (B // R) * (R // 2)
mode = ops.MODES["bf16"]
dev = torch.device("cuda:0")
img, pro, _ = synth.pairs(B, d, 4321, dev)
ls = torch.ones((), device=dev)
go = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


u_all, *_ = ops.l2norm(img, mode)
v_all, *_ = ops.l2norm(pro, mode)
x, y = img[off:off + n].contiguous(), pro[off:off + n].contiguous()
st4 = torch.empty((4, n), device=dev)
stats = torch.zeros((3, B), device=dev)
cs_all, rs_all, dg_all = stats.unbind(0)
rs, dg = rs_all[off:off + n], dg_all[off:off + n]
t_norm = timed(lambda: ops.l2norm_pair(x, y, mode, st4, stats))
u, v = ops.l2norm_pair(x, y, mode, st4, stats)
t_fwd = timed(lambda: ops.infonce_fwd_local(u, v_all, mode, d, off, B, ls, rs, cs_all, dg, sums_zeroed=True))
# complete statistics from an unsharded forward
rs_f, cs_f, dg_f = ops.infonce_fwd_local(u_all, v_all, mode, d, 0, B, ls)
t_loss = timed(lambda: ops.infonce_loss_local(rs_f[off:off + n], cs_f[off:off + n], dg_f[off:off + n], ls, B))
_, aux = ops.infonce_loss_local(rs_f[off:off + n], cs_f[off:off + n], dg_f[off:off + n], ls, B)
gs = aux[1:]
t_bwd = timed(lambda: ops.infonce_grad_pair_local(u, v_all, v, u_all, mode, d, off, B, ls, rs_f[off:off + n], cs_f,
                                                  cs_f[off:off + n], rs_f, gs))
acc_x, acc_y = ops.infonce_grad_pair_local(u, v_all, v, u_all, mode, d, off, B, ls, rs_f[off:off + n], cs_f,
                                           cs_f[off:off + n], rs_f, gs)
idx, nx, idy, ny = st4.unbind(0)
t_fin = timed(lambda: ops.infonce_grad_finish_pair(acc_x, acc_y, x, y, (idx, nx), (idy, ny), dg_f[off:off + n],
                                                   rs_f[off:off + n], cs_f[off:off + n], ls, go, go, B, gs, aux[0:1]))
tot = t_norm + t_fwd + t_loss + t_bwd + t_fin
print(f"B={B} d={d} ranks={R}: n={n} owned rows, parts={acc_x.shape[0]}")
print(f"  l2norm_pair {t_norm:8.1f} us\n  forward     {t_fwd:8.1f} us  ({2.0 * n * B * d / t_fwd / 1e6:7.1f} TFLOP/s)\n"
      f"  loss        {t_loss:8.1f} us\n  backward    {t_bwd:8.1f} us  ({4.0 * n * B * d / t_bwd / 1e6:7.1f} TFLOP/s credited)\n"
      f"  grad tail   {t_fin:8.1f} us\n  sum         {tot:8.1f} us   (each launch timed alone, L2 flushed before it)")
