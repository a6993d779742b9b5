"""GPU-time breakdown of one InfoNCE fwd+bwd step (each piece captured 20x in a CUDA graph, so
python/launch overhead on the host does not show up)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = ops.MODES[sys.argv[3] if len(sys.argv) > 3 else "bf16"]
img, pro, _ = synth.pairs(B, d, 1234, "cuda")
ls = torch.ones((), device="cuda")
go = torch.ones(1, device="cuda")
REP = 20


def gtime(fn, name):
    fn(); fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / (5 * REP) * 1e3
    print(f"{name:28s} {us:8.2f} us")
    return us


stats = torch.empty((7, B), device="cuda")
u, *_ = ops.l2norm(img, mode, True, stats[0], stats[1])
v, *_ = ops.l2norm(pro, mode, True, stats[2], stats[3])
ops.infonce_fwd_local(u, v, mode, d, 0, B, ls, stats[4], stats[5], stats[6])
loss, aux = ops.infonce_loss_local(stats[4], stats[5], stats[6], ls, B)
gs = torch.zeros(1, device="cuda")
acc = ops.infonce_grad_local(u, v, mode, d, 0, B, ls, stats[4], stats[5], gs)
tot = 0
tot += 2 * gtime(lambda: ops.l2norm(img, mode, True, stats[0], stats[1]), "l2norm (x1)")
tot += gtime(lambda: ops.infonce_fwd_local(u, v, mode, d, 0, B, ls, stats[4], stats[5], stats[6]), "fwd (+3 memsets)")
tot += gtime(lambda: ops.infonce_loss_local(stats[4], stats[5], stats[6], ls, B), "loss")
tot += 2 * gtime(lambda: ops.infonce_grad_local(u, v, mode, d, 0, B, ls, stats[4], stats[5], None), "grad (x1, no gs)")
gtime(lambda: ops.infonce_grad_local(u, v, mode, d, 0, B, ls, stats[4], stats[5], gs), "grad (x1, with gs)")
tot += 2 * gtime(lambda: ops.infonce_grad_finish(acc, img, pro, stats[0], stats[1], stats[2], stats[6], stats[4], stats[5], ls, go, B, torch.float32), "grad_finish (x1)")
tot += gtime(lambda: ops.infonce_dls(gs, aux[0:1], go, B), "dls")
print(f"sum of parts (2x l2norm, 2x grad, 2x finish): {tot:8.2f} us")


def step():
    l, st = ops.clip_loss_forward_state(img, pro, ls, img.shape[0], mode)
    ops.clip_loss_backward_state(go, img, pro, ls, st, img.shape[0], mode)


gtime(step, "whole step (custom ops)")
