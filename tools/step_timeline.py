"""Where the loss step's time goes at the bench shape: forward half, backward half and the whole step, each as one
CUDA graph, L2 flushed before every timed replay, CUDA events (bench.py's method).  usage: step_timeline.py [B d]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = ops.MODES["bf16"]
img, pro, _ = synth.pairs(B, d, 1234, "cuda")
ls = torch.ones((), device="cuda")
go = torch.ones(1, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def graphed(fn):
    fn()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    g.keep = keep
    return g.replay


def timed(fn, n=100, do_flush=True):
    for _ in range(5):
        if do_flush:
            flush.zero_()
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        if do_flush:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    # CUDA event stamps tick every 2.048 us on these boxes: medians land on multiples of it, the MEAN resolves finer
    return sum(ts) / len(ts) * 1e3, ts[0] * 1e3


loss, state = ops.clip_loss_forward_state(img, pro, ls, B, mode)


def fwd():
    return ops.clip_loss_forward_state(img, pro, ls, B, mode)


def bwd():
    return ops.clip_loss_backward_state(go, img, pro, ls, state, B, mode)


def whole():
    l, s = ops.clip_loss_forward_state(img, pro, ls, B, mode)
    return (l,) + tuple(ops.clip_loss_backward_state(go, img, pro, ls, s, B, mode))


def empty():
    return None


k = torch.zeros(1, device="cuda")
for name, fn in (("one tiny kernel (launch + event floor)", lambda: k.add_(1)), ("forward half (normalise + forward + loss)", fwd),
                 ("backward half (backward + tail)", bwd), ("whole step", whole)):
    g = graphed(fn)
    m, b = timed(g)
    m2, b2 = timed(g, do_flush=False)
    if name == "whole step":   # the same C calls launched directly on the stream (no graph): the host runs ahead under the flush
        md, bd = timed(fn)
        print(f"{'whole step, direct launches (no graph)':45s} flushed: mean {md:7.2f} us best {bd:7.2f}")
    print(f"{name:45s} flushed: mean {m:7.2f} us best {b:7.2f} | warm: mean {m2:7.2f} best {b2:7.2f}")
