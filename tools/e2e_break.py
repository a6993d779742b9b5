"""Where the end-to-end step time goes (host-resident inputs -> loss on the host), B=4096, d=256.

A  H2D only                      B  fwd+bwd, device inputs, sync per step
C  fwd+bwd, device inputs, no per-step sync (host enqueue rate)
D  serial e2e (bench.py's definition)      E  prefetched e2e (HostPairPrefetcher)
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from multimodal_plankton_recognition_b200 import CLIPLoss, synth  # noqa: E402


def timeit(fn, steps=100, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e6


def main():
    dev = torch.device("cuda", 0)
    n, d = 4096, 256
    img, pro, _ = synth.pairs(n, d, 1, dev)
    hx, hy = img.cpu().pin_memory(), pro.cpu().pin_memory()
    mod = CLIPLoss(precision="bf16").to(dev)

    def a():
        hx.to(dev, non_blocking=True)
        hy.to(dev, non_blocking=True)
        torch.cuda.synchronize()

    xd = img.clone().requires_grad_()
    yd = pro.clone().requires_grad_()

    def c():
        mod.logit_scale.grad = None
        xd.grad = None
        yd.grad = None
        mod(image_emb=xd, profile_emb=yd).backward()

    def b():
        c()
        torch.cuda.synchronize()

    def dd():
        x = hx.to(dev, non_blocking=True).requires_grad_()
        y = hy.to(dev, non_blocking=True).requires_grad_()
        mod.logit_scale.grad = None
        loss = mod(image_emb=x, profile_emb=y)
        loss.backward()
        return float(loss)

    bx, by = torch.empty_like(img), torch.empty_like(pro)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        bx.copy_(hx, non_blocking=True); by.copy_(hy, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        bx.copy_(hx, non_blocking=True); by.copy_(hy, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(f"A0 h2d 8 MiB back-to-back (events) {us:8.1f} us  = {8.388608e6 / us / 1e3:.1f} GB/s")
    print(f"A h2d only            {timeit(a):8.1f} us")
    print(f"B fwd+bwd sync        {timeit(b):8.1f} us")
    print(f"C fwd+bwd enqueue     {timeit(c):8.1f} us")
    print(f"D serial e2e          {timeit(dd):8.1f} us")

    try:
        from multimodal_plankton_recognition_b200.prefetch import HostPairPrefetcher
    except ImportError:
        return
    steps = 200
    def batches(k):
        for _ in range(k):
            yield hx, hy
    for depth in (2, 3):
        pf = HostPairPrefetcher(batches(steps + 5), dev, depth=depth)
        it = iter(pf)
        pending = None
        losses = []
        def one():
            nonlocal pending
            x, y = next(it)
            x.requires_grad_(); y.requires_grad_()
            mod.logit_scale.grad = None
            loss = mod(image_emb=x, profile_emb=y)
            loss.backward()
            host = pf.read_async(loss)
            if pending is not None:
                losses.append(pending())
            pending = host
        for _ in range(5):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        losses.append(pending())
        torch.cuda.synchronize()
        print(f"E prefetched depth={depth}  {(time.perf_counter() - t0) / steps * 1e6:8.1f} us  loss {losses[-1]:.6f}")


if __name__ == "__main__":
    main()
