"""Per-kernel timing sweep (CUDA events, 10 back-to-back launches) to separate fixed from per-tile cost."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth, _lib
from multimodal_plankton_recognition_b200.ann import GpuExactIndex

mode = ops.MODES["bf16"]
lib = _lib.load()


def timeit(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us


print("kernel      B      d  tiles/CTA  CTAs   us    clk/tile(@1.9GHz)  TFLOP/s(2*B*B*d per GEMM)")
for B, d in [(4096, 64), (4096, 256), (8192, 256), (16384, 256), (32768, 256), (4096, 512), (16384, 512), (32768, 512)]:
    img, pro, _ = synth.pairs(B, d, 1, "cuda")
    ls = torch.ones((), device="cuda")
    u, *_ = ops.l2norm(img, mode)
    v, *_ = ops.l2norm(pro, mode)
    rs = torch.empty(B, device="cuda"); cs = torch.empty(B, device="cuda"); dg = torch.empty(B, device="cuda")
    ops.infonce_fwd_local(u, v, mode, d, 0, B, ls, rs, cs, dg)
    rb = B // 128
    nseg = max(1, min(148 // rb, B // 128))
    tiles = -(-(B // 128) // nseg)
    t = timeit(lambda: ops.infonce_fwd_local(u, v, mode, d, 0, B, ls, rs, cs, dg))
    waves = -(-(rb * nseg) // 148)
    print(f"fwd    {B:6d} {d:5d} {tiles:6d} {rb * nseg:6d} {t:8.1f} {t * 1900 / (tiles * waves):10.0f} {2.0 * B * B * d / t / 1e6:10.1f}")
    parts = lib.plk_infonce_grad_parts(mode, B, B, d, B)
    z = 2 if d > 256 else 1
    nseg = parts
    tiles = -(-(B // 128) // nseg)
    acc = torch.empty((parts, B, d), device="cuda")
    def g():
        lib.check(lib.plk_infonce_grad(u.data_ptr(), v.data_ptr(), mode, u.stride(0), B, 0, B, d, B, ls.data_ptr(),
                                       rs.data_ptr(), cs.data_ptr(), acc.data_ptr(), None,
                                       torch.cuda.current_stream().cuda_stream))
    t = timeit(g)
    waves = -(-(rb * nseg * z) // 148)
    print(f"grad   {B:6d} {d:5d} {tiles:6d} {rb * nseg * z:6d} {t:8.1f} {t * 1900 / (tiles * waves):10.0f} {2.0 * B * B * d / t / 1e6:10.1f}")
    del acc

for nq, ng, d in [(8192, 131072, 512), (8192, 1 << 20, 512), (32768, 1 << 20, 512), (32768, 1 << 20, 256)]:
    gal, _ = synth.unit_embeddings(ng, d, 5, "cuda", 1)
    q, _ = synth.unit_embeddings(nq, d, 6, "cuda", 0)
    index = GpuExactIndex.from_device(gal, "bf16")
    t = timeit(lambda: index.search_device(q, 10), 3)
    print(f"topk  nq={nq} ng={ng} d={d}: {t / 1e3:8.2f} ms  {2.0 * nq * ng * d / t / 1e6:8.1f} TFLOP/s")
    del index, gal, q
