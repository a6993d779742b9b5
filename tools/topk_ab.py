"""A/B of the top-k candidate kernel's rasterisation (PLK_TOPK_RASTER=0: chunk-fastest, the round-1 order).
usage: python tools/topk_ab.py [nq ng d]   -- prints ms per search (CUDA events, 5 searches) and a checksum."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import synth
from multimodal_plankton_recognition_b200.ann import GpuExactIndex

shapes = [(8192, 262144, 512), (100000, 1000000, 512)]
if len(sys.argv) > 3:
    shapes = [tuple(int(x) for x in sys.argv[1:4])]
for nq, ng, d in shapes:
    gal, _ = synth.unit_embeddings(ng, d, 5, "cuda", 1)
    q, _ = synth.unit_embeddings(nq, d, 6, "cuda", 0)
    index = GpuExactIndex.from_device(gal, "bf16")
    idx, dist = index.search_device(q, 10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        idx, dist = index.search_device(q, 10)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"raster={os.environ.get('PLK_TOPK_RASTER', '1')} nq={nq} ng={ng} d={d}: {ms:8.3f} ms  "
          f"{2.0 * nq * ng * d / ms / 1e9:7.1f} TFLOP/s  checksum idx {int(idx.long().sum())} dist {float(dist.double().sum()):.6f}")
    del gal, q, index
