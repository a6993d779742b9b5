"""A few launches of one tensor-core kernel at a given shape (for `ncu --set full`).
usage: python tools/ncu_case.py {fwd|grad|topk} B d [ng]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_plankton_recognition_b200 import ops, synth
from multimodal_plankton_recognition_b200.ann import GpuExactIndex

kind, B, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = ops.MODES["bf16"]
if kind in ("fwd", "grad"):
    img, pro, _ = synth.pairs(B, d, 1234, "cuda")
    ls = torch.ones((), device="cuda")
    u, *_ = ops.l2norm(img, mode)
    v, *_ = ops.l2norm(pro, mode)
    for _ in range(3):
        rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, B, ls)
        if kind == "grad":
            acc = ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, B, ls, rs, cs, cs, rs, torch.zeros(1, device="cuda"))
else:
    ng = int(sys.argv[4])
    gal, _ = synth.unit_embeddings(ng, d, 5, "cuda", 1)
    q, _ = synth.unit_embeddings(B, d, 6, "cuda", 0)
    index = GpuExactIndex.from_device(gal, "bf16")
    for _ in range(3):
        index.search_device(q, 10)
torch.cuda.synchronize()
print("ok")
