"""In-kernel timeline of the tensor-core loss kernels (development aid).

Builds/loads the -DPLK_TRACE variant of the library (PLK_TRACE=1), runs one forward or backward
launch and prints, per CTA, the clock stamps relative to the CTA's entry.
usage: PLK_TRACE=1 python tools/trace_tc.py {fwd|grad} B d [n_ctas_to_print]
"""
import ctypes
import os
import sys

os.environ["PLK_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodal_plankton_recognition_b200 import _lib, ops, synth

kind, B, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nprint = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lib = _lib.load()
mode = ops.MODES["bf16"]
img, pro, _ = synth.pairs(B, d, 1234, "cuda")
ls = torch.ones((), device="cuda")
u, *_ = ops.l2norm(img, mode)
v, *_ = ops.l2norm(pro, mode)
SLOTS = 128
buf = torch.zeros(4096 * SLOTS, dtype=torch.int64, device="cuda")
setter = lib.cdll.plk_debug_set_trace
setter.argtypes = [ctypes.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


WARM = os.environ.get("PLK_TRACE_WARM", "0") == "1"   # 1: as inside a step (operands just written, L2-resident)
gs0 = torch.zeros(1, device="cuda")


def run():
    global u, v
    flush.zero_()
    if WARM:
        u, *_ = ops.l2norm(img, mode)
        v, *_ = ops.l2norm(pro, mode)
    if kind == "fwd":
        buf.zero_()
        ops.infonce_fwd_local(u, v, mode, d, 0, B, ls)
        return
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, B, ls)
    if not WARM:
        flush.zero_()
    buf.zero_()
    torch.cuda.synchronize()
    ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, B, ls, rs, cs, cs, rs, gs0)


assert setter(buf.data_ptr()) == 0
for it in range(3):
    run()
    torch.cuda.synchronize()
# the same launch between CUDA events (what bench.py's kernel_ms sees) against the CTAs' own wall clocks
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
if kind == "fwd":
    flush.zero_(); buf.zero_()
    ev[0].record(); ops.infonce_fwd_local(u, v, mode, d, 0, B, ls); ev[1].record()
else:
    rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, B, ls)
    flush.zero_(); buf.zero_()
    torch.cuda.synchronize()   # the backward launches with programmatic stream serialisation: it may start under buf.zero_()
    ev[0].record(); ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, B, ls, rs, cs, cs, rs, gs0); ev[1].record()
torch.cuda.synchronize()
print(f"CUDA events around the launch: {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
t = buf.view(-1, SLOTS).cpu().numpy()
used = [i for i in range(t.shape[0]) if t[i, 0] != 0]
print(f"{len(used)} CTAs traced")


def rel(row, k):
    return int(row[k] - row[0]) if row[k] else -1


tot = np.array([rel(t[i], 6) for i in used])
gt0 = np.array([t[i, 40] for i in used]); gt1 = np.array([t[i, 41] for i in used])
print(f"wall clock: first entry -> last exit {(gt1.max() - gt0.min()) / 1e3:.1f} us; entry spread {(gt0.max() - gt0.min()) / 1e3:.1f} us; "
      f"exit spread {(gt1.max() - gt1.min()) / 1e3:.1f} us; median lifetime {np.median(gt1 - gt0) / 1e3:.1f} us")
print(f"CTA lifetime (entry -> exit): min {tot.min()} median {int(np.median(tot))} max {tot.max()} cycles")
for i in used[:nprint] + used[-1:]:
    r = t[i]
    print(f"--- CTA {i}: setup {rel(r,1)}  A-ready(epi) {rel(r,2)}  mma-start {rel(r,3)}  first-chunk-mma {rel(r,4)}  "
          f"joined {rel(r,5)}  exit {rel(r,6)}")
    if r[15]:   # epilogue phases of tile 4 (tc4): relative to the named barrier at the top of the iteration
        names = ("ld done", "sempty sent", "G computed", "gempty seen", "G stored", "fenced")
        print("   epilogue tile 4: at barrier", rel(r, 14), "passed", rel(r, 15), "got S", rel(r, 84),
              " ".join(f"{n} +{int(r[8 + i] - r[84])}" for i, n in enumerate(names)))
    for name, base in (("S exec (probe)", 24), ("GV exec (probe)", 32)):
        vals = [int(r[base + k]) for k in range(16) if r[base + k]] if kind != "fwd" else []
        if vals:
            print(f"   {name}", " ".join(f"{x:6d}" for x in vals))
    if kind == "fwd":
        for name, base in (("w2 ld done   ", 8), ("w17 ld done  ", 24), ("w17 epi done ", 32)):
            vals = [rel(r, base + k) for k in range(8) if r[base + k]]
            print(f"   {name}", " ".join(f"{x:6d}" for x in vals))
    for name, base in (("tma issued   ", 48), ("S committed  ", 64), ("epi got S    ", 80), ("epi done     ", 96),
                       ("GV issued    ", 112)):
        vals = [rel(r, base + k) for k in range(16) if r[base + k]]
        if vals:
            print(f"   {name}", " ".join(f"{x:6d}" for x in vals))
