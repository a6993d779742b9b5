"""CPU emulation of the 16-bit tensor-core loss path: which rounding dominates the gradient error?

Rounds (a) the normalised operands u, v and (b) the recomputed softmax weights G to bf16 / fp16
independently (oracle.infonce.clip_loss_grads_rounded_operands) and reports the two gradient error metrics
of tests/test_gpu_loss.py against the fp64 closed form.  Everything else is fp64 here, so the numbers are
the floor a kernel with that operand format can reach.  No GPU needed.
  python tools/bf16_error_budget.py
"""
import glob
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import infonce as oinf   # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def case(B, d, seed=None):
    r = np.random.default_rng(B + d if seed is None else seed)
    cent = r.standard_normal((27, d))
    lab = r.integers(0, 27, B)
    z = r.standard_normal((B, d))
    img = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    pro = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    return img, pro


def report(tag, img, pro, ls, bk, ref_dx, ref_dy):
    print(tag)
    for op_fmt, g_fmt in itertools.product(("f64", "bf16", "fp16"), ("f64", "bf16", "fp16")):
        dx, dy = oinf.clip_loss_grads_rounded_operands(img, pro, ls, bk, op_fmt, g_fmt)
        print(f"  operands {op_fmt:5s} G {g_fmt:5s}: max {rel(dx, ref_dx):.2e} {rel(dy, ref_dy):.2e}"
              f"  l2 {rel_l2(dx, ref_dx):.2e} {rel_l2(dy, ref_dy):.2e}")


if __name__ == "__main__":
    for (B, d, bk, ls) in [(1024, 384, 8, 2.659), (2048, 256, 1, 2.659), (1024, 256, 1, 1.0), (2048, 256, 1, 3.7)]:
        img, pro = case(B, d)
        ref = oinf.clip_loss_closed_form(img, pro, ls, bk)
        report(f"B={B} d={d} buckets={bk} ls={ls}", img, pro, ls, bk, ref["d_image"], ref["d_profile"])
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "loss_*.npz"))):
        if "edge" in p:
            continue
        g = np.load(p)
        report(os.path.basename(p), g["image"], g["profile"], float(g["logit_scale"]), int(g["buckets"]),
               g["d_image_f64"], g["d_profile_f64"])
