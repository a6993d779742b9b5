"""cProfile of the host side of one CLIPLoss fwd+bwd (device-resident inputs, no per-step sync)."""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
from multimodal_plankton_recognition_b200 import CLIPLoss, synth  # noqa: E402

dev = torch.device("cuda", 0)
img, pro, _ = synth.pairs(4096, 256, 1, dev)
mod = CLIPLoss(precision="bf16").to(dev)
xd = img.clone().requires_grad_()
yd = pro.clone().requires_grad_()


def step():
    mod.logit_scale.grad = None
    xd.grad = None
    yd.grad = None
    mod(image_emb=xd, profile_emb=yd).backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
