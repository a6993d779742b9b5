"""Small fwd+bwd / retrieval cases touching every kernel once (CLIP + SigLIP, bf16 + fp32, ragged shapes):
a quick end-to-end sanity run, e.g. after changing a launch configuration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodal_plankton_recognition_b200 import ANNClassifier, CLIPLoss, SigLIPLoss, synth

dev = torch.device("cuda:0")
for cls in (CLIPLoss, SigLIPLoss):
    for prec in ("bf16", "fp32"):
        for B, d, bk in ((384, 256, 1), (500, 200, 5), (256, 512, 2), (1152, 256, 1), (896, 512, 1), (640, 320, 1)):
            img, pro, _ = synth.pairs(B, d, 3, dev)
            mod = cls(precision=prec).to(dev)
            x, y = img.requires_grad_(), pro.requires_grad_()
            loss = mod(image_emb=x, profile_emb=y, buckets=bk)
            loss.backward()
            torch.cuda.synchronize()
            assert torch.isfinite(loss) and torch.isfinite(x.grad).all()
# N1: fused projection + normalisation (d = 96 / 320 / 512: one and two MMAs per K step)
for prec in ("bf16", "fp32"):
    for B, f, d in ((300, 200, 96), (256, 192, 512), (384, 320, 320)):
        mod = CLIPLoss(precision=prec).to(dev)
        pi, pp = torch.nn.Linear(f, d, bias=False).to(dev), torch.nn.Linear(f + 8, d, bias=False).to(dev)
        xi = torch.randn(B, f, device=dev, requires_grad=True)
        xp = torch.randn(B, f + 8, device=dev, requires_grad=True)
        loss = mod.forward_projected(xi, xp, pi, pp)
        loss.backward()
        torch.cuda.synchronize()
        assert torch.isfinite(loss) and torch.isfinite(xi.grad).all() and torch.isfinite(pi.weight.grad).all()
g = np.random.default_rng(0)
gal = g.standard_normal((700, 128)).astype(np.float32)
q = g.standard_normal((130, 128)).astype(np.float32)
for prec in ("bf16", "fp32"):
    ANNClassifier(gal, g.integers(0, 5, 700), metric="euclidean", plk_precision=prec).predict(q, k=5, epsilon=.3)
torch.cuda.synchronize()
print("sanitize cases ok")
