"""Synthetic "CytoSense-shaped" embeddings: 27 classes with a long-tailed size
distribution, a class centroid plus a per-sample latent shared by the two modalities plus
modality noise -- un-normalised, like the projection outputs of reference src/model.py:80-83."""
from __future__ import annotations

import numpy as np
import torch

N_CLASSES = 27


def class_sizes(n: int, n_classes: int = N_CLASSES) -> np.ndarray:
    """Long-tailed (Zipf-like) class sizes summing to n, every class non-empty."""
    w = 1.0 / np.arange(1, n_classes + 1) ** 1.1
    sizes = np.maximum(1, np.floor(w / w.sum() * n)).astype(np.int64)
    sizes[0] += n - sizes.sum()
    return sizes


def labels_for(n: int, seed: int, n_classes: int = N_CLASSES) -> np.ndarray:
    lab = np.repeat(np.arange(n_classes), class_sizes(n, n_classes))
    np.random.default_rng(seed).shuffle(lab)
    return lab


def pairs(n: int, d: int, seed: int = 1234, device="cpu", n_classes: int = N_CLASSES, separation: float = 1.0):
    """-> (image_emb [n,d] fp32, profile_emb [n,d] fp32, labels [n] int64) on `device`.
    `separation` scales the class centroids: 1.0 gives well separated classes, ~0.15 a k-NN accuracy of
    80-90 % at d = 512 (the few-shot benchmark wants votes that are not unanimous)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cent = torch.randn(n_classes, d, generator=g)
    lab = torch.from_numpy(labels_for(n, seed, n_classes))
    dev = torch.device(device)
    gd = torch.Generator(device=dev).manual_seed(seed + 1)
    c = separation * cent.to(dev)[lab.to(dev)]
    z = torch.randn(n, d, device=dev, generator=gd)
    img = c + 0.5 * z + 0.3 * torch.randn(n, d, device=dev, generator=gd)
    pro = c + 0.5 * z + 0.3 * torch.randn(n, d, device=dev, generator=gd)
    return img, pro, lab.to(dev)


def unit_embeddings(n: int, d: int, seed: int, device="cpu", modality: int = 0):
    """L2-normalised fp32 embeddings as the benchmark scripts eat them (experiments.ipynb cell 4)."""
    img, pro, lab = pairs(n, d, seed, device)
    e = img if modality == 0 else pro
    return torch.nn.functional.normalize(e), lab
