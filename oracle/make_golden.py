"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):

    python oracle/make_golden.py

* loss_*.npz      inputs + outputs of the reference's own ``CLIPLoss``
                  (reference src/coordination.py:17-47), fp64 and fp32, incl. autograd
                  gradients w.r.t. both embeddings and ``logit_scale``.
* siglip_*.npz    the same for the reference's ``SigLIPLoss`` (reference src/coordination.py:67-95),
                  incl. the gradient w.r.t. ``bias``.
* bench_*.npz     inputs + every predicted label array of the reference's own benchmark drivers
                  (reference scripts/benchmark_cross.py:24-87, scripts/benchmark_cross_folds.py:24-85,
                  scripts/benchmark_raw.py:24-50, scripts/benchmark_folds.py:24-51),
                  same index injection as ann_*.
* ann_*.npz       inputs + outputs of the reference's own ``ANNClassifier``
                  (reference src/ann.py:6-34) executed with ``oracle.ann.ExactIndex``
                  injected as the module ``pynndescent`` (the real package is not
                  installed here), i.e. the reference's hstack / weights /
                  weighted_mode code runs unmodified.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from oracle.ann import ExactIndex  # noqa: E402

stub = types.ModuleType("pynndescent")
stub.NNDescent = ExactIndex
sys.modules["pynndescent"] = stub

from src.coordination import CLIPLoss, SigLIPLoss  # noqa: E402  (the reference)
from src.ann import ANNClassifier      # noqa: E402  (the reference)


def synth_pairs(B, d, seed, n_classes=27):
    g = np.random.default_rng(seed)
    cent = g.standard_normal((n_classes, d))
    lab = g.integers(0, n_classes, B)
    z = g.standard_normal((B, d))
    img = cent[lab] + 0.5 * z + 0.3 * g.standard_normal((B, d))
    pro = cent[lab] + 0.5 * z + 0.3 * g.standard_normal((B, d))
    return img.astype(np.float32), pro.astype(np.float32), lab


def loss_case(name, B, d, buckets, ls, seed, tweak=None):
    img, pro, _ = synth_pairs(B, d, seed)
    if tweak:
        tweak(img, pro)
    out = {"image": img, "profile": pro, "buckets": buckets, "logit_scale": np.float64(ls)}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        mod = CLIPLoss().to(dt)
        with torch.no_grad():
            mod.logit_scale.fill_(ls)
        x = torch.tensor(img, dtype=dt, requires_grad=True)
        y = torch.tensor(pro, dtype=dt, requires_grad=True)
        loss = mod(image_emb=x, profile_emb=y, buckets=buckets)
        loss.backward()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"d_image_{tag}"] = x.grad.numpy()
        out[f"d_profile_{tag}"] = y.grad.numpy()
        out[f"d_logit_scale_{tag}"] = mod.logit_scale.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **out)
    print("wrote", name, float(out["loss_f64"]))


def siglip_case(name, B, d, buckets, ls, bias, seed, tweak=None):
    """inputs + outputs of the reference's own ``SigLIPLoss`` (reference src/coordination.py:67-95)."""
    img, pro, _ = synth_pairs(B, d, seed)
    if tweak:
        tweak(img, pro)
    out = {"image": img, "profile": pro, "buckets": buckets, "logit_scale": np.float64(ls), "bias": np.float64(bias)}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        mod = SigLIPLoss().to(dt)
        with torch.no_grad():
            mod.logit_scale.fill_(ls)
            mod.bias.fill_(bias)
        x = torch.tensor(img, dtype=dt, requires_grad=True)
        y = torch.tensor(pro, dtype=dt, requires_grad=True)
        loss = mod(image_emb=x, profile_emb=y, buckets=buckets)
        loss.backward()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"d_image_{tag}"] = x.grad.numpy()
        out[f"d_profile_{tag}"] = y.grad.numpy()
        out[f"d_logit_scale_{tag}"] = mod.logit_scale.grad.numpy()
        out[f"d_bias_{tag}"] = mod.bias.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"siglip_{name}.npz"), **out)
    print("wrote siglip", name, float(out["loss_f64"]))


def ann_case(name, ng, nq, d, k, seed, n_classes=9, dup=False, two_mod=False):
    g = np.random.default_rng(seed)
    cent = g.standard_normal((n_classes, d))
    yg = g.integers(0, n_classes, ng)
    yq = g.integers(0, n_classes, nq)

    def emb(lab, noise):
        e = cent[lab] + noise * g.standard_normal((len(lab), d))
        return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)

    gal = emb(yg, 0.9)
    q1 = emb(yq, 0.9)
    q2 = emb(yq, 1.2)
    if dup:  # zero-distance rows + exact vote ties
        q1[: min(8, nq)] = gal[: min(8, nq)]
        gal[ng // 2: ng // 2 + 4] = gal[:4]
    clf = ANNClassifier(gal, yg, n_neighbors=32, metric="euclidean", diversify_prob=0.0,
                        pruning_degree_multiplier=3.0, low_memory=False, random_state=0)
    qs = (q1, q2) if two_mod else (q1,)
    nbrs = clf.kneighbors(*qs, k=k, epsilon=.3)
    pred = clf.predict(*qs, k=k, epsilon=.3)
    out = {"gallery": gal, "labels": yg, "k": k, "pred": pred}
    for m, (q, (i, dd)) in enumerate(zip(qs, nbrs)):
        out[f"query{m}"] = q
        out[f"idx{m}"] = i
        out[f"dist{m}"] = dd
    np.savez_compressed(os.path.join(OUT, f"ann_{name}.npz"), **out)
    print("wrote ann", name, pred[:8])


def bench_case(name, n_classes, per_class, d, n, repeats, K, seed):
    """Run the reference's own drivers -- `benchmark()` of scripts/benchmark_cross.py:24-87 and of
    scripts/benchmark_cross_folds.py:24-85 -- unmodified, with oracle.ann.ExactIndex standing in for
    pynndescent, on a synthetic embedding set; store the inputs, the `random` seed and every predicted
    label array (as class ids, key order = run, k, set-up)."""
    import importlib.util
    import random
    from sklearn.preprocessing import LabelEncoder

    def load(path, modname):
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    cross = load("/root/reference/scripts/benchmark_cross.py", "ref_benchmark_cross")
    folds = load("/root/reference/scripts/benchmark_cross_folds.py", "ref_benchmark_cross_folds")
    g = np.random.default_rng(seed)
    cent = g.standard_normal((n_classes, d))
    lab = np.repeat(np.arange(n_classes), per_class)
    g.shuffle(lab)

    def emb(noise):
        e = cent[lab] + noise * g.standard_normal((len(lab), d))
        return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)

    img, pro = emb(0.45 * np.sqrt(d)), emb(0.55 * np.sqrt(d))   # hard enough that the votes matter (acc ~ 0.5-0.8)
    names = np.array([f"class_{c:02d}" for c in lab])
    coder = LabelEncoder().fit(names)
    out = {"image": img, "profile": pro, "names": names, "n": n, "repeats": repeats, "K": np.array(K), "seed": seed}

    def flatten(res, tag):
        for run in sorted(res):
            out[f"{tag}/true/{run}"] = coder.transform(res[run]["true"])
            for k in K:
                for setup, pred in res[run]["pred"][k].items():
                    out[f"{tag}/pred/{run}/{k}/{setup}"] = coder.transform(pred)

    random.seed(seed)
    flatten(cross.benchmark((img, pro, names), coder, n, repeats, K), "cross")
    half = len(lab) // 2
    random.seed(seed)
    flatten(folds.benchmark((img[:half], pro[:half], names[:half]), (img[half:], pro[half:], names[half:]),
                            coder, n, repeats, K), "folds")
    # the single set-up drivers: I+P gallery, (I, P) query
    raw = load("/root/reference/scripts/benchmark_raw.py", "ref_benchmark_raw")
    jfolds = load("/root/reference/scripts/benchmark_folds.py", "ref_benchmark_folds")

    def flatten_joint(res, tag):
        for run in sorted(res):
            out[f"{tag}/true/{run}"] = coder.transform(res[run]["true"])
            for k in K:
                out[f"{tag}/pred/{run}/{k}"] = coder.transform(res[run]["pred"][k])

    random.seed(seed)
    flatten_joint(raw.benchmark((img, pro, names), coder, n, repeats, K), "joint")
    random.seed(seed)
    flatten_joint(jfolds.benchmark((img[:half], pro[:half], names[:half]), (img[half:], pro[half:], names[half:]),
                                   coder, n, repeats, K), "jointfolds")
    np.savez_compressed(os.path.join(OUT, f"bench_{name}.npz"), **out)
    print("wrote bench", name, len(out), "arrays")


def main():
    os.makedirs(OUT, exist_ok=True)
    loss_case("b64_d128_k1", 64, 128, 1, 1.0, 1)
    loss_case("b256_d512_k1", 256, 512, 1, 1.0, 2)           # BASELINE config[0]
    loss_case("b256_d512_k4", 256, 512, 4, 1.0, 3)
    loss_case("b192_d256_k3_ls2.66", 192, 256, 3, 2.659, 4)  # ln(1/0.07)
    loss_case("b100_d72_k1_ls0", 100, 72, 1, 0.0, 5)         # ragged B and d

    def tiny(img, pro):      # near-zero-norm rows (eps path) + duplicated rows (ties)
        img[3] = 0.0
        pro[5] = 1e-20
        img[7] = img[8]
        pro[7] = pro[8]
    loss_case("b64_d64_edge", 64, 64, 2, 1.0, 6, tweak=tiny)

    siglip_case("b64_d128_k1", 64, 128, 1, 1.0, -10.0, 21)           # the reference's initial parameters
    siglip_case("b256_d256_k4", 256, 256, 4, 1.0, -10.0, 22)
    siglip_case("b192_d256_k3_ls2.3_b-4", 192, 256, 3, 2.3, -4.0, 23)  # a trained-looking temperature / bias
    siglip_case("b100_d72_k1_ls0_b0", 100, 72, 1, 0.0, 0.0, 24)      # ragged B and d, logits around 0
    siglip_case("b64_d64_edge", 64, 64, 2, 1.0, -10.0, 25, tweak=tiny)

    bench_case("c9_p40_d64_n4", 9, 40, 64, 4, 2, (1, 3, 5), 31)

    ann_case("g432_q300_d512_k9", 432, 300, 512, 9, 11, n_classes=27)
    ann_case("g96_q64_d64_k5_dup", 96, 64, 64, 5, 12, dup=True)
    ann_case("g200_q128_d128_k3_two", 200, 128, 128, 3, 13, two_mod=True)
    ann_case("g50_q40_d32_k1", 50, 40, 32, 1, 14)


if __name__ == "__main__":
    main()
