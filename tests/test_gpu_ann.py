"""Parity of the CUDA retrieval / k-NN path (through the C ABI) against the oracle and the
golden vectors produced by the reference's own ANNClassifier code."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files
from oracle import ann as oann

pytestmark = pytest.mark.gpu
KW = dict(n_neighbors=32, metric="euclidean", diversify_prob=0.0, pruning_degree_multiplier=3.0,
          low_memory=False, random_state=0)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("path", golden_files("ann_"), ids=os.path.basename)
def test_golden_reference_vectors(path, precision):
    from multimodal_plankton_recognition_b200 import ANNClassifier
    g = np.load(path)
    k = int(g["k"])
    clf = ANNClassifier(g["gallery"], g["labels"], plk_precision=precision, **KW)
    qs = [g[f"query{m}"] for m in range(2) if f"query{m}" in g]
    nb = clf.kneighbors(*qs, k=k, epsilon=.3)
    assert len(nb) == len(qs)
    for m, (i, dd) in enumerate(nb):
        assert i.dtype == np.int32 and dd.dtype == np.float32 and i.shape == (len(qs[m]), k)
        np.testing.assert_array_equal(dd, g[f"dist{m}"])        # exact distances, bit for bit
        same = i == g[f"idx{m}"]
        # indices may only differ inside groups of exactly equal distance (duplicated gallery rows)
        assert same.all() or (g[f"dist{m}"][~same] == dd[~same]).all()
    np.testing.assert_array_equal(clf.predict(*qs, k=k, epsilon=.3), g["pred"])


def _clustered(n, d, n_classes, seed, noise=0.9):
    r = np.random.default_rng(seed)
    cent = r.standard_normal((n_classes, d))
    lab = r.integers(0, n_classes, n)
    e = cent[lab] + noise * r.standard_normal((n, d))
    return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32), lab


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ng,nq,d,k", [(5000, 700, 512, 10), (1234, 333, 200, 9), (300, 50, 72, 1),
                                        (20000, 256, 256, 10)])
def test_against_oracle(ng, nq, d, k, precision):
    from multimodal_plankton_recognition_b200 import ANNClassifier
    gal, yg = _clustered(ng, d, 27, 1)
    q, _ = _clustered(nq, d, 27, 2, noise=1.1)
    want = oann.OracleANNClassifier(gal, yg)
    wi, wd = want.kneighbors(q, k=k)[0]
    clf = ANNClassifier(gal, yg, plk_precision=precision, **KW)
    gi, gd = clf.kneighbors(q, k=k, epsilon=.3)[0]
    np.testing.assert_allclose(gd, wd, rtol=1e-6, atol=0)
    mism = gi != wi
    # a differing index is only acceptable at a (near-)tie: gap below the fp tolerance
    assert mism.mean() < 1e-3 and np.abs(gd[mism] - wd[mism]).max(initial=0) < 1e-6
    np.testing.assert_array_equal(clf.predict(q, k=k, epsilon=.3), want.predict(q, k=k))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sliced_upload_equals_one_upload(precision):
    """Large host query matrices are uploaded in slices under the search of the previous slice
    (GpuExactIndex.search_host): same neighbours, distances and labels as one upload, ragged last slice included."""
    from multimodal_plankton_recognition_b200 import ANNClassifier
    gal, yg = _clustered(3000, 128, 27, 1)
    q, _ = _clustered(1111, 128, 27, 2, noise=1.1)
    clf = ANNClassifier(gal, yg, plk_precision=precision, **KW)
    wi, wd = clf.kneighbors(q, k=10)[0]
    wl = clf.predict(q, k=10)
    clf.index.PIPELINE_MIN_BYTES = 0
    clf.index.PIPELINE_SLICE_BYTES = 256 * 128 * 4 + 7       # 5 slices of 256 queries, the last one ragged (87)
    gi, gd = clf.kneighbors(q, k=10)[0]
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gd, wd)
    np.testing.assert_array_equal(clf.predict(q, k=10), wl)
    np.testing.assert_array_equal(clf.predict_multi_k(q, ks=(1, 5, 10))[10], wl)


def test_two_modalities_and_fold_setups():
    """The 8 set-ups of reference scripts/benchmark_cross.py:57-86 on a synthetic fold: labels identical to
    the oracle for k in (1,3,5,7,9) and the BASELINE k=10."""
    from multimodal_plankton_recognition_b200 import ANNClassifier
    img, lab = _clustered(1500, 512, 27, 5)
    pro = (img + 0.05 * np.random.default_rng(6).standard_normal(img.shape)).astype(np.float32)
    pro /= np.linalg.norm(pro, axis=1, keepdims=True)
    tr, te = np.arange(0, 432), np.arange(432, 1500)
    setups = {
        "I": (img[tr], lab[tr]), "P": (pro[tr], lab[tr]),
        "I+P": (np.concatenate((img[tr], pro[tr])), np.tile(lab[tr], 2)),
    }
    for name, (gx, gy) in setups.items():
        ora = oann.OracleANNClassifier(gx, gy)
        clf = ANNClassifier(gx, gy, plk_precision="fp32", **KW)
        queries = [(img[te],), (pro[te],)] + ([(img[te], pro[te])] if name != "I+P" else [])
        for X in queries:
            for k in (1, 3, 5, 7, 9, 10):
                np.testing.assert_array_equal(clf.predict(*X, k=k, epsilon=.3), ora.predict(*X, k=k))


def test_large_gallery_properties_bf16():
    """Size-independent checks at a size the oracle cannot finish: distances ascending, equal to
    the recomputed distance of the returned index, and the k-th distance is a true lower bound
    for a random sample of other gallery rows."""
    from multimodal_plankton_recognition_b200.ann import GpuExactIndex
    ng, nq, d, k = 300_000, 4096, 512, 10
    gen = torch.Generator(device="cuda").manual_seed(0)
    gal = torch.nn.functional.normalize(torch.randn(ng, d, device="cuda", generator=gen))
    q = torch.nn.functional.normalize(torch.randn(nq, d, device="cuda", generator=gen) + 0.3 * gal[:nq])
    index = GpuExactIndex(gal.cpu().numpy(), precision="bf16")
    idx, dist = index.search_device(q, k)
    torch.cuda.synchronize()
    assert (dist[:, 1:] >= dist[:, :-1]).all()
    rec = (q[:, None, :].double() - gal[idx.long()].double()).pow(2).sum(-1).sqrt().float()
    assert torch.equal(rec, dist) or (rec - dist).abs().max() < 1e-6
    # brute force on a slice of the queries (fp32 on the GPU, independent of libplk)
    sub = slice(0, 256)
    full = torch.cdist(q[sub].double(), gal.double()).float()
    best = full.topk(k, dim=1, largest=False)
    assert (best.values - dist[sub]).abs().max() < 1e-6
    assert (best.indices.int() == idx[sub]).float().mean() > 0.999


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_predict_multi_k_equals_one_search_per_k(precision):
    """One search at k_max + prefix votes gives the labels of the reference's per-k
    loop (scripts/benchmark_cross.py:57-64), for one and two query modalities, incl. k > gallery size."""
    from multimodal_plankton_recognition_b200 import ANNClassifier
    img, lab = _clustered(900, 256, 27, 15)
    pro = (img + 0.05 * np.random.default_rng(16).standard_normal(img.shape)).astype(np.float32)
    pro /= np.linalg.norm(pro, axis=1, keepdims=True)
    tr, te = np.arange(0, 300), np.arange(300, 900)
    ora = oann.OracleANNClassifier(img[tr], lab[tr])
    clf = ANNClassifier(img[tr], lab[tr], plk_precision=precision, **KW)
    ks = (1, 3, 5, 7, 9, 10)
    for X in ((img[te],), (pro[te],), (img[te], pro[te])):
        got = clf.predict_multi_k(*X, ks=ks, epsilon=.3)
        assert sorted(got) == sorted(ks)
        for k in ks:
            np.testing.assert_array_equal(got[k], clf.predict(*X, k=k, epsilon=.3))
            np.testing.assert_array_equal(got[k], ora.predict(*X, k=k))
    small = ANNClassifier(img[:5], lab[:5], plk_precision=precision, **KW)
    got = small.predict_multi_k(img[te][:17], ks=(3, 9))
    np.testing.assert_array_equal(got[9], small.predict(img[te][:17], k=9))
    with pytest.raises(ValueError):
        clf.predict_multi_k(img[te], ks=())


@pytest.mark.parametrize("precision", ["bf16", "fp16", "fp32"])
def test_reference_driver_k_values_up_to_51(precision):
    """K = (1, 3, 9, 15, 31, 51) of reference scripts/benchmark_raw.py:82 / benchmark_folds.py:70 on the DEFAULT
    16-bit index: k > 26 is served by the fp32 candidate search (the tensor-core kernel keeps 32 candidates
    per query), with the same exact re-score -- neighbours, distances and labels equal to the oracle."""
    from multimodal_plankton_recognition_b200 import ANNClassifier, harness
    gal, yg = _clustered(3000, 128, 27, 21)
    q, _ = _clustered(400, 128, 27, 22, noise=1.1)
    ora = oann.OracleANNClassifier(gal, yg)
    clf = ANNClassifier(gal, yg, plk_precision=precision, **KW)
    K = (1, 3, 9, 15, 31, 51)
    for k in (26, 27, 31, 51, 58):
        wi, wd = ora.kneighbors(q, k=k)[0]
        gi, gd = clf.kneighbors(q, k=k, epsilon=.3)[0]
        np.testing.assert_allclose(gd, wd, rtol=1e-6, atol=0)
        mism = gi != wi
        assert mism.mean() < 1e-3 and np.abs(gd[mism] - wd[mism]).max(initial=0) < 1e-6
    got = clf.predict_multi_k(q, ks=K, epsilon=.3)
    for k in K:
        np.testing.assert_array_equal(got[k], ora.predict(q, k=k))
    with pytest.raises(ValueError):
        clf.kneighbors(q, k=59)
    assert harness is not None


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", golden_files("bench_"), ids=os.path.basename)
def test_benchmark_harness_matches_reference_drivers(path, precision):
    """harness.cross_benchmark / cross_benchmark_folds against the label arrays the reference's own
    drivers produced (scripts/benchmark_cross.py, scripts/benchmark_cross_folds.py run by
    oracle/make_golden.py): same `random` stream -> same galleries -> identical predictions for every
    run, k and set-up."""
    import random
    from sklearn.preprocessing import LabelEncoder
    from multimodal_plankton_recognition_b200 import harness
    g = np.load(path)
    img, pro, names = g["image"], g["profile"], g["names"]
    n, repeats, K, seed = int(g["n"]), int(g["repeats"]), tuple(int(k) for k in g["K"]), int(g["seed"])
    coder = LabelEncoder().fit(names)

    def check(res, tag):
        assert sorted(res) == list(range(repeats))
        n_checked = 0
        for run in res:
            np.testing.assert_array_equal(coder.transform(res[run]["true"]), g[f"{tag}/true/{run}"])
            assert sorted(res[run]["pred"]) == sorted(K)
            for k in K:
                assert len(res[run]["pred"][k]) == 8
                for setup, pred in res[run]["pred"][k].items():
                    np.testing.assert_array_equal(coder.transform(pred), g[f"{tag}/pred/{run}/{k}/{setup}"],
                                                  err_msg=f"{tag} run {run} k {k} {setup}")
                    n_checked += 1
        assert n_checked == repeats * len(K) * 8

    random.seed(seed)
    check(harness.cross_benchmark((img, pro, names), coder, n, repeats, K, plk_precision=precision), "cross")
    half = len(names) // 2
    random.seed(seed)
    check(harness.cross_benchmark_folds((img[:half], pro[:half], names[:half]), (img[half:], pro[half:], names[half:]),
                                        coder, n, repeats, K, plk_precision=precision), "folds")

    def check_joint(res, tag):
        for run in range(repeats):
            np.testing.assert_array_equal(coder.transform(res[run]["true"]), g[f"{tag}/true/{run}"])
            for k in K:
                np.testing.assert_array_equal(coder.transform(res[run]["pred"][k]), g[f"{tag}/pred/{run}/{k}"],
                                              err_msg=f"{tag} run {run} k {k}")

    random.seed(seed)
    check_joint(harness.joint_benchmark((img, pro, names), coder, n, repeats, K, plk_precision=precision), "joint")
    random.seed(seed)
    check_joint(harness.joint_benchmark_folds((img[:half], pro[:half], names[:half]),
                                              (img[half:], pro[half:], names[half:]), coder, n, repeats, K,
                                              plk_precision=precision), "jointfolds")
    kept = harness.keep_frequent((img, pro, names), coder, 40)      # every class has 40 samples: all kept, grouped by class
    assert len(kept[2]) == len(names) and (np.diff(coder.transform(kept[2])) >= 0).all()
