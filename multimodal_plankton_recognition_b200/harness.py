"""Few-shot retrieval benchmark of the reference's drivers on the GPU path.

`cross_benchmark` reproduces `benchmark()` of reference scripts/benchmark_cross.py:24-87 -- same
arguments, same use of Python's `random` stream for the gallery draw (so a caller that seeds it like
the reference's `main()` gets the same galleries), same nested result
``{run: {"pred": {k: {setup: class names}}, "true": class names}}`` -- with two differences in HOW:
the gallery index is `ANNClassifier` of this package (exact search on the tensor cores), and every
(gallery, query) set-up is searched ONCE at max(K) and voted for each k from the prefixes
(`predict_multi_k`) instead of once per k.  `cross_benchmark_folds` does the same for the
train-fold / test-fold variant (reference scripts/benchmark_cross_folds.py:24-85); `joint_benchmark`
/ `joint_benchmark_folds` are the single set-up "I+P gallery, (I, P) query" of
reference scripts/benchmark_raw.py:24-50 and scripts/benchmark_folds.py:24-51
(result ``{run: {"pred": {k: class names}, "true": class names}}``).
"""
from __future__ import annotations

import random

import numpy as np

from .ann import ANNClassifier

# the reference drivers' index settings (scripts/benchmark_cross.py:30-37); the exact index ignores
# the graph-construction knobs but takes `metric`
ANN_KWARGS = dict(n_neighbors=32, metric="euclidean", diversify_prob=0.0, pruning_degree_multiplier=3.0,
                  low_memory=False, random_state=0)

# gallery modality -> (set-up name, query modalities) in the drivers' order
_SETUPS = {
    "I": (("I - I", ("I",)), ("I - P", ("P",)), ("I - I+P", ("I", "P"))),
    "P": (("P - I", ("I",)), ("P - P", ("P",)), ("P - I+P", ("I", "P"))),
    "I+P": (("I+P - I", ("I",)), ("I+P - P", ("P",))),
}


def draw_gallery(labels, n):
    """n random members of every class, classes in sorted order, drawn with `random.sample` from the
    ascending index list of the class (reference scripts/benchmark_cross.py:14-21)."""
    labels = np.asarray(labels)
    positions = np.arange(len(labels))
    picked = []
    for cls in np.unique(labels):
        picked += random.sample(list(positions[labels == cls]), n)
    return np.array(picked)


def keep_frequent(data, coder, th):
    """Drop the classes with fewer than `th` samples (reference scripts/benchmark_cross.py:98-108)."""
    images, profiles, names = data
    label = coder.transform(names)
    ids, counts = np.unique(label, return_counts=True)
    keep = np.concatenate([np.where(label == c)[0] for c in ids[counts >= th]])
    return images[keep], profiles[keep], names[keep]


def _run_setups(gallery, gallery_labels, queries, coder, K, ann_kwargs):
    """gallery: {"I": [n,d], "P": [n,d]}; queries likewise -> {k: {setup: class names}}"""
    pred = {k: {} for k in K}
    for gal_kind, setups in _SETUPS.items():
        if gal_kind == "I+P":
            gx = np.concatenate((gallery["I"], gallery["P"]))
            gy = np.tile(gallery_labels, (2,))
        else:
            gx, gy = gallery[gal_kind], gallery_labels
        clf = ANNClassifier(gx, gy, **ann_kwargs)
        for name, kinds in setups:
            by_k = clf.predict_multi_k(*[queries[m] for m in kinds], ks=K, epsilon=.3)
            for k in K:
                pred[k][name] = coder.inverse_transform(by_k[k])
    return pred


def cross_benchmark(data, coder, n, repeats, K, **ann_overrides):
    """data = (images [N,d], profiles [N,d], class names [N]); per run: draw n per class as gallery, the
    rest are queries; 8 set-ups x every k in K."""
    images, profiles, names = data
    labels = coder.transform(names)
    everything = set(range(len(labels)))
    kw = {**ANN_KWARGS, **ann_overrides}
    results = {}
    for run in range(repeats):
        tr = draw_gallery(labels, n)
        te = list(everything - set(tr))
        pred = _run_setups({"I": images[tr], "P": profiles[tr]}, labels[tr], {"I": images[te], "P": profiles[te]},
                           coder, K, kw)
        results[run] = {"pred": pred, "true": coder.inverse_transform(labels[te])}
    return results


def cross_benchmark_folds(train, test, coder, n, repeats, K, **ann_overrides):
    """train / test = (images, profiles, class names) of two folds; per run the gallery is n per class of
    the train fold, the whole test fold is queried."""
    image_train, profile_train, name_train = train
    image_test, profile_test, name_test = test
    label_train, label_test = coder.transform(name_train), coder.transform(name_test)
    kw = {**ANN_KWARGS, **ann_overrides}
    results = {}
    for run in range(repeats):
        tr = draw_gallery(label_train, n)
        pred = _run_setups({"I": image_train[tr], "P": profile_train[tr]}, label_train[tr],
                           {"I": image_test, "P": profile_test}, coder, K, kw)
        results[run] = {"pred": pred, "true": coder.inverse_transform(label_test)}
    return results


def _joint(gallery_i, gallery_p, gallery_labels, query_i, query_p, coder, K, ann_kwargs):
    clf = ANNClassifier(np.concatenate((gallery_i, gallery_p)), np.tile(gallery_labels, (2,)), **ann_kwargs)
    by_k = clf.predict_multi_k(query_i, query_p, ks=K, epsilon=.3)
    return {k: coder.inverse_transform(by_k[k]) for k in K}


def joint_benchmark(data, coder, n, repeats, K, **ann_overrides):
    """Both modalities in the gallery and in the query; n per class drawn from `data`, the rest queried."""
    images, profiles, names = data
    labels = coder.transform(names)
    everything = set(range(len(labels)))
    kw = {**ANN_KWARGS, **ann_overrides}
    results = {}
    for run in range(repeats):
        tr = draw_gallery(labels, n)
        te = list(everything - set(tr))
        results[run] = {"pred": _joint(images[tr], profiles[tr], labels[tr], images[te], profiles[te], coder, K, kw),
                        "true": coder.inverse_transform(labels[te])}
    return results


def joint_benchmark_folds(train, test, coder, n, repeats, K, **ann_overrides):
    """Both modalities in the gallery (n per class of the train fold) and in the query (the test fold)."""
    image_train, profile_train, name_train = train
    image_test, profile_test, name_test = test
    label_train, label_test = coder.transform(name_train), coder.transform(name_test)
    kw = {**ANN_KWARGS, **ann_overrides}
    results = {}
    for run in range(repeats):
        tr = draw_gallery(label_train, n)
        results[run] = {"pred": _joint(image_train[tr], profile_train[tr], label_train[tr], image_test, profile_test,
                                       coder, K, kw),
                        "true": coder.inverse_transform(label_test)}
    return results
