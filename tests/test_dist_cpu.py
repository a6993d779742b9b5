"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-block sharded loss with all-gather
/ all-reduce, the bucket-aligned no-exchange path, DDP gradient scaling, and gallery-sharded
retrieval with candidate merge.  The CUDA entry points are replaced by the fp64 stand-ins of
tests/_emul.py; the oracle on the concatenated batch / gallery is the ground truth."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _patch():
    import _emul
    from multimodal_plankton_recognition_b200 import ann, ops
    for name in ("l2norm", "l2norm_pair", "infonce_fwd_local", "infonce_loss_local", "infonce_grad_pair_local",
                 "infonce_grad_finish", "infonce_grad_finish_pair", "infonce_dls", "clip_loss_forward_state",
                 "clip_loss_backward_state"):
        setattr(ops, name, getattr(_emul, name))
    ann.GpuExactIndex = _emul.CpuExactIndex
    ann.topk_merge_device = _emul.topk_merge_device
    ann.knn_vote_device = _emul.knn_vote_device


def _worker(rank, port, case, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        _patch()
        q.put((rank, case(rank)))
    finally:
        dist.destroy_process_group()


def _run(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, case, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=180) for _ in range(WORLD))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return out


def _data(B=96, d=24, seed=0):
    r = np.random.default_rng(seed)
    img = r.standard_normal((B, d)).astype(np.float32)
    pro = (img + 0.8 * r.standard_normal((B, d))).astype(np.float32)
    return img, pro


def _loss_worker(buckets, grad_scale, rank):
    from multimodal_plankton_recognition_b200 import dist as pdist, ops
    img, pro = _data()
    n = img.shape[0] // WORLD
    x = torch.tensor(img[rank * n:(rank + 1) * n], requires_grad=True)
    y = torch.tensor(pro[rank * n:(rank + 1) * n], requires_grad=True)
    ls = torch.tensor(1.3, requires_grad=True)
    loss = pdist.sharded_clip_loss(x, y, ls, buckets, ops.PLK_F32, None, grad_scale)
    (loss * 2.0).backward()
    return float(loss.detach()), x.grad.numpy(), y.grad.numpy(), float(ls.grad)


def _loss_case(buckets, grad_scale):
    import functools
    return functools.partial(_loss_worker, buckets, grad_scale)


@pytest.mark.parametrize("buckets", [1, 2, 3, 4])   # 1,3: real exchange; 2,4: bucket-aligned, no exchange
def test_sharded_loss_matches_oracle_on_concatenated_batch(buckets):
    from oracle import infonce as oinf
    img, pro = _data()
    ref = oinf.clip_loss_closed_form(img, pro, 1.3, buckets, grad_out=2.0)
    out = _run(_loss_case(buckets, "none"))
    n = img.shape[0] // WORLD
    for r in range(WORLD):
        loss, dx, dy, dls = out[r]
        assert loss == pytest.approx(ref["loss"], rel=1e-6)          # same global loss on every rank
        np.testing.assert_allclose(dx, ref["d_image"][r * n:(r + 1) * n], rtol=2e-4, atol=1e-8)
        np.testing.assert_allclose(dy, ref["d_profile"][r * n:(r + 1) * n], rtol=2e-4, atol=1e-8)
        assert dls == pytest.approx(ref["d_logit_scale"], rel=1e-5)  # all-reduced: identical on every rank


def test_ddp_gradient_scaling():
    from oracle import infonce as oinf
    img, pro = _data()
    ref = oinf.clip_loss_closed_form(img, pro, 1.3, 1, grad_out=2.0)
    out = _run(_loss_case(1, "ddp"))
    n = img.shape[0] // WORLD
    # DDP averages parameter grads over ranks: local embedding grads are pre-multiplied by the world size
    np.testing.assert_allclose(out[1][1], WORLD * ref["d_image"][n:], rtol=2e-4, atol=1e-8)
    assert out[0][3] == pytest.approx(ref["d_logit_scale"], rel=1e-5)


def _module_worker(d, buckets, rank):
    from multimodal_plankton_recognition_b200 import CLIPLoss
    img, pro = _data(96, d, seed=3)
    n = img.shape[0] // WORLD
    x = torch.tensor(img[rank * n:(rank + 1) * n], requires_grad=True)
    y = torch.tensor(pro[rank * n:(rank + 1) * n], requires_grad=True)
    mod = CLIPLoss(precision="fp32", sharded=True)
    loss = mod(image_emb=x, profile_emb=y, buckets=buckets)
    loss.backward()
    return float(loss.detach()), x.grad.numpy(), y.grad.numpy(), float(mod.logit_scale.grad)


@pytest.mark.parametrize("d,buckets", [(128, 2), (24, 2), (128, 3)])   # in-kernel DDP factor / host factor / not aligned
def test_sharded_module_path_with_ddp_scaling(d, buckets):
    """`CLIPLoss(sharded=True)` end to end (module -> dist -> composite calls): global loss on every rank,
    embedding gradients pre-multiplied by the world size (DDP averages them back), `d logit_scale` global."""
    import functools
    from oracle import infonce as oinf
    img, pro = _data(96, d, seed=3)
    ref = oinf.clip_loss_closed_form(img, pro, 1.0, buckets)
    out = _run(functools.partial(_module_worker, d, buckets))
    n = img.shape[0] // WORLD
    for r in range(WORLD):
        loss, dx, dy, dls = out[r]
        assert loss == pytest.approx(ref["loss"], rel=1e-6)
        np.testing.assert_allclose(dx, WORLD * ref["d_image"][r * n:(r + 1) * n], rtol=2e-4, atol=1e-8)
        np.testing.assert_allclose(dy, WORLD * ref["d_profile"][r * n:(r + 1) * n], rtol=2e-4, atol=1e-8)
        assert dls == pytest.approx(ref["d_logit_scale"], rel=1e-5)


def _ann_case(rank):
    from multimodal_plankton_recognition_b200.dist import ShardedANNClassifier
    r = np.random.default_rng(5)
    gal = r.standard_normal((150, 16)).astype(np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    lab = r.integers(0, 5, 150)
    q = r.standard_normal((40, 16)).astype(np.float32)
    cut = 60                                             # uneven shards: 60 + 90 rows
    sl = slice(0, cut) if rank == 0 else slice(cut, 150)
    clf = ShardedANNClassifier(gal[sl], lab[sl], plk_device="cpu")
    (idx, dd), = clf.kneighbors(q, k=7)
    return idx, dd, clf.predict(q, k=7), clf.predict(q, q[::-1].copy(), k=3)


def test_sharded_retrieval_matches_unsharded_oracle():
    from oracle import ann as oann
    r = np.random.default_rng(5)
    gal = r.standard_normal((150, 16)).astype(np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    lab = r.integers(0, 5, 150)
    q = r.standard_normal((40, 16)).astype(np.float32)
    ref = oann.OracleANNClassifier(gal, lab)
    (wi, wd), = ref.kneighbors(q, k=7)
    out = _run(_ann_case)
    for rk in range(WORLD):
        idx, dd, pred, pred2 = out[rk]
        np.testing.assert_array_equal(idx, wi)
        np.testing.assert_array_equal(dd, wd)
        np.testing.assert_array_equal(pred, ref.predict(q, k=7))
        np.testing.assert_array_equal(pred2, ref.predict(q, q[::-1].copy(), k=3))
