"""Parity of the SHARDED CUDA path on one GPU: the ranks of `dist.sharded_fwd` / `dist.sharded_bwd` and of
`dist.ShardedANNClassifier` are run one after the other on cuda:0 through the same `ops.*` / `ann.*` calls
(the C ABI), with the collectives replaced by what they compute (concatenation for the all-gathers, a sum
for the all-reduces).  This is what exercises, on hardware, the kernels with `row_offset != 0`,
`n_rows < n_cols` (owned row block against the whole batch), partial column sums, `gallery_offset != 0`
and `plk_topk_merge` -- compared with the UNSHARDED oracle on the concatenated batch
(reference src/coordination.py:26-47, src/ann.py:15-25).  The collectives themselves are covered by the gloo
world-2 tests (tests/test_dist_cpu.py) and by the parity objects bench.py emits at N > 1."""
import numpy as np
import pytest
import torch

from oracle import ann as oann
from oracle import infonce as oinf

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "bf16": 2e-3, "fp16": 2e-3}


def _pairs(B, d, seed):
    r = np.random.default_rng(seed)
    cent = r.standard_normal((27, d))
    lab = r.integers(0, 27, B)
    z = r.standard_normal((B, d))
    img = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    pro = (cent[lab] + 0.5 * z + 0.3 * r.standard_normal((B, d))).astype(np.float32)
    return img, pro


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def sharded_step_on_one_gpu(img, pro, ls_val, buckets, counts, precision, grad_out=1.0):
    """The general (not bucket-aligned) branch of dist.sharded_fwd + dist.sharded_bwd for ranks owning
    `counts[r]` consecutive rows each.  -> (loss, d_image, d_profile, d_logit_scale) of the global batch."""
    from multimodal_plankton_recognition_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.MODES[precision]
    B, d = img.shape
    bs = B // buckets
    offs = np.concatenate(([0], np.cumsum(counts)))
    assert offs[-1] == B
    ls = torch.tensor(float(ls_val), device=dev)
    go = torch.full((1,), float(grad_out), device=dev)
    ranks = []
    for r, n in enumerate(counts):                      # local normalisation (+ zero fill of the accumulators)
        o = int(offs[r])
        x = torch.tensor(img[o:o + n], device=dev)
        y = torch.tensor(pro[o:o + n], device=dev)
        st4 = torch.empty((4, n), device=dev)
        rs = torch.empty(n, device=dev)
        cs_part = torch.empty(B, device=dev)
        u, v = ops.l2norm_pair(x, y, mode, st4, rs, cs_part)
        ranks.append(dict(o=o, n=n, x=x, y=y, st4=st4, rs=rs, cs_part=cs_part, u=u, v=v,
                          dg=torch.empty(n, device=dev)))
    u_all = torch.cat([k["u"] for k in ranks])          # all_gather_into_tensor
    v_all = torch.cat([k["v"] for k in ranks])
    for k in ranks:                                     # owned rows x all columns
        ops.infonce_fwd_local(k["u"], v_all, mode, d, k["o"], bs, ls, k["rs"], k["cs_part"], k["dg"],
                              sums_zeroed=True)
    cs_all = torch.stack([k["cs_part"] for k in ranks]).sum(0)      # all_reduce(SUM) of the partial column sums
    rs_all = torch.cat([k["rs"] for k in ranks])                    # all_gather of the row sums
    loss = 0.0
    dls = 0.0
    dxs, dys = [], []
    for k in ranks:
        o, n = k["o"], k["n"]
        part, aux = ops.infonce_loss_local(k["rs"], cs_all[o:o + n], k["dg"], ls, B)
        loss += float(part)                                         # all_reduce(SUM) of the scalar
        rs_own, cs_own = rs_all[o:o + n], cs_all[o:o + n]
        gs = aux[1:]
        acc_x, acc_y = ops.infonce_grad_pair_local(k["u"], v_all, k["v"], u_all, mode, d, o, bs, ls, rs_own, cs_all,
                                                   cs_own, rs_all, gs)
        idx, nx, idy, ny = k["st4"].unbind(0)
        dx, dy, dl = ops.infonce_grad_finish_pair(acc_x, acc_y, k["x"], k["y"], (idx, nx), (idy, ny), k["dg"], rs_own,
                                                  cs_own, ls, go, go, B, gs, aux[0:1])
        dxs.append(dx)
        dys.append(dy)
        dls += float(dl)
    torch.cuda.synchronize()
    return loss, torch.cat(dxs).cpu().numpy(), torch.cat(dys).cpu().numpy(), dls


def _check(got, ref, tol):
    loss, dx, dy, dls = got
    assert abs(loss - ref["loss"]) / abs(ref["loss"]) < tol, ("loss", loss, ref["loss"])
    for name, a, b in (("d_image", dx, ref["d_image"]), ("d_profile", dy, ref["d_profile"])):
        assert _rel(a, b) < tol, (name, "max", _rel(a, b))
        assert _rel_l2(a, b) < tol, (name, "l2", _rel_l2(a, b))
    assert abs(dls - ref["d_logit_scale"]) <= tol * max(abs(ref["d_logit_scale"]), 1e-3), \
        ("d_logit_scale", dls, ref["d_logit_scale"])


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("B,d,buckets,counts", [
    (4096, 256, 1, [512] * 8),                 # BASELINE config[1] shape over 8 ranks, one bucket
    (4096, 256, 8, [1024] * 4),                # buckets cut across nothing, ranks hold 2 buckets each
    (1536, 192, 3, [384] * 4),                 # a bucket (512 rows) spans rank boundaries
    (1000, 200, 1, [250] * 4),                 # nothing a multiple of the tile
    (777, 96, 1, [300, 77, 400]),              # ragged shards
    (1280, 256, 1, [640, 384, 256]),           # paired-CTA backward: 5 / 3 / 2 row blocks (phantom block pads the odd ones)
    (900, 128, 1, [300, 600]),                 # paired-CTA backward at d = 128, row_offset not a multiple of the tile
])
def test_row_sharded_loss_matches_unsharded_oracle(B, d, buckets, counts, precision):
    img, pro = _pairs(B, d, B + d)
    ref = oinf.clip_loss_closed_form(img, pro, 1.0, buckets, grad_out=0.75)
    _check(sharded_step_on_one_gpu(img, pro, 1.0, buckets, counts, precision, grad_out=0.75), ref, TOL[precision])


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_c3_shape_row_block(precision):
    """BASELINE config[2] geometry at a size the CPU oracle finishes: d = 512, global batch 8192, four
    ranks of 2048 rows -- the streaming backward `infonce_grad_tc<8,4,*>` with n_rows < n_cols, partial
    column sums, row_offset up to 6144."""
    B, d = 8192, 512
    img, pro = _pairs(B, d, 5)
    ref = oinf.clip_loss_closed_form(img, pro, 1.0, 1)
    _check(sharded_step_on_one_gpu(img, pro, 1.0, 1, [2048] * 4, precision), ref, TOL[precision])


def test_c3_full_size_sharded_equals_unsharded_kernels():
    """BASELINE config[2] at FULL size (global batch 32768, d = 512, 8 ranks of 4096 rows), where the B x B
    oracle does not fit: the rank-sharded bf16 path must reproduce the single-GPU bf16 path on the same
    batch (which the smaller cases pin to the oracle) -- same operands, same logits, sums in another order."""
    from multimodal_plankton_recognition_b200 import CLIPLoss
    B, d, R = 32768, 512, 8
    img, pro = _pairs(B, d, 9)
    got = sharded_step_on_one_gpu(img, pro, 1.0, 1, [B // R] * R, "bf16")
    dev = torch.device("cuda:0")
    mod = CLIPLoss(precision="bf16").to(dev)
    x = torch.tensor(img, device=dev, requires_grad=True)
    y = torch.tensor(pro, device=dev, requires_grad=True)
    loss = mod(image_emb=x, profile_emb=y, buckets=1)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(got[0] - float(loss)) <= 1e-5 * abs(float(loss))
    # the bf16 rounding of G flips where the sum-exp atomics land in another order: ~1e-5 relative
    assert _rel(got[1], x.grad.cpu().numpy()) < 2e-4 and _rel(got[2], y.grad.cpu().numpy()) < 2e-4
    assert abs(got[3] - float(mod.logit_scale.grad)) <= 2e-4 * max(abs(float(mod.logit_scale.grad)), 1e-3)
    # size-independent properties of the loss on this data: finite, below log(B) + margin
    assert np.isfinite(got[0]) and 0.0 < got[0] < np.log(B) + 1.0
    assert np.isfinite(got[1]).all() and np.isfinite(got[2]).all()


# ---------------------------------------------------------------------------------------------------------
# retrieval: gallery shards with gallery_offset + plk_topk_merge vs the unsharded oracle
# ---------------------------------------------------------------------------------------------------------
def _clustered(n, d, n_classes, seed, noise=0.9):
    r = np.random.default_rng(seed)
    cent = r.standard_normal((n_classes, d))
    lab = r.integers(0, n_classes, n)
    e = cent[lab] + noise * r.standard_normal((n, d))
    return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32), lab


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ng,nq,d,k,shards", [(6000, 500, 512, 10, [1500] * 4), (2500, 300, 200, 9, [7, 1200, 993, 300]),
                                               (40000, 256, 256, 10, [5000] * 8)])
def test_gallery_sharded_search_matches_unsharded_oracle(ng, nq, d, k, shards, precision):
    """dist.ShardedANNClassifier.search_device / predict with the all-gather replaced by a concatenation."""
    from multimodal_plankton_recognition_b200 import ann
    dev = torch.device("cuda:0")
    gal, yg = _clustered(ng, d, 27, 1)
    q, _ = _clustered(nq, d, 27, 2, noise=1.1)
    want = oann.OracleANNClassifier(gal, yg)
    wi, wd = want.kneighbors(q, k=k)[0]
    q32 = torch.tensor(q, device=dev)
    offs = np.concatenate(([0], np.cumsum(shards)))
    lists_i, lists_d = [], []
    for r, n in enumerate(shards):
        o = int(offs[r])
        index = ann.GpuExactIndex(gal[o:o + n], precision=precision, device=dev, gallery_offset=o)
        k_loc = min(k, n)
        i, dd = index.search_device(q32, k_loc)
        if k_loc < k:          # short shard: pad with empty slots (dist.ShardedANNClassifier.search_device)
            i = torch.cat((i, torch.full((nq, k - k_loc), -1, device=dev, dtype=torch.int32)), 1)
            dd = torch.cat((dd, torch.full((nq, k - k_loc), float("inf"), device=dev)), 1)
        assert int(i[i >= 0].min()) >= o and int(i.max()) < o + n       # global indices of THIS shard
        lists_i.append(i)
        lists_d.append(dd)
    cand_i = torch.cat(lists_i, 1).contiguous()          # all_gather + permute of dist.merge_shard_results
    cand_d = torch.cat(lists_d, 1).contiguous()
    gi, gd = ann.topk_merge_device(cand_i, cand_d, k)
    labels = torch.tensor(yg.astype(np.int64), device=dev)
    pred = ann.knn_vote_device(gi, gd, labels).cpu().numpy()
    gi, gd = gi.cpu().numpy(), gd.cpu().numpy()
    np.testing.assert_allclose(gd, wd, rtol=1e-6, atol=0)
    mism = gi != wi
    assert mism.mean() < 1e-3 and np.abs(gd[mism] - wd[mism]).max(initial=0) < 1e-6
    np.testing.assert_array_equal(pred, want.predict(q, k=k))
