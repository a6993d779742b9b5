"""torchrun --nproc-per-node N tools/dist_check.py : sharded loss / retrieval on N GPUs vs the
unsharded single-GPU run of the same global problem (rank 0 computes both)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from multimodal_plankton_recognition_b200 import ANNClassifier, CLIPLoss, synth
from multimodal_plankton_recognition_b200.dist import ShardedANNClassifier

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for prec, tol in (("fp32", 2e-5), ("bf16", 2e-3)):
    for B, d, buckets in ((2048, 256, 1), (1024, 512, 1), (1536, 192, 3), (1024, 128, 2 * world)):
        img, pro, _ = synth.pairs(B, d, 77, "cpu")
        n = B // world
        x = img[rank * n:(rank + 1) * n].to(dev).requires_grad_()
        y = pro[rank * n:(rank + 1) * n].to(dev).requires_grad_()
        mod = CLIPLoss(precision=prec, sharded=True).to(dev)
        from multimodal_plankton_recognition_b200 import dist as pdist, ops
        loss = pdist.sharded_clip_loss(x, y, mod.logit_scale, buckets, ops.MODES[prec], None, "none")
        loss.backward()
        ref = CLIPLoss(precision="fp32").to(dev)
        xf, yf = img.to(dev).requires_grad_(), pro.to(dev).requires_grad_()
        lref = ref(image_emb=xf, profile_emb=yf, buckets=buckets)
        lref.backward()
        gx = xf.grad[rank * n:(rank + 1) * n]
        e_loss = abs(float(loss) - float(lref)) / abs(float(lref))
        e_gx = float((x.grad - gx).abs().max() / gx.abs().max())
        e_ls = abs(float(mod.logit_scale.grad) - float(ref.logit_scale.grad)) / max(abs(float(ref.logit_scale.grad)), 1e-6)
        good = max(e_loss, e_gx, e_ls) < tol
        ok &= good
        if rank == 0:
            print(f"loss[{prec}] B={B} d={d} bk={buckets}: loss {e_loss:.1e} dI {e_gx:.1e} dls {e_ls:.1e} {'OK' if good else 'FAIL'}", flush=True)

# fused cross-GPU scalar exchange (NVLink peer memory inside the gradient-tail kernel) vs NCCL
try:
    from multimodal_plankton_recognition_b200 import dist as pdist, ops
    xg = pdist.XGpuScalars(dev)
    img, pro, _ = synth.pairs(512 * world, 256, 5 + rank, dev)
    ls = torch.ones((), device=dev)
    go = torch.ones(1, device=dev)
    for rep in range(5):   # several epochs: parity slots are reused
        loss_p, state = pdist.sharded_fwd(img, pro, ls, world, ops.MODES["bf16"], None, reduce_scalars=False)
        want_loss = loss_p.clone()
        dx, dy, dls_p = pdist.sharded_bwd(state, go, "none", reduce_scalars=False, xgpu=xg)
        want = torch.stack((want_loss, dls_p.clone()))
        dist.all_reduce(want)
        torch.cuda.synchronize()
        err = float((xg.out2 - want).abs().max() / want.abs().max())
        good = err < 1e-6
        ok &= good
    if rank == 0:
        print(f"fused xgpu scalars: out2={xg.out2.tolist()} nccl={want.tolist()} {'OK' if good else 'FAIL'}", flush=True)
except Exception as e:
    ok = False
    print(f"rank {rank}: fused xgpu exchange failed: {e!r}", flush=True)

# the module path: CLIPLoss(sharded=True) -- bucket-aligned steps run without a collective launch (the loss
# kernel and the gradient-tail kernel exchange the scalars over peer memory); three steps on one module
for prec, tol in (("fp32", 2e-5), ("bf16", 2e-3)):
    mod = CLIPLoss(precision=prec, sharded=True).to(dev)
    ref = CLIPLoss(precision="fp32").to(dev)
    for step, (B, d, buckets) in enumerate(((1024, 256, world), (2048, 128, 2 * world), (1024, 256, 1))):
        img, pro, _ = synth.pairs(B, d, 90 + step, "cpu")
        n = B // world
        x = img[rank * n:(rank + 1) * n].to(dev).requires_grad_()
        y = pro[rank * n:(rank + 1) * n].to(dev).requires_grad_()
        mod.logit_scale.grad = None
        ref.logit_scale.grad = None
        loss = mod(image_emb=x, profile_emb=y, buckets=buckets)
        loss.backward()
        xf, yf = img.to(dev).requires_grad_(), pro.to(dev).requires_grad_()
        lref = ref(image_emb=xf, profile_emb=yf, buckets=buckets)
        lref.backward()
        gx = xf.grad[rank * n:(rank + 1) * n] * world          # grad_scale="ddp": pre-multiplied by the world size
        e_loss = abs(float(loss.detach()) - float(lref.detach())) / abs(float(lref.detach()))
        e_gx = float((x.grad - gx).abs().max() / gx.abs().max())
        e_ls = abs(float(mod.logit_scale.grad) - float(ref.logit_scale.grad)) / max(abs(float(ref.logit_scale.grad)), 1e-6)
        good = max(e_loss, e_gx, e_ls) < tol
        ok &= good
        if rank == 0:
            how = "peer memory" if (mod._xgpu is not None and buckets % world == 0) else "NCCL"
            print(f"module[{prec}] B={B} d={d} bk={buckets} ({how}): loss {e_loss:.1e} dI {e_gx:.1e} dls {e_ls:.1e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)

gal, lab = synth.unit_embeddings(40000, 256, 3, "cpu", 1)
q, _ = synth.unit_embeddings(2000, 256, 4, "cpu", 0)
shard = 40000 // world
sl = slice(rank * shard, (rank + 1) * shard if rank < world - 1 else 40000)
clf = ShardedANNClassifier(gal[sl].numpy(), lab[sl].numpy(), plk_precision="bf16")
(idx, dd), = clf.kneighbors(q.numpy(), k=10)
pred = clf.predict(q.numpy(), k=10)
if rank == 0:
    one = ANNClassifier(gal.numpy(), lab.numpy(), plk_precision="fp32")
    (wi, wd), = one.kneighbors(q.numpy(), k=10)
    good = np.array_equal(dd, wd) and (idx == wi).mean() > 0.999 and np.array_equal(pred, one.predict(q.numpy(), k=10))
    ok &= good
    print(f"retrieval sharded over {world}: dist equal {np.array_equal(dd, wd)} idx match {(idx == wi).mean():.4f} "
          f"labels equal {np.array_equal(pred, one.predict(q.numpy(), k=10))} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
