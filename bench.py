#!/usr/bin/env python
"""bench.py -- throughput of the cross-modal similarity hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched by torchrun)
  python bench.py --impl reference ...                            (CPU arm: the oracle port)

Primary line (BASELINE.json config[1]): symmetric InfoNCE forward+backward, batch 4096 x d=256,
bf16 tensor-core path, `value` = pairs/s with inputs resident in HBM (CUDA-graph replay of the
whole step, L2 flushed between steps), `e2e` = the same through `CLIPLoss` forward + backward on
pinned HOST batches (every step: H2D of both embedding matrices through prefetch.HostPairPrefetcher,
D2H of the loss, all inside the timed region; on one GPU the module replays CUDA graphs --
`CLIPLoss.graphed`); `e2e.serial_*` = eager, nothing overlapped.
N>1: weak scaling -- every rank owns one bucket of 4096 pairs of a global batch 4096*N
(reference `buckets` semantics, sharded on bucket boundaries; the two per-rank scalars are summed
inside the gradient-tail kernel over NVLink peer memory).
Extra objects on the same line: `c3` (global batch 32768, d=512, row-block sharded with
all-gather / all-reduce: the north-star multi-GPU loss), `retrieval` (top-10, gallery sharded) and,
on one GPU, `siglip` (SigLIP fwd+bwd at the primary shape, with its own CPU port beside it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_C2, D_C2 = 4096, 256
B_C3, D_C3 = 32768, 512
L2_FLUSH_BYTES = 256 << 20


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"), hbm=p["hbm_gbs"],
                    source="measured")
    except Exception:
        return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """Polls NVML (SM clock + throttle reasons) in a thread while the timed regions run."""

    def __init__(self, index: int, interval: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.interval = interval
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        import statistics
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CLIPLoss (reference src/coordination.py:26-47), all host threads
# -------------------------------------------------------------------------------------------------
def cpu_loss_step_time(B, d, steps, warmup):
    import torch
    from oracle.infonce import clip_loss_materialised
    from multimodal_plankton_recognition_b200 import synth
    torch.set_num_threads(os.cpu_count())
    torch.set_float32_matmul_precision("highest")
    img, pro, _ = synth.pairs(B, d, 1234, "cpu")
    ls = torch.ones((), requires_grad=True)
    x, y = img.requires_grad_(), pro.requires_grad_()
    times = []
    for it in range(warmup + steps):
        x.grad = y.grad = ls.grad = None
        t0 = time.perf_counter()
        clip_loss_materialised(x, y, ls, 1).backward()
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    return times


def run_reference(args, rank):
    if rank != 0:
        return
    times = cpu_loss_step_time(B_C2, D_C2, args.steps, args.warmup)
    total = sum(times)
    val = B_C2 * len(times) / total
    line = {
        "impl": "reference", "metric": "InfoNCE fwd+bwd pairs/s", "value": val, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"symmetric InfoNCE fwd+bwd, batch {B_C2} per GPU x d={D_C2}, f32 "
                               f"(BASELINE config[1]); N>1: global batch {B_C2 * args.gpus}, buckets={args.gpus}",
                   "global_batch": B_C2 * args.gpus, "d": D_C2, "buckets": args.gpus, "logit_scale": 1.0,
                   "note": "oracle port of reference CLIPLoss on host cores (/root/reference is Python and does "
                           "not travel to the GPU box); each step is one bucket of 4096 pairs -- the buckets "
                           "of the block-diagonal problem are independent, so pairs/s does not depend on N"},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{len(times)} full fwd+bwd steps at B={B_C2}, d={D_C2}, fp32, "
                                   f"best {1e3 * min(times):.1f} ms"},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, warmup, flush, sync_all):
    """Device time of `steps` calls of fn (CUDA events on the current stream, L2 flushed before
    each timed call, the flush itself outside the event pair).  Returns total milliseconds."""
    import torch
    for _ in range(warmup):
        flush.zero_()
        fn()
    sync_all()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    sync_all()
    return sum(a.elapsed_time(b) for a, b in evs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the c3 / retrieval extra objects")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from multimodal_plankton_recognition_b200 import CLIPLoss, _lib, ops, synth
    from multimodal_plankton_recognition_b200 import dist as pdist

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    mode = ops.MODES[args.precision]
    peaks = _peaks()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    flush = torch.empty(L2_FLUSH_BYTES, device=dev, dtype=torch.uint8)
    n, d = B_C2, D_C2
    Bg = n * world
    img, pro, _ = synth.pairs(n, d, 1234 + rank, dev)
    mod = CLIPLoss(precision=args.precision, sharded=world > 1).to(dev)
    ls = mod.logit_scale.detach()
    go = torch.ones(1, device=dev)

    # ---- value: the whole fwd+bwd step, inputs resident in HBM, captured in a CUDA graph ----
    xg = None
    if world == 1:
        def raw_step():
            loss, state = ops.clip_loss_forward_state(img, pro, ls, n, mode)
            return (loss,) + tuple(ops.clip_loss_backward_state(go, img, pro, ls, state, n, mode))
    else:
        # bucket-aligned sharding: the data path has no exchange; the two per-rank scalars (loss,
        # d logit_scale) are summed over the ranks INSIDE the gradient-tail kernel through NVLink peer
        # memory (plk_infonce_grad_finish_pair_xgpu), so the replayed graph is the whole step.
        if os.environ.get("PLK_BENCH_NCCL_SCALARS", "0") != "1":
            try:
                xg = pdist.XGpuScalars(dev)
            except Exception as e:      # no symmetric memory on this box: fall back to one NCCL all-reduce
                if rank == 0:
                    print(f"bench: symmetric memory unavailable ({e!r}); using NCCL for the scalars", file=sys.stderr)

        def raw_step():
            loss, state = pdist.sharded_fwd(img, pro, ls, world, mode, None, reduce_scalars=False)
            return (state[-2],) + tuple(pdist.sharded_bwd(state, go, "ddp", reduce_scalars=False, xgpu=xg))

    raw_step()
    torch.cuda.synchronize()
    l0 = lib.plk_launch_count()
    raw_step()
    launches_per_step = lib.plk_launch_count() - l0
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            raw_step()
    torch.cuda.current_stream().wait_stream(side)
    sync_all()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = raw_step()
    if world == 1 or xg is not None:
        step_fn = graph.replay
    else:
        def step_fn():
            graph.replay()
            dist.all_reduce(outs[0])    # [2] = (loss, d logit_scale) partials in one collective

    with ClockSampler(local, float(os.environ.get('PLK_BENCH_CLOCK_INTERVAL', '0.02'))) as clocks:
        total_ms = max_over_ranks(timed_steps(step_fn, args.steps, args.warmup, flush, sync_all))
        ms_per_step = total_ms / args.steps
        value = Bg / (ms_per_step * 1e-3)

        # ---- roofline: the dominant kernel (one direction of the recompute backward), timed alone ----
        u, idx, nx, _ = ops.l2norm(img, mode)
        v, idy, ny, _ = ops.l2norm(pro, mode)
        rs, cs, dg = ops.infonce_fwd_local(u, v, mode, d, 0, n, ls)
        def graphed(fn):   # one launch captured in a graph: host-side call overhead stays out of the timing
            fn()
            st_ = torch.cuda.Stream()
            st_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st_):
                fn()
            torch.cuda.current_stream().wait_stream(st_)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                keep = fn()
            g_.keep = keep
            return g_.replay

        k_ms = timed_steps(graphed(lambda: ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, n, ls, rs, cs, cs, rs, None)),
                           min(args.steps, 100), 3, flush, sync_all) / min(args.steps, 100)
        f_ms = timed_steps(graphed(lambda: ops.infonce_fwd_local(u, v, mode, d, 0, n, ls, rs, cs, dg)),
                           min(args.steps, 100), 3, flush, sync_all) / min(args.steps, 100)
        algo_flops = 4.0 * n * n * d          # the two reference GEMMs (dU = G V, dV = G^T U) this launch replaces
        achieved = algo_flops / (k_ms * 1e-3) / 1e12

        # ---- e2e: public API, pinned host inputs, H2D + D2H inside the timed region ----
        hx, hy = img.cpu().pin_memory(), pro.cpu().pin_memory()
        e2e_steps = min(args.steps, 100)
        clocks.interval = max(clocks.interval, 0.1)   # host-bound region: poll NVML less often

        # (1) serial: copy -> step -> read the loss, nothing overlapped (latency of one step)
        def serial_step():
            x = hx.to(dev, non_blocking=True).requires_grad_()
            y = hy.to(dev, non_blocking=True).requires_grad_()
            mod.logit_scale.grad = None
            loss = mod(image_emb=x, profile_emb=y, buckets=world)
            loss.backward()
            return float(loss.detach())          # D2H read of the step's result (synchronises)

        for _ in range(3):
            serial_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            serial_step()
        sync_all()
        serial_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

        # (2) streamed: the same per-step work (8 MiB H2D of that step's inputs, CLIPLoss forward +
        # backward, D2H read of its loss) with the copies on a side stream two batches ahead
        # (prefetch.HostPairPrefetcher) and each loss read one step late, so PCIe and the GPU overlap.
        from multimodal_plankton_recognition_b200.prefetch import HostPairPrefetcher
        e2e_warm = 5
        pf = HostPairPrefetcher(((hx, hy) for _ in range(e2e_steps + e2e_warm)), dev, depth=3)
        feed = iter(pf)
        losses, pending = [], [None]
        # one GPU: forward and backward of the module replay CUDA graphs (CLIPLoss.graphed =
        # torch.cuda.make_graphed_callables), which takes the host out of the critical path
        step_fn_e2e = None
        if world == 1 and os.environ.get("PLK_BENCH_GRAPHED_E2E", "1") == "1":
            try:
                step_fn_e2e = mod.graphed(img, pro)
            except Exception as e:
                print(f"bench: CLIPLoss.graphed unavailable ({e!r}); eager e2e", file=sys.stderr)

        def streamed_step():
            x, y = next(feed)
            x.requires_grad_()
            y.requires_grad_()
            mod.logit_scale.grad = None
            loss = step_fn_e2e(x, y) if step_fn_e2e is not None else mod(image_emb=x, profile_emb=y, buckets=world)
            loss.backward()
            read = pf.read_async(loss)
            if pending[0] is not None:
                losses.append(pending[0]())
            pending[0] = read

        for _ in range(e2e_warm):
            streamed_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            streamed_step()
        losses.append(pending[0]())          # the last step's loss: the queue is drained inside the timed region
        sync_all()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        e2e_value = Bg / (e2e_ms * 1e-3)
        assert all(l == l for l in losses[-e2e_steps:]), "NaN loss in the e2e run"

    line = {
        "metric": "InfoNCE fwd+bwd pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {
            "workload": f"symmetric InfoNCE fwd+bwd, batch {n} per GPU x d={d}, {args.precision} "
                        f"(BASELINE config[1]); N>1: global batch {Bg}, buckets={world} sharded on "
                        f"bucket boundaries",
            "global_batch": Bg, "d": d, "buckets": world, "logit_scale": 1.0,
            "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)",
            "timed_path": "CUDA-graph replay of plk_clip_loss_forward + plk_clip_loss_backward" if world == 1
                          else ("CUDA-graph replay of dist.sharded_fwd + dist.sharded_bwd; the (loss, d logit_scale) sum over ranks is "
                                "fused into the gradient-tail kernel (NVLink peer memory)" if xg is not None else
                                "CUDA-graph replay of dist.sharded_fwd + dist.sharded_bwd, then one NCCL all-reduce of (loss, d logit_scale)"),
            "parallelism": f"dp{world}"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 2 * n * d * 4, "d2h_bytes_per_step": 4,
                "path": ("CLIPLoss.graphed (forward + backward as CUDA graphs) " if step_fn_e2e is not None else
                         "CLIPLoss.forward + backward ") +
                        "on batches staged by prefetch.HostPairPrefetcher "
                        "(pinned host -> HBM on a copy stream, 2 batches ahead; each loss read back one step late)",
                "serial_ms_per_step": serial_ms, "serial_value": Bg / (serial_ms * 1e-3)},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "roofline": {"bound": "tensor", "kernel": "infonce_grad_tc2 (recompute backward, both directions in one launch)", "achieved": achieved,
                     "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                     "peak_source": f"{peaks['source']} burst", "kernel_ms": k_ms,
                     "algorithmic_flops_per_launch": algo_flops,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, one
                     # `ncu --set full` capture (profiles/r1_ncu_full_summary.txt): the 4 MiB of 16-bit
                     # operands; the partial-gradient slabs stay in L2
                     "traffic": 4287232 if (n, d, args.precision) == (4096, 256, "bf16") else None,
                     "executed_flops_per_launch": 2.0 * algo_flops,
                     "executed_tflops": 2.0 * achieved,
                     "note": "the recompute backward executes S = a.b^T once per direction on top of the two "
                             "credited GEMMs: executed tensor work is twice the algorithmic numerator",
                     "step_frac_of_peak": 6.0 * n * n * d / (ms_per_step * 1e-3) / 1e12 / peaks["bf16"],
                     "fwd_kernel_ms": f_ms},
        "clocks": clocks.summary(),
    }

    if rank == 0 and world == 1:
        cpu_steps = 6
        times = cpu_loss_step_time(n, d, cpu_steps, 2)
        line["cpu_baseline"] = {"value": n * len(times) / sum(times), "unit": "pairs/s",
                                "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{cpu_steps} full fwd+bwd steps of the same workload (B={n}, d={d}, fp32 "
                                          f"torch CPU port of reference CLIPLoss), best {1e3 * min(times):.1f} ms"}

    if not args.no_extras:
        try:
            line["c3"] = bench_c3(args, world, rank, dev, mode, flush, sync_all, max_over_ranks, peaks)
        except Exception as e:  # extras must never take the primary line down
            line["c3"] = {"error": repr(e)}
        try:
            line["retrieval"] = bench_retrieval(args, world, rank, dev, sync_all, max_over_ranks, peaks)
        except Exception as e:
            line["retrieval"] = {"error": repr(e)}

        if world == 1:
            try:
                line["siglip"] = bench_siglip(args, dev, mode, flush, sync_all, peaks)
            except Exception as e:
                line["siglip"] = {"error": repr(e)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_c3(args, world, rank, dev, mode, flush, sync_all, max_over_ranks, peaks):
    """BASELINE config[2]: global batch 32768, d=512, one bucket, rows sharded across the ranks
    (all-gather of normalised embeddings, all-reduce of column sums; both backward passes local)."""
    import torch
    from multimodal_plankton_recognition_b200 import CLIPLoss, synth
    n = B_C3 // world
    img, pro, _ = synth.pairs(n, D_C3, 4321 + rank, dev)
    mod = CLIPLoss(precision=args.precision, sharded=world > 1).to(dev)
    x, y = img.requires_grad_(), pro.requires_grad_()

    def step():
        x.grad = y.grad = mod.logit_scale.grad = None
        mod(image_emb=x, profile_emb=y, buckets=1).backward()

    steps = 10
    ms = max_over_ranks(timed_steps(step, steps, 3, flush, sync_all)) / steps
    flops = 6.0 * B_C3 * B_C3 * D_C3
    return {"workload": f"InfoNCE fwd+bwd global batch {B_C3}, d={D_C3}, row-block sharded over {world} GPU(s)",
            "value": B_C3 / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "scaling": "strong",
            "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "frac_of_peak_all_gpus": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_sustained"] * world)}


def bench_siglip(args, dev, mode, flush, sync_all, peaks):
    """SigLIP fwd+bwd at the primary shape (B=4096, d=256), CUDA-graph replay of
    plk_siglip_loss_forward + plk_siglip_loss_backward, L2 flushed between steps; CPU port timed beside it."""
    import torch
    from multimodal_plankton_recognition_b200 import ops, synth
    n, d = B_C2, D_C2
    img, pro, _ = synth.pairs(n, d, 1234, dev)
    ls = torch.ones((), device=dev)
    bias = torch.full((), -10.0, device=dev)
    go = torch.ones(1, device=dev)

    def raw_step():
        loss, state = ops.siglip_loss_forward_state(img, pro, ls, bias, n, mode)
        return (loss,) + tuple(ops.siglip_loss_backward_state(go, img, pro, ls, bias, state, n, mode))

    raw_step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            raw_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = raw_step()
    steps = min(args.steps, 100)
    ms = timed_steps(graph.replay, steps, 3, flush, sync_all) / steps
    out = {"workload": f"SigLIP fwd+bwd, batch {n} x d={d}, {args.precision}", "value": n / (ms * 1e-3),
           "unit": "pairs/s", "ms_per_step": ms,
           "algorithmic_tflops": 6.0 * n * n * d / (ms * 1e-3) / 1e12,
           "frac_of_peak": 6.0 * n * n * d / (ms * 1e-3) / 1e12 / peaks["bf16"]}
    del keep
    try:
        from oracle import siglip as osig
        torch.set_num_threads(os.cpu_count() or 1)
        xc, yc = img.cpu().requires_grad_(), pro.cpu().requires_grad_()
        lc, bc = torch.ones((), requires_grad=True), torch.full((), -10.0, requires_grad=True)
        times = []
        for _ in range(3):
            xc.grad = yc.grad = lc.grad = bc.grad = None
            t0 = time.perf_counter()
            osig.siglip_loss_materialised(xc, yc, lc, bc, 1).backward()
            times.append(time.perf_counter() - t0)
        out["cpu_baseline"] = {"value": n / min(times), "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"3 fwd+bwd steps of the same workload, best {1e3 * min(times):.1f} ms"}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def bench_retrieval(args, world, rank, dev, sync_all, max_over_ranks, peaks):
    """BASELINE config[3]: 1M-row gallery (sharded over the ranks), 100k queries, d=512, top-10."""
    import torch
    import torch.distributed as dist
    from multimodal_plankton_recognition_b200 import synth
    from multimodal_plankton_recognition_b200.ann import GpuExactIndex
    ng, nq, d, k = 1_000_000, 100_000, 512, 10
    shard = ng // world
    gal, _ = synth.unit_embeddings(shard, d, 99 + rank, dev, modality=1)
    q, _ = synth.unit_embeddings(nq, d, 7, dev, modality=0)
    index = GpuExactIndex.from_device(gal, precision="bf16", gallery_offset=rank * shard)
    del gal

    def search():
        idx, dst = index.search_device(q, k)
        if world > 1:
            from multimodal_plankton_recognition_b200.dist import merge_shard_results
            idx, dst = merge_shard_results(idx, dst, k, None)
        return idx, dst

    search()
    sync_all()
    steps = 3
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record()
        search()
        b.record()
    sync_all()
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / steps
    flops = 2.0 * nq * ng * d
    return {"workload": f"cosine/euclidean top-{k}: {ng} gallery rows sharded over {world} GPU(s), {nq} queries, d={d}, "
                        f"bf16 candidates + exact fp32 re-score",
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_search": ms,
            "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "frac_of_peak_all_gpus": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_sustained"] * world)}


if __name__ == "__main__":
    main()
