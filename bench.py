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


def _workload(n, d, world):
    """Same string in both arms (the arithmetic type is the line's `dtype`)."""
    return (f"symmetric InfoNCE fwd+bwd, batch {n} per GPU x d={d} (BASELINE config[1]); "
            f"N>1: global batch {n * world}, buckets={world} sharded on bucket boundaries")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"), hbm=p["hbm_gbs"],
                    source="measured")
    except Exception:
        return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def _ncu_traffic(kernel, n, d, precision):
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        e = t.get(kernel)
        if e and (e.get("B"), e.get("d"), e.get("precision")) == (n, d, precision):
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


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
        "config": {"workload": _workload(B_C2, D_C2, args.gpus),
                   "global_batch": B_C2 * args.gpus, "d": D_C2, "buckets": args.gpus, "logit_scale": 1.0,
                   "parallelism": f"dp{args.gpus}",
                   "note": "oracle port of reference CLIPLoss on host cores, fp32 (/root/reference is Python and does "
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
        # An event pair around ONE graph launch also times the launch itself: a graph holding a single one-thread
        # kernel measures ~6 us on these boxes (`event_floor_ms`).  Reported beside kernel_ms, never subtracted from it:
        # the MARGINAL launch duration = (graph of 4 backward launches - graph of 1) / 3, L2 flushed before each replay.
        tiny = torch.zeros(1, device=dev)
        floor_ms = timed_steps(graphed(lambda: tiny.add_(1.0)), min(args.steps, 100), 3, flush, sync_all) / min(args.steps, 100)

        def bwd4():
            for _ in range(4):
                keep = ops.infonce_grad_pair_local(u, v, v, u, mode, d, 0, n, ls, rs, cs, cs, rs, None)
            return keep

        k4_ms = timed_steps(graphed(bwd4), min(args.steps, 100), 3, flush, sync_all) / min(args.steps, 100)
        k_marginal_ms = (k4_ms - k_ms) / 3.0

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
        if os.environ.get("PLK_BENCH_GRAPHED_E2E", "1") == "1":
            # N > 1: bucket-aligned sharding with the in-kernel scalar exchange has no collective launch,
            # so the module's forward and backward can be graphed there as well
            try:
                step_fn_e2e = mod.graphed(img, pro, buckets=world)
            except Exception as e:
                if rank == 0:
                    print(f"bench: CLIPLoss.graphed unavailable ({e!r}); eager e2e", file=sys.stderr)
            if world > 1:      # every rank must take the same path
                ok = torch.tensor([1 if step_fn_e2e is not None else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok) == 0:
                    step_fn_e2e = None

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
            "workload": _workload(n, d, world),
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
        "roofline": {"bound": "tensor", "kernel": "infonce_grad_tc5 (recompute backward, both directions in one launch, tcgen05 cta_group::2)", "achieved": achieved,
                     "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                     "peak_source": f"{peaks['source']} burst", "kernel_ms": k_ms,
                     "algorithmic_flops_per_launch": algo_flops,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel at this shape,
                     # parsed from the committed `ncu --set full` capture (profiles/r2_traffic.json, written
                     # by tools/ncu_summary.py); null when there is no capture of the current kernel
                     "traffic": _ncu_traffic("infonce_grad_tc5", n, d, args.precision),
                     "executed_flops_per_launch": 2.0 * algo_flops,
                     "executed_tflops": 2.0 * achieved,
                     "note": "the recompute backward executes S = a.b^T once per direction on top of the two "
                             "credited GEMMs: executed tensor work is twice the algorithmic numerator",
                     "step_frac_of_peak": 6.0 * n * n * d / (ms_per_step * 1e-3) / 1e12 / peaks["bf16"],
                     "fwd_kernel_ms": f_ms,
                     "event_floor_ms": floor_ms,
                     "kernel_ms_marginal": k_marginal_ms,
                     "frac_marginal": algo_flops / (k_marginal_ms * 1e-3) / 1e12 / peaks["bf16"],
                     "timing_note": "kernel_ms / frac: CUDA events around a graph holding ONE launch of the kernel (includes "
                                    "the launch floor, event_floor_ms = the same measurement of a one-thread kernel); "
                                    "kernel_ms_marginal = (graph of 4 launches - graph of 1) / 3"},
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
            for key, fn in (("siglip", lambda: bench_siglip(args, dev, mode, flush, sync_all, peaks)),
                            ("n1", lambda: bench_n1(args, dev, mode, flush, sync_all, peaks)),
                            ("c1", lambda: bench_c1(args, dev, mode, flush, sync_all)),
                            ("c5", lambda: bench_c5(args, dev))):
                try:
                    line[key] = fn()
                except Exception as e:
                    line[key] = {"error": repr(e)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_c3(args, world, rank, dev, mode, flush, sync_all, max_over_ranks, peaks):
    """BASELINE config[2]: global batch 32768, d=512, one bucket, rows sharded across the ranks
    (all-gather of normalised embeddings, all-reduce of column sums; both backward passes local).
    N > 1 adds `parity` (rank 0 recomputes the WHOLE batch unsharded on its GPU and compares the global loss,
    d logit_scale and the gradients of its own rows) and `efficiency_vs_n1` against that unsharded run."""
    import torch
    import torch.distributed as dist
    from multimodal_plankton_recognition_b200 import CLIPLoss, synth
    n = B_C3 // world
    img, pro, _ = synth.pairs(n, D_C3, 4321 + rank, dev)
    mod = CLIPLoss(precision=args.precision, sharded=world > 1).to(dev)
    x, y = img.requires_grad_(), pro.requires_grad_()
    last = {}

    def step():
        x.grad = y.grad = mod.logit_scale.grad = None
        loss = mod(image_emb=x, profile_emb=y, buckets=1)
        loss.backward()
        last["loss"] = loss.detach()

    steps = 10
    ms_module = max_over_ranks(timed_steps(step, steps, 3, flush, sync_all)) / steps
    ms, path = ms_module, "CLIPLoss(sharded) module, eager (autograd + NCCL launches from the host)"
    if world > 1 and os.environ.get("PLK_BENCH_C3_GRAPH", "1") == "1":
        # the same step (dist.sharded_fwd + dist.sharded_bwd: kernels AND the three NCCL collectives) captured
        # in one CUDA graph: the host queues a single launch per step
        try:
            from multimodal_plankton_recognition_b200 import dist as pdist
            ls, go = mod.logit_scale.detach(), torch.ones(1, device=dev)
            xd, yd = img.detach(), pro.detach()

            def raw():
                loss, state = pdist.sharded_fwd(xd, yd, ls, 1, mode, None)
                return (loss,) + tuple(pdist.sharded_bwd(state, go, "ddp"))

            raw()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    raw()
            torch.cuda.current_stream().wait_stream(side)
            sync_all()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                keep = raw()
            ms_graph = max_over_ranks(timed_steps(graph.replay, steps, 3, flush, sync_all)) / steps
            # the replayed step reproduces the module's result
            assert abs(float(keep[0]) - float(last["loss"])) <= 1e-5 * abs(float(last["loss"])), \
                f"graph replay loss {float(keep[0])} != module loss {float(last['loss'])}"
            if ms_graph < ms:
                ms, path = ms_graph, "CUDA-graph replay of dist.sharded_fwd + dist.sharded_bwd (kernels + 3 NCCL collectives)"
            last["ms_graph"] = ms_graph
        except Exception as e:
            import traceback
            last["graph_error"] = repr(e) + " | " + traceback.format_exc()[-600:]
    flops = 6.0 * B_C3 * B_C3 * D_C3
    out = {"workload": f"InfoNCE fwd+bwd global batch {B_C3}, d={D_C3}, row-block sharded over {world} GPU(s)",
           "value": B_C3 / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "scaling": "strong", "timed_path": path,
           "ms_per_step_module_eager": ms_module,
           "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
           "frac_of_peak_all_gpus": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_sustained"] * world)}
    for k_ in ("ms_graph", "graph_error"):
        if k_ in last:
            out[k_] = last[k_]
    if world > 1:
        # grad_scale="ddp" (the module's default) pre-multiplies the embedding gradients by the world size
        got = (float(last["loss"]), x.grad.detach().clone() / world, y.grad.detach().clone() / world,
               float(mod.logit_scale.grad))
        if rank == 0:
            parts = [synth.pairs(n, D_C3, 4321 + r, dev) for r in range(world)]
            fx = torch.cat([p[0] for p in parts]).requires_grad_()
            fy = torch.cat([p[1] for p in parts]).requires_grad_()
            del parts
            one = CLIPLoss(precision=args.precision).to(dev)

            def step1():
                fx.grad = fy.grad = one.logit_scale.grad = None
                loss1 = one(image_emb=fx, profile_emb=fy, buckets=1)
                loss1.backward()
                last["loss1"] = loss1.detach()

            ms1 = timed_steps(step1, 3, 2, flush, torch.cuda.synchronize) / 3
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            out["parity"] = {"against": "unsharded single-GPU run of the same kernels on the concatenated batch (rank 0)",
                             "loss_rel": abs(got[0] - float(last["loss1"])) / abs(float(last["loss1"])),
                             "d_image_rel": rel(got[1], fx.grad[:n]), "d_profile_rel": rel(got[2], fy.grad[:n]),
                             "d_logit_scale_rel": abs(got[3] - float(one.logit_scale.grad)) / max(abs(float(one.logit_scale.grad)), 1e-12),
                             "rows_compared": n}
            out["ms_per_step_n1"] = ms1
            out["efficiency_vs_n1"] = ms1 / (world * ms)
            del fx, fy
        dist.barrier()
    return out


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
    """BASELINE config[3]: 1M-row gallery (sharded over the ranks), 100k queries, d=512, top-10.
    `value`: queries resident in HBM, search (+ all-gather / merge of the shard lists) timed with CUDA events;
    `roofline`: the candidate kernel (topk_tc_kernel + its list merge) alone; `e2e`: the reference-facing call
    `kneighbors(numpy queries) -> numpy (idx, dist)` with the H2D of the queries and the D2H of the result
    inside the timed region; `cpu_baseline` (N = 1): exact brute force on the host cores for a 1000-query
    subsample; N > 1: `parity` + `efficiency_vs_n1` against the unsharded index built on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from multimodal_plankton_recognition_b200 import ANNClassifier, _lib, synth
    from multimodal_plankton_recognition_b200.ann import GpuExactIndex
    ng, nq, d, k = 1_000_000, 100_000, 512, 10
    shard = ng // world
    gal, lab = synth.unit_embeddings(shard, d, 99 + rank, dev, modality=1)
    q, _ = synth.unit_embeddings(nq, d, 7, dev, modality=0)
    index = GpuExactIndex.from_device(gal, precision="bf16", gallery_offset=rank * shard)

    def search():
        idx, dst = index.search_device(q, k)
        if world > 1:
            from multimodal_plankton_recognition_b200.dist import merge_shard_results
            idx, dst = merge_shard_results(idx, dst, k, None)
        return idx, dst

    def timed(fn, steps, sync):
        fn()
        sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        sync()
        return sum(a.elapsed_time(b) for a, b in evs) / steps

    steps = 10
    ms = max_over_ranks(timed(search, steps, sync_all))
    flops = 2.0 * nq * ng * d
    out = {"workload": f"cosine/euclidean top-{k}: {ng} gallery rows sharded over {world} GPU(s), {nq} queries, d={d}, "
                       f"bf16 candidates + exact fp32 re-score",
           "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_search": ms, "searches_timed": steps,
           "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
           "frac_of_peak_all_gpus": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_sustained"] * world)}

    # roofline: the tensor-core candidate search of this rank's shard alone (plk_topk_candidates)
    lib = _lib.load()
    from multimodal_plankton_recognition_b200 import ops
    kc = 16
    q_op = ops.l2norm(q, index.mode, normalise=False)[0]
    ci = torch.empty((nq, kc), device=dev, dtype=torch.int32)
    ck = torch.empty((nq, kc), device=dev, dtype=torch.float32)
    wsb = lib.plk_topk_workspace_bytes(nq, index.n, d, kc, index.mode)
    ws = torch.empty(max(wsb, 16), device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream(dev).cuda_stream

    def cand():
        lib.check(lib.plk_topk_candidates(q_op.data_ptr(), index.g_op.data_ptr(), index.mode, q_op.stride(0),
                                          index.g_sqn.data_ptr(), nq, index.n, d, kc, index.gallery_offset,
                                          ci.data_ptr(), ck.data_ptr(), ws.data_ptr(), wsb, st), "plk_topk_candidates")

    k_ms = timed(cand, 5, torch.cuda.synchronize)
    ach = 2.0 * nq * index.n * d / (k_ms * 1e-3) / 1e12
    out["roofline"] = {"bound": "tensor", "kernel": "topk_tc_kernel (+ select_kernel merging the per-chunk lists)",
                       "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                       "frac": ach / peaks["bf16_sustained"], "frac_of_burst": ach / peaks["bf16"],
                       "peak_source": f"{peaks['source']} sustained (the kernel runs for {k_ms:.0f} ms)",
                       "kernel_ms": k_ms, "algorithmic_flops_per_launch": 2.0 * nq * index.n * d,
                       "traffic": _ncu_traffic("topk_tc_kernel", nq, d, "bf16")}
    del q_op, ci, ck, ws

    # e2e: numpy in, numpy out through the reference-facing class
    q_np = q.cpu().numpy()
    if world == 1:
        clf = ANNClassifier.from_index(index, lab.cpu().numpy())
    else:
        from multimodal_plankton_recognition_b200.dist import ShardedANNClassifier
        clf = ShardedANNClassifier.from_device(gal, lab, precision="bf16")
    e2e_steps = 3
    clf.kneighbors(q_np, k=k, epsilon=.3)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = clf.kneighbors(q_np, k=k, epsilon=.3)
    sync_all()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    out["e2e"] = {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_search": e2e_ms,
                  "h2d_bytes_per_step": int(q_np.nbytes), "d2h_bytes_per_step": int(res[0][0].nbytes + res[0][1].nbytes),
                  "path": ("ANNClassifier" if world == 1 else "dist.ShardedANNClassifier") +
                          ".kneighbors(numpy [100000, 512] fp32) -> numpy (idx int32, dist fp32); pageable host "
                          "arrays, synchronous, as scripts/benchmark_cross.py calls it"}

    if world == 1 and rank == 0:
        try:
            from oracle import ann as oann
            torch.set_num_threads(os.cpu_count() or 1)
            sub = 1000
            g_cpu, q_cpu = gal.cpu(), q[:sub].cpu()
            oann.brute_force_topk_blas(q_cpu[:100], g_cpu, k)
            t0 = time.perf_counter()
            bi, bd = oann.brute_force_topk_blas(q_cpu, g_cpu, k)
            dt = time.perf_counter() - t0
            gi, _ = index.search_device(q[:sub], k)
            agree = float((gi.cpu().long() == bi).float().mean())
            out["cpu_baseline"] = {"value": sub / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{sub} of the {nq} queries against the full {ng}-row gallery: exact fp32 "
                                             f"brute force (BLAS matmul + topk, oracle.ann.brute_force_topk_blas), "
                                             f"{dt:.2f} s; the reference's pynndescent search is absent from this image",
                                   "index_agreement_with_gpu": agree}
            del g_cpu, q_cpu
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}

    if world > 1:
        got_i, got_d = search()
        if rank == 0:
            full = torch.cat([synth.unit_embeddings(shard, d, 99 + r, dev, modality=1)[0] for r in range(world)])
            one = GpuExactIndex.from_device(full, precision="bf16")
            del full
            ms1 = timed(lambda: one.search_device(q, k), 3, torch.cuda.synchronize)
            sub = 1024
            wi, wd = one.search_device(q[:sub], k)
            same = wi == got_i[:sub]
            out["parity"] = {"against": "unsharded index over the concatenated gallery (rank 0), first 1024 queries",
                             "dist_max_abs_diff": float((wd - got_d[:sub]).abs().max()),
                             "index_mismatch_frac": float(1.0 - same.float().mean()),
                             "mismatches_only_at_equal_distance": bool((wd[~same] == got_d[:sub][~same]).all())}
            out["ms_per_search_n1"] = ms1
            out["efficiency_vs_n1"] = ms1 / (world * ms)
            del one
        dist.barrier()
    return out


def bench_n1(args, dev, mode, flush, sync_all, peaks):
    """the two bias-free projection Linears + the loss, forward and backward.  Fused path
    (`CLIPLoss.forward_projected`: projection + normalisation in one tcgen05 kernel per modality) against the
    reference's arrangement (nn.Linear under bf16 autocast -- its '16-mixed' trainer precision -- feeding the
    same loss module), and the projection kernel alone against the measured tensor peak."""
    import torch
    from multimodal_plankton_recognition_b200 import CLIPLoss, ops
    B, f_i, f_p, d = 4096, 1280, 192, 256      # EfficientNet-B0 features, profile-encoder features, BASELINE d
    g = torch.Generator(device="cpu").manual_seed(5)
    z = torch.randn(B, 64, generator=g)
    fi = (z @ torch.randn(64, f_i, generator=g) / 8 + 0.3 * torch.randn(B, f_i, generator=g)).to(dev)
    fp = (z @ torch.randn(64, f_p, generator=g) / 8 + 0.3 * torch.randn(B, f_p, generator=g)).to(dev)
    mod = CLIPLoss(precision=args.precision).to(dev)
    pi = torch.nn.Linear(f_i, d, bias=False).to(dev)
    pp = torch.nn.Linear(f_p, d, bias=False).to(dev)
    xi, xp = fi.clone().requires_grad_(), fp.clone().requires_grad_()
    params = (xi, xp, pi.weight, pp.weight, mod.logit_scale)

    def fused():
        for t in params:
            t.grad = None
        loss = mod.forward_projected(xi, xp, pi, pp)
        loss.backward()
        return loss

    def unfused():
        for t in params:
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a, b = pi(xi), pp(xp)
        loss = mod(image_emb=a.float(), profile_emb=b.float())
        loss.backward()
        return loss

    steps = min(args.steps, 50)
    out = {"workload": f"batch {B}: Linear({f_i}->{d}) + Linear({f_p}->{d}) (no bias) + symmetric InfoNCE, fwd+bwd incl. "
                       f"weight and feature gradients, {args.precision}",
           "timed_path": "eager autograd step (CUDA events, L2 flushed between steps)"}
    l_f, l_u = float(fused().detach()), float(unfused().detach())
    out["fused_ms_per_step"] = timed_steps(fused, steps, 3, flush, sync_all) / steps
    out["unfused_ms_per_step"] = timed_steps(unfused, steps, 3, flush, sync_all) / steps
    out["value"] = B / (out["fused_ms_per_step"] * 1e-3)
    out["unit"] = "pairs/s"

    def graph_of(fn):   # the eager step is host-bound (~0.4 ms of Python / autograd for ~0.15 ms of kernels)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            keep = fn()
        g_.keep = keep
        return g_.replay

    try:
        out["fused_graphed_ms_per_step"] = timed_steps(graph_of(fused), steps, 3, flush, sync_all) / steps
        out["unfused_graphed_ms_per_step"] = timed_steps(graph_of(unfused), steps, 3, flush, sync_all) / steps
        out["value"] = B / (out["fused_graphed_ms_per_step"] * 1e-3)
        out["timed_path"] = ("whole autograd step (forward_projected + backward, weight and feature gradients) replayed "
                             "as one CUDA graph, L2 flushed between steps; *_ms_per_step without 'graphed': eager")
    except Exception as e:   # capture is an optimisation of the measurement, not of the product
        out["graph_capture_error"] = repr(e)
    out["parity"] = {"loss_fused": l_f, "loss_linear_then_module": l_u, "rel_diff": abs(l_f - l_u) / abs(l_u),
                     "note": "the unfused arm rounds the projected embedding to bf16 (autocast) before the loss "
                             "normalises it; the fused kernel normalises the fp32 accumulator"}
    if mode != ops.MODES["fp32"]:
        odt = ops.OP_TORCH_DTYPE[mode]
        x16, w16 = fi.to(odt), pi.weight.detach().to(odt)
        k_ms = timed_steps(lambda: ops.project_normalise(x16, w16, mode), steps, 3, flush, sync_all) / steps
        flops = 2.0 * B * f_i * d
        out["roofline"] = {"bound": "tensor", "kernel": "proj_norm_tc2 (image modality: [4096 x 1280] x [256 x 1280]^T "
                           "+ squared norms + normalised operand)", "kernel_ms": k_ms,
                           "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peaks["bf16"], "unit": "TFLOP/s",
                           "frac": flops / (k_ms * 1e-3) / 1e12 / peaks["bf16"],
                           "note": "64 CTAs (32 row blocks x 2 output halves, 43 % of the SMs) and 21 MB of mandatory traffic: an "
                                   "occupancy-bound shape, timed with the host-side call inside the region"}
    return out


def bench_c1(args, dev, mode, flush, sync_all):
    """BASELINE config[0] (model_cards/example_multi.yaml scale): batch 256, d=512 -- coordination loss fwd+bwd
    plus top-10 retrieval over the batch (profile embeddings as gallery, image embeddings as queries).  Launch /
    latency bound (0.2 GFLOP): microseconds, no roofline fraction.  CPU port beside it."""
    import numpy as np
    import torch
    from multimodal_plankton_recognition_b200 import ops, synth
    from multimodal_plankton_recognition_b200.ann import GpuExactIndex
    B, d, k = 256, 512, 10
    img, pro, _ = synth.pairs(B, d, 11, dev)
    ls = torch.ones((), device=dev)
    go = torch.ones(1, device=dev)

    def raw_step():
        loss, state = ops.clip_loss_forward_state(img, pro, ls, B, mode)
        return (loss,) + tuple(ops.clip_loss_backward_state(go, img, pro, ls, state, B, mode))

    def graph_of(fn):
        fn()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()
        g.keep = keep
        return g.replay

    steps = min(args.steps, 100)
    loss_us = 1e3 * timed_steps(graph_of(raw_step), steps, 3, flush, sync_all) / steps
    un = torch.nn.functional.normalize
    gal, qry = un(pro), un(img)
    index = GpuExactIndex.from_device(gal, precision=args.precision if args.precision != "fp32" else "fp32")
    topk_us = 1e3 * timed_steps(graph_of(lambda: index.search_device(qry, k)), steps, 3, flush, sync_all) / steps
    out = {"workload": f"batch {B}, d={d}, {args.precision}: InfoNCE fwd+bwd + top-{k} retrieval over the batch",
           "loss_fwd_bwd_us": loss_us, "top10_us": topk_us, "us_per_step": loss_us + topk_us,
           "value": B / ((loss_us + topk_us) * 1e-6), "unit": "pairs/s",
           "timed_path": "CUDA-graph replays (loss: plk_clip_loss_forward + backward; retrieval: candidate search + "
                         "exact re-score), L2 flushed between steps"}
    try:
        from oracle import ann as oann
        from oracle.infonce import clip_loss_materialised
        torch.set_num_threads(os.cpu_count() or 1)
        xc, yc = img.cpu().requires_grad_(), pro.cpu().requires_grad_()
        lc = torch.ones((), requires_grad=True)
        tl, tk = [], []
        gal_np, q_np = gal.cpu().numpy(), qry.cpu().numpy()
        for _ in range(5):
            xc.grad = yc.grad = lc.grad = None
            t0 = time.perf_counter()
            clip_loss_materialised(xc, yc, lc, 1).backward()
            t1 = time.perf_counter()
            want = oann.ExactIndex(gal_np).query(q_np, k=k)
            t2 = time.perf_counter()
            tl.append(t1 - t0)
            tk.append(t2 - t1)
        gi, gd = index.search_device(qry, k)
        out["cpu_baseline"] = {"value": B / (min(tl) + min(tk)), "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                               "loss_fwd_bwd_us": 1e6 * min(tl), "top10_us": 1e6 * min(tk),
                               "sample": "5 repetitions of the same batch, best of 5 (oracle port of CLIPLoss + exact index)"}
        out["parity"] = {"top10_dist_equal": bool(np.array_equal(gd.cpu().numpy(), want[1])),
                         "top10_index_mismatch_frac": float((gi.cpu().numpy() != want[0]).mean())}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    return out


def bench_c5(args, dev):
    """BASELINE config[4]: the 5-fold few-shot k-NN benchmark of reference scripts/benchmark_cross_folds.py on
    synthetic CytoSense-shaped embeddings (9250 samples, 27 long-tailed classes, d=512, StratifiedKFold(5,
    shuffle, seed 0); galleries of n in (2,4,8,12,16) per class, K = (1,3,5,7,9) and the BASELINE k = 10, eight
    set-ups incl. image->profile and profile->image) through `harness.cross_benchmark_folds`.  The CPU port
    (oracle.ann.fold_benchmark_port, the reference's loop over the oracle's exact index) runs one fold at n = 16
    beside it; its labels must be IDENTICAL to the GPU harness's on that fold."""
    import random
    import numpy as np
    import torch
    from sklearn.model_selection import StratifiedKFold
    from sklearn.preprocessing import LabelEncoder
    from multimodal_plankton_recognition_b200 import harness, synth
    N, d = 9250, 512
    NS, K, repeats = (2, 4, 8, 12, 16), (1, 3, 5, 7, 9, 10), 1
    img, pro, lab = synth.pairs(N, d, 2024, "cpu", separation=0.15)     # k-NN accuracy 80-90 %: votes are contested
    un = torch.nn.functional.normalize
    img, pro, lab = un(img).numpy(), un(pro).numpy(), lab.numpy()
    names = np.array([f"class_{c:02d}" for c in lab])
    coder = LabelEncoder().fit(names)
    folds = list(StratifiedKFold(5, shuffle=True, random_state=0).split(img, lab))

    def fold_data(i):
        tr, te = folds[i]
        return (img[tr], pro[tr], names[tr]), (img[te], pro[te], names[te])

    def run_gpu():
        random.seed(0)
        res = {}
        for i in range(len(folds)):
            train, test = fold_data(i)
            res[i] = {n: harness.cross_benchmark_folds(train, test, coder, n, repeats, K, plk_precision="bf16")
                      for n in NS}
        return res

    run_gpu()                          # warm-up (module load, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = run_gpu()
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    n_pred = len(folds) * len(NS) * repeats * len(K) * 8
    queries = sum(len(folds[i][1]) for i in range(len(folds))) * len(NS) * repeats * len(K) * 8
    acc = {}
    for setup in ("I - P", "P - I", "I+P - I"):
        hit = tot = 0
        for i in res:
            r = res[i][16][0]
            hit += int((r["pred"][10][setup] == r["true"]).sum())
            tot += len(r["true"])
        acc[setup] = hit / tot
    out = {"workload": f"5-fold few-shot k-NN benchmark (scripts/benchmark_cross_folds.py): {N} samples, 27 classes, d={d}, "
                       f"n per class in {NS}, K={K}, 8 set-ups, {repeats} repeat(s) per fold",
           "seconds": gpu_s, "predict_calls_replaced": n_pred, "value": queries / gpu_s, "unit": "classified queries/s",
           "accuracy_k10_n16": acc,
           "timed_path": "harness.cross_benchmark_folds (numpy in, class names out; one search per set-up at max(K), "
                         "prefix votes per k), all folds, wall clock"}
    try:
        from oracle import ann as oann
        train, test = fold_data(0)
        sub = 600                                  # bounded CPU sample: the first 600 queries of fold 0
        test = tuple(t[:sub] for t in test)
        random.seed(0)
        gpu0 = harness.cross_benchmark_folds(train, test, coder, 16, 1, K, plk_precision="bf16")
        random.seed(0)
        t0 = time.perf_counter()
        cpu0 = oann.fold_benchmark_port(train, test, coder, 16, 1, K)
        cpu_s = time.perf_counter() - t0
        same = all(np.array_equal(gpu0[0]["pred"][kk][sname], cpu0[0]["pred"][kk][sname])
                   for kk in K for sname in cpu0[0]["pred"][kk])
        assert same, "k-NN fold labels differ from the CPU port"
        q0 = len(test[2]) * len(K) * 8
        out["cpu_baseline"] = {"value": q0 / cpu_s, "unit": "classified queries/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"fold 0, n = 16, first {sub} test samples, K = {K}, 8 set-ups ({q0} classified "
                                         f"queries) through oracle.ann.fold_benchmark_port (one predict per k, as the "
                                         f"reference's loop), {cpu_s:.1f} s"}
        out["parity"] = {"labels_identical_to_cpu_port": bool(same), "labels_compared": q0}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    main()
