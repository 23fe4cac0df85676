#!/usr/bin/env python
"""bench.py -- train samples/sec of the ViT-B/16 224px bf16 training step (BASELINE.json configs[1] / [2]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # the B200-native arm (this repo)
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port), host cores
  torchrun ... bench.py --gpus N ...                              # one rank per GPU, weak scaling (256 images / GPU)

One step = forward + cross-entropy + backward + (N > 1: bucketed gradient all-reduce) + Adam step on one batch of
synthetic images (randn, fp32 NCHW) and labels (uniform int64), random-init weights of the named architecture.
Rank 0 prints ONE JSON line.  ``value`` is measured with inputs resident in HBM; ``e2e`` goes through the public
module API with pinned HOST buffers, the host->device copy of every step's batch and a device->host read of the loss
inside the timed region.
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

FLOP_PER_IMAGE_FWD_BWD = 105_382_969_344  # BASELINE.md section 2 (GEMM-only, 3x forward)
CONFIG_NAME = "vit_b16"
PER_GPU_BATCH = 256


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm=p["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; nvidia-smi fields of the profiling recipe)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # pragma: no cover
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (pinned bit-for-bit to the reference) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_samples_per_sec(steps: int, warmup: int, budget_s: float):
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vit_oracle as vo

    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        usable = os.cpu_count() or 1
    cfg = vo.vit_config(CONFIG_NAME)
    # "all the host threads it can use": on the shared GPU hosts 128 logical CPUs are visible but oversubscribing them
    # is slower than using fewer, so a few short probes pick the fastest thread count (reported as `cores`)
    cores = usable
    if usable > 16:
        probe_sd = vo.init_state_dict(cfg, seed=0, perturb=False)
        px, py = vo.synthetic_batch(cfg, 4, seed=0)
        best = None
        for n in sorted({usable, max(1, usable // 2), max(1, usable // 4), 16}, reverse=True):
            torch.set_num_threads(n)
            pp = {k: v.clone().requires_grad_(True) for k, v in probe_sd.items()}
            vo.cross_entropy(vo.classifier_forward(pp, px, cfg), py).backward()  # warm
            t0 = time.perf_counter()
            vo.cross_entropy(vo.classifier_forward(pp, px, cfg), py).backward()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
        cores = best[1]
    torch.set_num_threads(cores)
    sd = vo.init_state_dict(cfg, seed=0, perturb=False)
    params = [v.clone().requires_grad_(True) for v in sd.values()]
    keys = list(sd.keys())
    opt = torch.optim.Adam(params, lr=1e-3)

    def step(x, y):
        # fp32: CPU bf16 autocast is slower than fp32, so fp32 is the reference's best CPU path
        pd = dict(zip(keys, params))
        loss = vo.cross_entropy(vo.classifier_forward(pd, x, cfg), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    # calibrate the bounded sample (micro-batch) so that (steps + warmup) steps fit the time budget
    x, y = vo.synthetic_batch(cfg, 16, seed=0)
    step(x, y)  # cold step (thread pool spin-up, allocator): not representative
    t0 = time.perf_counter()
    step(x, y)
    per_img = (time.perf_counter() - t0) / 16
    # many-core hosts need a reasonably large micro-batch to use their threads; keep it within the time budget
    mb = int(max(8, min(64, budget_s / max(1e-6, per_img * (steps + warmup)))))
    x, y = vo.synthetic_batch(cfg, mb, seed=1)
    for _ in range(warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, y)
    dt = time.perf_counter() - t0
    return mb * steps / dt, dt / steps * 1e3, mb, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    v, ms, mb, cores = cpu_reference_samples_per_sec(args.steps, max(1, args.warmup), budget_s=150.0)
    sample = f"micro-batch {mb} images x {args.steps} steps, fp32, torch CPU, {cores} threads"
    line = {
        "impl": "reference", "metric": "train samples/sec, ViT-B/16 224px, fwd+bwd+adam step", "value": round(v, 3),
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 classifier 224x224, 1000 classes (BASELINE.json configs[1])", "micro_batch": mb},
        "cpu_baseline": {"value": round(v, 3), "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import cflearn_b200  # noqa: F401
    from cflearn_b200 import _cabi, dp, ops, registry
    from cflearn_b200.optim import ArenaAdam

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    if not _cabi.available():
        raise SystemExit(f"bench.py: {_cabi.load_error()}")
    info = dp.ddp_info()
    rank, world, local = info if info is not None else (0, 1, 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cfg = dict(img_size=224, patch_size=16, in_channels=3, latent_dim=768, num_layers=12, num_classes=1000)
    torch.manual_seed(0)
    model = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=1000, img_size=224, latent_dim=768, encoder="vit",
                                                        encoder_config=dict(patch_size=16, num_layers=12))).to(dev)
    model.arena.ensure()
    if world > 1:
        dp.broadcast_parameters(model)
        dp.attach_reducer(model)
    use_graph = not (args.no_graph or args.overlap_dp)
    if use_graph and world > 1:
        model.engine.reducer = None  # graphed DP step: one all-reduce of the flat gradient arena after the replay
    opt = ArenaAdam(model, lr=1e-3, capturable=use_graph)
    B = PER_GPU_BATCH
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)  # rank r uses its own data seed
    n_host = 2
    host_x = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 1000, (B, 1), generator=g).pin_memory() for _ in range(n_host)]
    dev_x = [h.to(dev) for h in host_x]
    dev_y = [h.to(dev) for h in host_y]
    loss_host = torch.zeros(1).pin_memory()

    gstep = None
    if use_graph:  # zero_grad + fwd + CE + bwd (+ Adam at N = 1) as ONE CUDA graph; N > 1: all-reduce + Adam follow it
        from cflearn_b200.optim import GraphedTrainStep

        model.arena.ensure()
        gstep = GraphedTrainStep(model, opt, B)

    def do_step(x, y):
        if gstep is not None:
            return gstep.step(x, y)
        opt.zero_grad()  # schema.py:984 (backward then overwrites the gradient arena instead of accumulating)
        loss = model.train_step(x, y)
        opt.step()
        return loss

    def step_resident(i):
        return do_step(dev_x[i % n_host], dev_y[i % n_host])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident timing ---------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step_resident(i)
    e1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    if gstep is not None:  # graph replays launch the captured kernels without going through the host-side counter
        launches = gstep.launches_per_replay * args.steps
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = loss.item()
    if not (final_loss == final_loss and abs(final_loss) < 1e4):
        raise SystemExit(f"bench.py: loss diverged ({final_loss})")

    # ---- end to end: pinned host batches -> H2D on a copy stream (prefetched one step ahead) -> step -> loss D2H --
    copy_stream = torch.cuda.Stream()
    stage_x = [torch.empty_like(dev_x[0]) for _ in range(2)]
    stage_y = [torch.empty_like(dev_y[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            stage_x[s].copy_(host_x[i % n_host], non_blocking=True)
            stage_y[s].copy_(host_y[i % n_host], non_blocking=True)
            ready[s].record(copy_stream)

    loss_dev = [torch.zeros(1, device=dev) for _ in range(2)]
    loss_pinned = [torch.zeros(1).pin_memory() for _ in range(2)]
    losses_seen = []

    def e2e_loop(n):
        pending = None
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        prefetch(0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            ls = do_step(stage_x[s], stage_y[s])
            consumed[s].record(torch.cuda.current_stream())
            # device->host read of the loss every step (the reference's per-step .item(), models/common.py:42); the
            # read of step i is waited for after step i+1 has been queued so the GPU never idles on the host
            loss_dev[i % 2].copy_(ls.reshape(1))
            if pending is not None:
                pending[0].synchronize()
                losses_seen.append(float(loss_pinned[pending[1]][0]))
            loss_pinned[i % 2].copy_(loss_dev[i % 2], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            pending = (ev, i % 2)
        if pending is not None:
            pending[0].synchronize()
            losses_seen.append(float(loss_pinned[pending[1]][0]))

    e2e_loop(max(1, min(3, args.warmup)))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    d2h = 4

    # ---- roofline of the dominant kernel: the tcgen05 GEMM (FeedForward up-projection, fused bias+GELU) ----------
    peaks = _peaks()
    M, N, K = B * 197, 3072, 768
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    o0 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BIAS_GELU_BF16, out0=o0, out1=o1)
    torch.cuda.synchronize()
    reps = 20
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        ops.gemm(a, w, bias=bias, epilogue=ops.EPI_BIAS_GELU_BF16, out0=o0, out1=o1)
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / reps
    achieved = 2.0 * M * N * K / (k_ms * 1e-3) / 1e12
    step_tflops = FLOP_PER_IMAGE_FWD_BWD * B / (ms_step * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_kernel<EPI_BIAS_GELU_BF16> 50432x3072x768", "achieved": round(achieved, 1),
        "peak": peaks["burst"], "unit": "TFLOP/s", "frac": round(achieved / peaks["burst"], 4),
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the committed ncu --set full
        # capture (profiles/r01_ncu_summary.md: 82.4 MB + 566.2 MB); algorithmic bytes are 702 MB, so nothing is re-read
        "traffic": 648528640, "traffic_unit": "B/launch (ncu, profiles/r01_ncu_summary.md)",
        "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)",
        "kernel_ms": round(k_ms, 4),
        "step": {"achieved": round(step_tflops, 1), "peak": peaks["sustained"], "frac": round(step_tflops / peaks["sustained"], 4),
                 "note": "whole step, GEMM-only FLOPs 105.38 GFLOP/image, vs sustained cuBLAS bf16 peak"},
    }

    # ---- CPU baseline (rank 0, N == 1 only): the oracle on the host cores, bounded sample -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, mb, cores = cpu_reference_samples_per_sec(2, 1, budget_s=20.0)
        cpu = {"value": round(v, 3), "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"micro-batch {mb} images x 2 steps (fp32 oracle port of the reference path, torch CPU, {cores} threads)"}

    if rank == 0:
        line = {
            "metric": "train samples/sec, ViT-B/16 224px, fwd+bwd+adam step", "value": round(value, 1), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ViT-B/16 classifier 224x224, 1000 classes, batch 256 per GPU (BASELINE.json configs[1]/[2])",
                       "global_batch": world * B, "seq_len": 197, "parallelism": f"dp{world}",
                       "optimizer": "adam (fused arena kernel, inside the timed region)",
                       "cuda_graph": bool(use_graph),
                       "l2": "per-step working set (> 15 GB of activations) exceeds the 126 MB L2; no explicit flush needed"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": round(e2e_ms / args.steps, 3)},
            "roofline": roofline, "cpu_baseline": cpu, "loss": round(final_loss, 4),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--overlap-dp", action="store_true", help="N > 1: eager launches with the bucketed, overlapped all-reduce (implies --no-graph)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
