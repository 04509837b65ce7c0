#!/usr/bin/env python
"""Headline benchmark: KernelWeighting forward+backward throughput.

Metric (BASELINE.json): Msamples/s, samples = B*spp*H*W per step, one step =
the reference call pattern for one batch: `spp` calls of KernelWeighting
forward + backward on data [B,3,H,W] x weights [B,K,K,H,W] (the model calls the
op once per sample index, sbmc/models.py:195-206 of the reference).
Default workload = BASELINE.json configs[1]: B=4, spp=8, 1280x720, K=21, fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (launched by torch.distributed.run, one rank per GPU): the image is cut
into N row bands (H-sharding); every rank owns a band of
`--h` rows (weak scaling: the image is N*h rows tall), exchanges the K-1 halo
rows of `data` (forward) and of `d_data` (backward) with its neighbours over
NCCL, and the rank-0 line reports the whole-job Msamples/s (max over ranks).

One JSON line is printed by rank 0; see DESIGN.md section "Measurement" for the
meaning of every key (roofline / cpu_baseline / e2e / clocks / gpu_launches).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "Msamples/s"


def _metric_name():
    """BASELINE.json's metric string (the driver matches on it)."""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as fid:
            return json.load(fid)["metric"]
    except (OSError, ValueError, KeyError):
        return ("Msamples/s (spp\u00d7H\u00d7W) KernelWeighting fwd+bwd @ spp=8 K=21 720p, "
                "1/2/4/8 GPU")


METRIC = _metric_name()


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="kernel_weighting", choices=["kernel_weighting", "tiles"],
                   help="kernel_weighting: the headline metric (default).  tiles: the tile "
                        "reader side benchmark (benchmarks/tiles_bench.py) with its CPU leg")
    p.add_argument("--tile-size", type=int, default=80)
    p.add_argument("--b", type=int, default=4)
    p.add_argument("--spp", type=int, default=8)
    p.add_argument("--h", type=int, default=720)
    p.add_argument("--w", type=int, default=1280)
    p.add_argument("--k", type=int, default=21)
    p.add_argument("--c", type=int, default=3)
    p.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    p.add_argument("--no-train-e2e", dest="no_train_e2e", action="store_true",
                   help="skip the end-to-end training line of the secondary block")
    p.add_argument("--no-secondary", action="store_true",
                   help="skip the configs 3 / 4 / 5 block (callers of the hot path)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=1)
    p.add_argument("--cpu-seconds", type=float, default=12.0,
                   help="target CPU time of the cpu_baseline sample")
    return p.parse_args()


def workload(a):
    return {"workload": "KernelWeighting fwd+bwd, B=%d spp=%d %dx%d K=%d C=%d fp32 "
                        "(BASELINE configs[1]; %d calls of [B,C,H,W]x[B,K,K,H,W] per step)"
                        % (a.b, a.spp, a.w, a.h, a.k, a.c, a.spp),
            "B": a.b, "spp": a.spp, "H": a.h, "W": a.w, "K": a.k, "C": a.c,
            "l2_policy": "inputs larger than L2 (6.5 GB of weights per call, a "
                         "different weight buffer per sample index)"}


# --------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# --------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if t0 is not None:                       # samples taken inside the timed region
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.06]
            rows = inside or rows
        for _, row in rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, power) if p >= 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(smax),
                "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# the reference arm / cpu_baseline: the CPU restatement on the host cores
# --------------------------------------------------------------------------
def host_cores():
    """The host cores this process may run on (the affinity mask, not what a
    launcher wrote into OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuArm:
    """The ONE routine both CPU legs time (`cpu_baseline` of the B200 line and
    `--impl reference`): the oracle's fwd+bwd on one image of the workload
    ([1,C,H,W] x [1,K,K,H,W]) with every host core, whatever OMP_NUM_THREADS says."""

    def __init__(self, a):
        import torch as th
        import oracle
        self.a, self.oracle = a, oracle
        oracle.set_num_threads(host_cores())
        self.cores = oracle.num_threads()
        th.manual_seed(0)
        self.data = 2 * th.randn(1, a.c, a.h, a.w)
        self.weights = th.randn(1, a.k, a.k, a.h, a.w)
        self.d_out = th.randn(1, a.c, a.h, a.w)
        self.d_sw = th.randn(1, a.h, a.w)
        self.out = th.empty_like(self.data)
        self.sum_w = th.empty(1, a.h, a.w)
        self.d_data = th.empty_like(self.data)
        self.d_weights = th.empty_like(self.weights)
        self.once()                              # page faults, thread pool
        self.once()

    def once(self):
        o = self.oracle
        t0 = time.perf_counter()
        o.kernel_weighting_cpu_float32(self.data, self.weights, self.out, self.sum_w)
        o.kernel_weighting_grad_cpu_float32(self.data, self.weights, self.sum_w, self.d_out,
                                            self.d_sw, self.d_data, self.d_weights)
        return time.perf_counter() - t0

    def run(self, reps):
        """Median seconds per image over `reps` repetitions."""
        ts = sorted(self.once() for _ in range(reps))
        return ts[len(ts) // 2]

    def describe(self, reps):
        a = self.a
        return ("each step = 1 image of the workload ([1,%d,%d,%d]x[1,%d,%d,%d,%d], fwd+bwd) "
                "instead of B*spp=%d; median of %d; %d OpenMP threads; C restatement of the "
                "Halide CPU schedule (oracle/sbmc_oracle.c)"
                % (a.c, a.h, a.w, a.k, a.k, a.h, a.w, a.b * a.spp, reps, self.cores))


def cpu_sample(a, target_seconds, reps_min=3):
    """(Msamples/s, cores, description, seconds per image) of the CPU arm."""
    arm = CpuArm(a)
    one = arm.once()
    reps = max(reps_min, min(50, int(target_seconds / max(one, 1e-3))))
    dt = arm.run(reps)
    return a.h * a.w / dt / 1e6, arm.cores, arm.describe(reps), dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(a)
    for _ in range(a.warmup):
        arm.once()
    dt = arm.run(a.steps)
    value = a.h * a.w / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores,
                         "kind": "port", "sample": arm.describe(a.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "Halide (the reference's code generator) cannot be built here; this is "
                "the schedule-faithful C/OpenMP restatement of its CPU path",
    }))


# --------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------
def run_b200(a):
    import torch as th
    import torch.distributed as dist
    from sbmc_b200 import _lib, halide_ops
    from sbmc_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not th.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, C, H, W, K, spp = a.b, a.c, a.h, a.w, a.k, a.spp
    gen = th.Generator(device=dev).manual_seed(1234 + rank)
    plan = sharding.BandPlan(H * world, world, K, K) if world > 1 else None
    parity = sharded_parity_check(th, dist, sharding, halide_ops, dev, rank, world) \
        if world > 1 else None
    # resident inputs: one weight buffer per sample index (spp x 6.5 GB at config 2)
    if plan is None:
        data = [2 * th.randn(B, C, H, W, device=dev, generator=gen) for _ in range(spp)]
    else:
        # sharded: `data` / `d_data` live in their halo-extended layout (the band is a
        # view), the exchange runs on a side stream (sharding.HaloPipeline)
        pipe = sharding.HaloPipeline(plan, rank, (B, C, H, W), dev)
        data = [pipe.new_ext() for _ in range(spp)]
        for d in data:
            d.zero_()
            pipe.band(d).normal_(generator=gen).mul_(2)
        d_ext = [pipe.new_ext(), pipe.new_ext()]
    weights = []
    for _ in range(spp):
        wt = th.empty(B, K, K, H, W, device=dev)
        for b in range(B):                       # bounded temporaries
            wt[b].normal_(generator=gen)
        weights.append(wt)
    d_out = th.randn(B, C, H, W, device=dev, generator=gen)
    d_sw = th.randn(B, H, W, device=dev, generator=gen)
    out = th.empty(B, C, H, W, device=dev)
    sum_w = th.empty(B, H, W, device=dev)
    d_data = th.empty(B, C, H, W, device=dev)
    d_weights = th.empty(B, K, K, H, W, device=dev)

    def step():
        if plan is None:
            for s in range(spp):
                halide_ops.kernel_weighting_cuda_float32(data[s], weights[s], out, sum_w)
                halide_ops.kernel_weighting_grad_cuda_float32(
                    data[s], weights[s], sum_w, d_out, d_sw, d_data, d_weights)
            return
        # every call exchanges its own halos (the op cannot know that the bench reuses
        # its inputs); the exchange of call s + 1 and the d_data reduction of call s run
        # on the side stream while the kernels of the neighbouring calls execute
        tok = pipe.exchange_async(data[0])
        red = [None, None]
        for s in range(spp):
            pipe.wait(tok)
            if s + 1 < spp:
                tok = pipe.exchange_async(data[s + 1])
            sharding.kernel_weighting_fwd_band(plan, rank, data[s], weights[s], out, sum_w)
            pipe.wait(red[s & 1])                       # d_ext[s & 1] is free again
            sharding.kernel_weighting_bwd_band(plan, rank, data[s], weights[s], d_out, d_sw,
                                               d_ext[s & 1], d_weights)
            red[s & 1] = pipe.reduce_async(d_ext[s & 1])
        pipe.wait(red[0])
        pipe.wait(red[1])

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                  # runs through warm-up and the timed region
    for _ in range(a.warmup):
        step()
    barrier()
    _lib.timing_collect()
    _lib.timing_enable(True)
    launches0 = _lib.launch_count()
    ev0, ev1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    _lib.timing_enable(False)
    kernels = _lib.timing_collect()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    t = th.tensor([ms], device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / a.steps
    samples_step = B * spp * H * W * world
    value = samples_step / (ms_step * 1e-3) / 1e6

    # roofline of the dominant kernel (largest share of the step)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks \
        else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    per_sample = {"kw_fwd": 4 * (K * K + 2 * C + 1),            # W + D -> out + sum_w
                  "kw_bwd_dweights": 4 * (K * K + 2 * C + 1),   # D + dO + dSw -> dW
                  "kw_bwd_ddata": 4 * (K * K + 2 * C)}          # W + dO -> dD
    launch_samples = B * H * W
    table = {}
    for name, (kms, cnt) in kernels.items():
        if name in per_sample and cnt:
            gbs = per_sample[name] * launch_samples / (kms / cnt * 1e-3) / 1e9
            table[name] = {"launches": cnt, "avg_ms": kms / cnt, "GB/s": gbs,
                           "frac": gbs / peak, "share_of_step": kms / (ms_step * a.steps)}
    roofline = None
    if table:
        top = max(table, key=lambda n: table[n]["avg_ms"] * table[n]["launches"])
        traffic = None
        if (B, C, H, W, K) == (4, 3, 720, 1280, 21):     # the shape ncu captured
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[top]["bytes"]
            except (OSError, ValueError, KeyError):
                pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": table[top]["GB/s"],
                    "peak": peak, "unit": "GB/s", "frac": table[top]["frac"],
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": per_sample[top] * launch_samples,
                    "step_frac": 4 * (2 * K * K + K * K + 5 * C + 2) * samples_step / world
                    / (ms_step * 1e-3) / 1e9 / peak,
                    "kernels": table}

    # e2e: the same step through the host-buffer entry points (pinned host memory)
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, th, halide_ops, dev, world, dist)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cores, desc, _ = cpu_sample(a, a.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}

    # free the headline's buffers, then the callers of the hot path (configs 3, 4, 5)
    del weights, d_weights, data, out, sum_w, d_data, d_out, d_sw
    th.cuda.empty_cache()
    secondary = None
    if not a.no_secondary:
        secondary = run_secondary(a, th, dist, dev, rank, world)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload(a),
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "e2e": e2e, "cpu_baseline": cpu, "secondary": secondary,
            "tolerance": "fp32 vs the CPU oracle: |got-ref| <= 1e-5 (|ref| + sum|terms|) per "
                         "element and ||got-ref|| <= 1e-5 ||ref|| (tests/util.py); "
                         "scatter2gather bit-exact",
        }
        if parity is not None:
            line["parity_ok"] = parity["ok"]
            line["parity"] = parity
        if world > 1:
            line["config"]["parallelism"] = (
                "H-sharding: %d row bands of %d rows, NCCL halo exchange of %d data rows "
                "(fwd) and d_data rows (bwd) per call, on a side stream" % (world, H, K - 1))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sharded_parity_check(th, dist, sharding, halide_ops, dev, rank, world):
    """Every rank: the sharded op (row bands + halo pipeline) on a small image against
    the unsharded op on the whole image (same seeded inputs on every rank)."""
    n, c, k, w = 2, 3, 21, 256
    h = 32 * world
    g = th.Generator(device=dev).manual_seed(99)
    data = 2 * th.randn(n, c, h, w, device=dev, generator=g)
    weights = th.randn(n, k, k, h, w, device=dev, generator=g)
    d_out = th.randn(n, c, h, w, device=dev, generator=g)
    d_sw = th.randn(n, h, w, device=dev, generator=g)
    out, sum_w = th.empty_like(data), th.empty(n, h, w, device=dev)
    d_data, d_weights = th.empty_like(data), th.empty_like(weights)
    halide_ops.kernel_weighting_cuda_float32(data, weights, out, sum_w)
    halide_ops.kernel_weighting_grad_cuda_float32(data, weights, sum_w, d_out, d_sw, d_data,
                                                  d_weights)
    plan = sharding.BandPlan(h, world, k, k)
    y0, y1 = plan.y0[rank], plan.y1[rank]
    rows = y1 - y0
    pipe = sharding.HaloPipeline(plan, rank, (n, c, rows, w), dev)
    ext, dext = pipe.new_ext(), pipe.new_ext()
    pipe.band(ext).copy_(data[:, :, y0:y1])
    pipe.wait(pipe.exchange_async(ext))
    wb = weights[:, :, :, y0:y1].contiguous()
    ob, sb = th.empty(n, c, rows, w, device=dev), th.empty(n, rows, w, device=dev)
    dwb = th.empty_like(wb)
    sharding.kernel_weighting_fwd_band(plan, rank, ext, wb, ob, sb)
    sharding.kernel_weighting_bwd_band(plan, rank, ext, wb, d_out[:, :, y0:y1].contiguous(),
                                       d_sw[:, y0:y1].contiguous(), dext, dwb)
    pipe.wait(pipe.reduce_async(dext))
    th.cuda.synchronize()

    def rel(got, ref):
        return ((got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)).item()
    errs = {"output": rel(ob, out[:, :, y0:y1]), "sum_w": rel(sb, sum_w[:, y0:y1]),
            "d_weights": rel(dwb, d_weights[:, :, :, y0:y1]),
            "d_data": rel(pipe.band(dext), d_data[:, :, y0:y1])}
    worst = th.tensor([max(errs.values())], device=dev, dtype=th.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {"ok": bool(worst.item() <= 1e-5), "max_rel_err_over_ranks": worst.item(),
            "check": "sharded (bands + side-stream halo exchange) vs unsharded op, "
                     "[%d,%d,%d,%d] K=%d, every rank, norm-wise" % (n, c, h, w, k)}


def _timed_cuda(th, fn, warmup, steps):
    for _ in range(warmup):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_secondary(a, th, dist, dev, rank, world):
    """BASELINE.json configs 3, 4 (rank 0, one GPU) and 5 (all ranks, N > 1): the
    callers of the hot path, timed after the headline.  Not part of `value`."""
    from sbmc_b200 import _lib, interfaces, models, sharding
    sec = {}
    th.manual_seed(0)
    try:
        if rank == 0:
            # config 3: Multisteps(93, 3) eval forward, spp=4, 1280x720, bf16 convs + fp32 splat
            net = models.Multisteps(93, 3).to(dev).eval().to(memory_format=th.channels_last)
            net.bf16_chains = net.bf16_unet = True
            batch = {"radiance": th.rand(1, 4, 3, 720, 1280, device=dev),
                     "features": th.randn(1, 4, 93, 720, 1280, device=dev),
                     "global_features": th.randn(1, 3, 1, 1, device=dev)}

            def fwd():
                with th.no_grad():
                    return net(batch)["radiance"]
            l0 = _lib.launch_count()
            ms = _timed_cuda(th, fwd, 2, 5)
            sec["config3_forward"] = {
                "ms": ms, "Msamples_per_s": 4 * 720 * 1280 / ms / 1e3,
                "workload": "Multisteps(93,3) eval forward, spp=4, 1280x720, K=21, 1 GPU",
                "path": "bf16 NHWC pipeline: pipelined tcgen05 1x1 chains (+ fused sample mean), "
                        "tcgen05 3x3 implicit-GEMM U-net convs (bias+act fused), fused fp32 splat",
                "repo_launches_per_forward": (_lib.launch_count() - l0) / 7.0}
            del net, batch
            th.cuda.empty_cache()
            # config 4: train step B=8, spp=8, 128x128, K=21: fwd + loss + bwd + clip + Adam,
            # once in strict fp32 (the reference's arithmetic) and once with cuDNN's TF32
            # convolution paths allowed (PyTorch's default; labelled, never silent)
            batch = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev),
                     "features": th.randn(8, 8, 93, 128, 128, device=dev),
                     "global_features": th.randn(8, 3, 1, 1, device=dev),
                     "target_image": th.rand(8, 3, 128, 128, device=dev)}
            sec["config4_train_step"] = {
                "workload": "Multisteps(93,3) train step, B=8, spp=8, 128x128, K=21, "
                            "fwd + TonemappedRelativeMSE + bwd + clip + Adam, 1 GPU",
                "path": "fp32_strict / tf32_convs_allowed: the reference's op chain (cuDNN "
                        "convolutions) with the repo's fused splat forward / backward and fused "
                        "clip+Adam; bf16_unet_*: U-net forward and data-gradient convolutions on "
                        "the repo's tcgen05 kernel; bf16_train_pipeline*: sbmc_b200/"
                        "train_pipeline.py -- every GEMM-shaped pass except the 3x3 weight "
                        "gradients on repo tcgen05 kernels"}
            for label, tf32, own, graph in (
                    ("fp32_strict", False, 0, False),
                    ("tf32_convs_allowed", True, 0, False),
                    ("tf32_convs_allowed_cuda_graph", True, 0, True),
                    ("bf16_unet_own_kernels_rest_fp32_strict", False, 1, False),
                    ("bf16_unet_own_kernels_rest_tf32", True, 1, False),
                    ("bf16_train_pipeline", False, 2, False),
                    ("bf16_train_pipeline_cuda_graph", False, 2, True)):
                th.manual_seed(0)
                net = models.Multisteps(93, 3).to(dev).train()
                net.bf16_unet_train = own == 1
                net.bf16_train = own == 2
                iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True,
                                                                fused_optimizer=True,
                                                                allow_tf32=tf32, cuda_graph=graph)

                def train():
                    return iface.train_step(batch)
                ms = _timed_cuda(th, train, 2, 4)
                sec["config4_train_step"][label] = {
                    "ms": ms, "Msamples_per_s": 8 * 8 * 128 * 128 / ms / 1e3,
                    "precision": ("fp32 storage and accumulation, TF32 %s for cuDNN / cuBLAS"
                                  % ("ALLOWED" if tf32 else "off"))
                    + ("; U-nets in bf16 (fp32 accumulate): forward and data-gradient "
                       "convolutions on csrc/conv3x3.cu, weight gradients on cuDNN bf16"
                       if own else "")
                    + ("; 1x1 chains in bf16: forward and data-gradient layers on "
                       "csrc/linear.cu, weight / bias gradients on csrc/wgrad.cu; 3x3 weight "
                       "gradients on cuDNN bf16; no fp32 convolution left" if own == 2 else "")
                    + ("; whole step replayed from one CUDA graph" if graph else "")}
                del net, iface
            th.backends.cudnn.allow_tf32 = False
            th.backends.cuda.matmul.allow_tf32 = False
            del batch
            th.cuda.empty_cache()
    except Exception as exc:                      # the headline line must still be printed
        sec["error_config34"] = "%s: %s" % (type(exc).__name__, exc)
    try:
        if rank == 0 and not getattr(a, "no_train_e2e", False):
            # config 4 END TO END through the callers' own loop: 128 x 128 x 8 spp tiles on disk
            # (the renderer's .bin + LZ4 format, synthetic content) -> PrefetchLoader (GPU inflate
            # + assembly, 16 batches decoded per launch on a side stream) -> training step
            import contextlib
            import io
            from benchmarks import train_e2e_bench
            with contextlib.redirect_stdout(io.StringIO()):
                line = train_e2e_bench.main(["--prefetch", "16", "--tiles", "5", "--repeat", "48",
                                             "--steps", "64"])
            sec["config4_train_end_to_end"] = {
                k: line[k] for k in ("end_to_end_ms_per_step", "end_to_end_Msamples_per_s",
                                     "loader_only_ms_per_batch", "step_only_ms")}
            sec["config4_train_end_to_end"]["workload"] = line["config"]["workload"] + "; " + \
                line["config"]["path"] + "; device_prefetch 16; steady state of one epoch"
            th.cuda.empty_cache()
    except Exception as exc:
        sec["error_config4_e2e"] = "%s: %s" % (type(exc).__name__, exc)
    try:
        if world > 1:
            # config 5: tiled inference of one 3840x2160 spp=8 frame, one row band per rank,
            # halo exchange before every U-net + final gather (strong scaling of one frame)
            H5, W5, spp5 = 2160, 3840, 8
            net = models.Multisteps(93, 3).to(dev).eval().to(memory_format=th.channels_last)
            net.bf16_chains = net.bf16_unet = True
            for prm in net.parameters():
                dist.broadcast(prm.data, 0)
            lo, hi = sharding.halo_mode_rows(H5, world, net.ksize, rank)
            g = th.Generator(device=dev).manual_seed(7 + rank)
            band = {"radiance": th.rand(1, spp5, 3, hi - lo, W5, device=dev, generator=g),
                    "features": th.randn(1, spp5, 93, hi - lo, W5, device=dev, generator=g),
                    "global_features": th.randn(1, 3, 1, 1, device=dev)}

            def frame():
                with th.no_grad():
                    return sharding.multisteps_forward_halo(net, band, rank, world,
                                                            image_height=H5, row0=lo)["radiance"]
            frame()
            th.cuda.synchronize()
            dist.barrier()
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                frame()
            e1.record()
            th.cuda.synchronize()
            t = th.tensor([e0.elapsed_time(e1) / 2], device=dev, dtype=th.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec["config5_tiled_inference"] = {
                "s_per_frame": t.item() / 1e3, "Msamples_per_s": spp5 * H5 * W5 / t.item() / 1e3,
                "workload": "Multisteps(93,3) 3840x2160 spp=8 K=21 on %d GPUs: row bands, NCCL "
                            "halo exchange before every U-net, final all-gather; band inputs "
                            "resident per rank; max over ranks" % world}
            del net, band
            th.cuda.empty_cache()
    except Exception as exc:
        sec["error_config5"] = "%s: %s" % (type(exc).__name__, exc)
    return sec if rank == 0 else None


def run_e2e(a, th, halide_ops, dev, world, dist):
    """One step = spp calls of fwd+bwd through the *_cpu_float32 entry points
    (HOST tensors, pinned): every call streams its inputs H2D and its outputs
    D2H inside the timed region."""
    B, C, H, W, K, spp = a.b, a.c, a.h, a.w, a.k, a.spp
    pin = dict(pin_memory=True)
    data = th.randn(B, C, H, W, **pin).mul_(2)
    weights = th.empty(B, K, K, H, W, **pin)
    weights[0].normal_()
    for b in range(1, B):                 # values do not matter for the timing
        weights[b].copy_(weights[0])
    d_out = th.randn(B, C, H, W, **pin)
    d_sw = th.randn(B, H, W, **pin)
    out = th.empty(B, C, H, W, **pin)
    sum_w = th.empty(B, H, W, **pin)
    d_data = th.empty(B, C, H, W, **pin)
    d_weights = th.empty(B, K, K, H, W, **pin)

    def step():
        for _ in range(spp):
            halide_ops.kernel_weighting_cpu_float32(data, weights, out, sum_w)
            halide_ops.kernel_weighting_grad_cpu_float32(
                data, weights, sum_w, d_out, d_sw, d_data, d_weights)

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    # warm-up: one call pair (allocates the staging buffers)
    halide_ops.kernel_weighting_cpu_float32(data, weights, out, sum_w)
    halide_ops.kernel_weighting_grad_cpu_float32(data, weights, sum_w, d_out, d_sw,
                                                 d_data, d_weights)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    t = th.tensor([dt], device=dev, dtype=th.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = t.item() / a.e2e_steps
    nb = lambda x: x.numel() * 4
    h2d = spp * (2 * nb(weights) + 2 * nb(data) + nb(d_out) + nb(d_sw))
    d2h = spp * (nb(out) + nb(sum_w) + nb(d_data) + nb(d_weights))
    return {"value": B * spp * H * W * world / dt / 1e6, "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": a.e2e_steps, "ms_per_step": dt * 1e3,
            "api": "sbmc_b200.halide_ops.kernel_weighting{,_grad}_cpu_float32 "
                   "(sbmc_kernel_weighting_{fwd,bwd}_host_f32), pinned host tensors"}


def run_tiles(a):
    """Side benchmark of the tile reader.  The CPU leg (the reference's reader
    restated: oracle/tiles_ref.py + oracle/lz4_oracle.c, single thread) lives here
    because bench.py is the one benchmark that may execute oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import tiles_bench

    def cpu_seconds_per_tile(bufs):
        from oracle import tiles_ref
        t0 = time.perf_counter()
        for b in bufs:
            tiles_ref.read_tile(b)
        return (time.perf_counter() - t0) / len(bufs)

    argv = ["--w", str(a.w), "--h", str(a.h), "--ts", str(a.tile_size), "--spp", str(a.spp),
            "--steps", str(min(a.steps, 5))]
    tiles_bench.main(argv, None if a.no_cpu_baseline else cpu_seconds_per_tile)


def main():
    a = parse_args()
    if a.workload == "tiles":
        run_tiles(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
