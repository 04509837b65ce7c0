#!/usr/bin/env python
"""Times the pieces of the sharded KernelWeighting call (torchrun, NCCL)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist
from sbmc_b200 import sharding, halide_ops


def timeit(name, fn, iters=50, rank=0):
    for _ in range(5):
        fn()
    th.cuda.synchronize(); dist.barrier(); th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    t_cpu = (time.perf_counter() - t0) / iters * 1e3
    th.cuda.synchronize()
    if rank == 0:
        print("%-40s gpu %.3f ms/iter   cpu-enqueue %.3f ms/iter" % (name, a.elapsed_time(b) / iters, t_cpu), flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, C, H, W, K = 4, 3, 720, 1280, 21
    plan = sharding.BandPlan(H * world, world, K)
    data = th.randn(B, C, H, W, device=dev)
    weights = th.randn(B, K, K, H, W, device=dev)
    d_out = th.randn(B, C, H, W, device=dev); d_sw = th.randn(B, H, W, device=dev)
    out = th.empty_like(data); sum_w = th.empty(B, H, W, device=dev)
    d_data = th.empty_like(data); d_weights = th.empty_like(weights)
    ext = sharding.exchange_halo(plan, rank, data)
    dext = th.empty_like(ext)
    timeit("exchange_halo", lambda: sharding.exchange_halo(plan, rank, data), rank=rank)
    timeit("reduce_halo", lambda: sharding.reduce_halo(plan, rank, dext, out=d_data), rank=rank)
    timeit("fwd unsharded", lambda: halide_ops.kernel_weighting_cuda_float32(data, weights, out, sum_w), rank=rank)
    timeit("fwd sharded (ext given)", lambda: sharding.kernel_weighting_fwd_sharded(plan, rank, data, weights, out, sum_w, data_ext=ext), rank=rank)
    timeit("fwd sharded (with exchange)", lambda: sharding.kernel_weighting_fwd_sharded(plan, rank, data, weights, out, sum_w), rank=rank)
    timeit("bwd unsharded", lambda: halide_ops.kernel_weighting_grad_cuda_float32(data, weights, sum_w, d_out, d_sw, d_data, d_weights), rank=rank)
    timeit("bwd sharded (ext given)", lambda: sharding.kernel_weighting_bwd_sharded(plan, rank, data, weights, d_out, d_sw, d_data, d_weights, data_ext=ext), rank=rank)

    def both():
        e = sharding.kernel_weighting_fwd_sharded(plan, rank, data, weights, out, sum_w)
        sharding.kernel_weighting_bwd_sharded(plan, rank, data, weights, d_out, d_sw, d_data, d_weights, data_ext=e)
    timeit("fwd+bwd sharded", both, rank=rank)

    def both_un():
        halide_ops.kernel_weighting_cuda_float32(data, weights, out, sum_w)
        halide_ops.kernel_weighting_grad_cuda_float32(data, weights, sum_w, d_out, d_sw, d_data, d_weights)
    timeit("fwd+bwd unsharded", both_un, rank=rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
