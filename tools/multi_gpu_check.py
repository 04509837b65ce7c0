#!/usr/bin/env python
"""Multi-GPU parity check of the H-sharded KernelWeighting (run under torchrun,
one rank per GPU, NCCL): every rank computes the unsharded op on its own GPU and
its band through ShardedKernelWeighting (halo exchange + halo reduce over NCCL),
then compares.  Prints one line per rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist

import sbmc_b200.functions as funcs
from sbmc_b200 import sharding


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (n, c, h, w, k) in [(2, 3, 64 * world, 256, 21), (1, 3, 37 * world + 3, 132, 5),
                            (1, 3, 270 * world, 3840, 21)]:
        g = th.Generator(device=dev).manual_seed(5)          # same tensors on all ranks
        data = (2 * th.randn(n, c, h, w, device=dev, generator=g)).requires_grad_(True)
        weights = th.randn(n, k, k, h, w, device=dev, generator=g).requires_grad_(True)
        d_out = th.randn(n, c, h, w, device=dev, generator=g)
        d_sw = th.randn(n, h, w, device=dev, generator=g)
        out, sw = funcs.KernelWeighting.apply(data, weights)
        th.autograd.backward([out, sw], [d_out, d_sw])
        plan = sharding.BandPlan(h, world, k)
        db = plan.band(rank, data.detach(), 2).contiguous().requires_grad_(True)
        wb = plan.band(rank, weights.detach(), 3).contiguous().requires_grad_(True)
        ob, sb = sharding.ShardedKernelWeighting.apply(db, wb, plan, rank)
        th.autograd.backward([ob, sb], [plan.band(rank, d_out, 2).contiguous(),
                                        plan.band(rank, d_sw, 1).contiguous()])
        full_out = sharding.gather_bands(plan, rank, ob.detach(), 2)
        checks = {
            "out": th.equal(ob, plan.band(rank, out, 2)),
            "sum_w": th.equal(sb, plan.band(rank, sw, 1)),
            "d_weights": th.equal(wb.grad, plan.band(rank, weights.grad, 3)),
            "d_data": (db.grad - plan.band(rank, data.grad, 2)).abs().max().item()
            <= 1e-5 * data.grad.abs().max().item(),
            "gather": th.equal(full_out, out.detach()),
        }
        ok = ok and all(checks.values())
        print("rank %d/%d shape %s: %s" % (rank, world, (n, c, h, w, k), checks), flush=True)
        del data, weights, out, sw, db, wb, ob, sb
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
