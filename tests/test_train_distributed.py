"""Data-parallel training step (interfaces.SampleBasedDenoiserInterface(distributed=True))
over gloo, world size 2, on CPU: two ranks with different batches end up with identical
parameters, equal to one process training on the union of the batches."""
import os
import socket

import torch as th
import torch.multiprocessing as mp

from sbmc_b200 import interfaces


class _Stub(th.nn.Module):
    """Stands in for Multisteps: dict in, {"radiance": ...} out, a few parameter tensors."""

    def __init__(self):
        super(_Stub, self).__init__()
        self.a = th.nn.Conv2d(3, 8, 3, padding=1)
        self.b = th.nn.Conv2d(8, 3, 3, padding=1)

    def forward(self, batch):
        x = batch["radiance"].mean(1)
        return {"radiance": self.b(th.relu(self.a(x)))}


def _batches(seed, n):
    g = th.Generator().manual_seed(seed)
    return [{"radiance": th.rand(2, 2, 3, 12, 12, generator=g),
             "target_image": th.rand(2, 3, 12, 12, generator=g)} for _ in range(n)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    th.manual_seed(100 + rank)                   # different initial weights: rank 0's win
    iface = interfaces.SampleBasedDenoiserInterface(_Stub(), lr=1e-2, distributed=True)
    for b in _batches(7 + rank, 3):
        iface.train_step(b)
    th.save([p.detach().clone() for p in iface.model.parameters()], os.path.join(out, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_data_parallel_step_matches_training_on_the_union(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = th.load(str(tmp_path / "rank0.pt"))
    r1 = th.load(str(tmp_path / "rank1.pt"))
    for a, b in zip(r0, r1):
        assert th.equal(a, b)
    # one process, batches concatenated (mean loss over equal-sized halves = mean of the means)
    th.manual_seed(100)
    ref = interfaces.SampleBasedDenoiserInterface(_Stub(), lr=1e-2)
    for b0, b1 in zip(_batches(7, 3), _batches(8, 3)):
        ref.train_step({k: th.cat([b0[k], b1[k]]) for k in b0})
    for a, p in zip(r0, ref.model.parameters()):
        assert th.allclose(a, p, rtol=1e-4, atol=1e-6), (a - p).abs().max()
