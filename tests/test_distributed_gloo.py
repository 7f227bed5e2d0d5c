"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat gradient all-reduce, parameter
broadcast and batch sharding.  The CUDA kernels are not involved (they need a GPU); what is checked is that
two ranks with different local gradients end up with the DDP average and identical parameters."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cim_quantization_b200.distributed import FlatGradAllReducer, broadcast_parameters, shard_batch
        import cim_quantization_b200 as cq
        torch.manual_seed(100 + rank)  # different init per rank on purpose
        m = cq.Conv2dLSQCiM(4, 16, (3, 3), 1, 1, 1, 1, False, nbits_w=3, nbits_a=3, xbar=16, adcbits=1.5)
        broadcast_parameters(m, 0)
        ref = [p.detach().clone() for p in m.parameters()]
        gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
        dist.all_gather(gathered, ref[0])
        same = all(torch.equal(gathered[0], g) for g in gathered)
        # local "gradients": rank-dependent constants
        for i, p in enumerate(m.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        red = FlatGradAllReducer(m.parameters())
        red.all_reduce_()
        ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(m.parameters()))
        x = torch.arange(8 * 3).view(8, 3)
        sh = shard_batch(x, rank, world)
        # lazy-init statistics reduced over the global batch: every rank gets the value one process
        # would compute on the concatenation of the shards
        from cim_quantization_b200 import distributed as CD
        full = torch.linspace(-1.0, 3.0, 16)
        mine = shard_batch(full, rank, world)
        gmean = CD.global_mean_(mine.abs().mean())
        gmin = CD.global_min_(mine.min())
        gsum, ranks = CD.global_sum_(mine.to(torch.int64).abs().sum())
        stats_ok = (torch.allclose(gmean, full.abs().mean()) and gmin.item() == full.min().item()
                    and gsum.item() == full.to(torch.int64).abs().sum().item() and ranks == world)
        CD.set_sync_lazy_init(False)
        local = CD.global_mean_(mine.abs().mean())
        stats_ok = stats_ok and torch.allclose(local, mine.abs().mean())
        CD.set_sync_lazy_init(True)
        ret[rank] = (same, ok, sh[0, 0].item(), red.nbytes, stats_ok)
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_two_ranks():
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        same, ok, first, nbytes, stats_ok = ret[rank]
        assert stats_ok, "lazy-init statistics are not the global-batch values"
        assert same, "parameters differ across ranks after broadcast"
        assert ok, "all-reduced gradient is not the DDP average"
        assert first == rank * 12
        assert nbytes > 0


def test_single_process_reducer_is_identity():
    from cim_quantization_b200.distributed import FlatGradAllReducer
    p = torch.nn.Parameter(torch.ones(5))
    p.grad = torch.arange(5.0)
    r = FlatGradAllReducer([p])
    r.all_reduce_()
    assert torch.equal(p.grad, torch.arange(5.0))
    with pytest.raises(ValueError):
        FlatGradAllReducer([])
