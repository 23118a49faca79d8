"""Multi-GPU host logic on CPU: world_size-2 `gloo` process group.
Covers ray sharding, the single flat all-reduce bucket, max/sum over ranks and the distributed
branch of SGD_Optimizer.process_gradient (per-rank partial gradients + error statistics summed,
every rank ending with the same processed gradient).  The trace itself needs a GPU; here a stub
engine evaluates a differentiable per-ray error on the rank's shard so the reduction logic is
exercised end to end."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorflowraytrace_b200 import optimizer, parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


class _StubSystem:
    def update(self):
        pass


class _StubEngine:
    """Stands in for OpticalEngine: `rays` are this rank's shard of per-ray coefficients."""

    def __init__(self, rays):
        self.rays = rays
        self.optical_system = _StubSystem()

    def clear_ray_history(self):
        pass

    def ray_trace(self, depth):
        pass


def _error_function(params):
    def f(engine):
        a, b = params
        r = engine.rays
        return (r[:, 0] * a.sum() - r[:, 1]) ** 2 + (b ** 2).sum() * r[:, 2] * r[:, 0]
    return f


def _make_problem():
    g = torch.Generator().manual_seed(5)
    rays = torch.rand((1001, 3), generator=g, dtype=torch.float64)
    a = torch.linspace(0.1, 0.9, 7, dtype=torch.float64)
    b = torch.linspace(-0.4, 0.4, 5, dtype=torch.float64)
    return rays, a, b


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and parallel.is_distributed()
    rays, a, b = _make_problem()
    lo, hi = parallel.shard_range(rays.shape[0], rank, world)
    params = [a.clone().requires_grad_(True), b.clone().requires_grad_(True)]
    opt = optimizer.SGD_Optimizer(_StubEngine(rays[lo:hi]), params, _error_function(params), 3,
                                  learning_rate=0.5, grad_clip=1e9)
    grads, mean_err = opt.process_gradient([None, None])
    # bucket all-reduce keeps shapes / None entries; max and sum over ranks
    t = [torch.full((3,), float(rank + 1)), None, torch.full((2, 2), 10.0 * (rank + 1))]
    parallel.allreduce_sum_(t)
    assert t[1] is None and torch.all(t[0] == 3.0) and torch.all(t[2] == 30.0)
    assert parallel.max_over_ranks(float(rank)) == world - 1
    assert parallel.sum_over_ranks(1.5) == 1.5 * world
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), g0=grads[0].numpy(), g1=grads[1].numpy(),
             err=float(mean_err))
    dist.barrier()
    dist.destroy_process_group()


def test_process_gradient_allreduce_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-process answer over all rays
    rays, a, b = _make_problem()
    params = [a.clone().requires_grad_(True), b.clone().requires_grad_(True)]
    opt = optimizer.SGD_Optimizer(_StubEngine(rays), params, _error_function(params), 3,
                                  learning_rate=0.5, grad_clip=1e9)
    want, want_err = opt.process_gradient([None, None])
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    for r in range(2):   # every rank ends with the single-process gradient and mean error
        assert np.allclose(got[r]["g0"], want[0].numpy(), rtol=1e-6)
        assert np.allclose(got[r]["g1"], want[1].numpy(), rtol=1e-6)
        assert np.allclose(got[r]["err"], float(want_err), rtol=1e-6)
