"""CPU: the N>1 host logic (view sharding + one all-reduce of the contiguous gradient arena) with two gloo ranks.
The rasteriser itself needs a GPU, so the inner step is a stand-in that writes a deterministic per-view gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gstex_cuda_b200.pipeline import DataParallelTrainStep


class _FakeInner:
    """Mimics FusedTrainStep.step: accumulates a known gradient per view into one flat arena whose LAST slot is the loss
    accumulator (the all-reduce sums gradients and loss in one collective), and zero-fills when given no views."""

    def __init__(self, size):
        self.grad_arena = torch.zeros(size + 1)
        self.loss = self.grad_arena[-1:]

    def step(self, cameras, targets):
        self.grad_arena.zero_()
        for (vm, _), tgt in zip(cameras, targets):
            v = float(vm[0, 0])
            self.grad_arena[:-1] += v * torch.arange(1, self.grad_arena.numel(), dtype=torch.float32)
            self.loss += v + float(tgt.sum())
        return self.loss


def _worker(rank, world, port, nviews, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cams = [(torch.full((4, 4), float(v + 1)), torch.eye(4)) for v in range(nviews)]
    targets = [torch.full((2, 2, 3), 0.5 * v) for v in range(nviews)]
    dp = DataParallelTrainStep(_FakeInner(17), rank, world)
    loss = dp.step(cams, targets)
    torch.save({"arena": dp.inner.grad_arena.clone(), "loss": loss.clone(),
                "mine": DataParallelTrainStep.shard(nviews, rank, world)}, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


import pytest


@pytest.mark.parametrize("nviews", [7, 1])  # 1 view on 2 ranks: rank 1's shard is empty and must contribute zeros
def test_two_rank_gloo_allreduce_equals_single_rank(tmp_path, nviews):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nviews, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    # shards partition the views
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(nviews))
    # every rank ends with the same arena, equal to a single-rank run over all views
    cams = [(torch.full((4, 4), float(v + 1)), torch.eye(4)) for v in range(nviews)]
    targets = [torch.full((2, 2, 3), 0.5 * v) for v in range(nviews)]
    single = DataParallelTrainStep(_FakeInner(17), 0, 1)
    loss1 = single.step(cams, targets)
    for r in res:
        torch.testing.assert_close(r["arena"], single.inner.grad_arena)
        torch.testing.assert_close(r["loss"], loss1)


def test_wrapper_rejects_caller_owned_gradient_views():
    """Gradients placed outside inner.grad_arena would be left out of the all-reduce."""
    inner = _FakeInner(4)
    inner.has_grad_views = True
    with pytest.raises(RuntimeError, match="grad_views"):
        DataParallelTrainStep(inner, 0, 2)


def test_shard_partitions():
    for nv, ws in ((64, 8), (7, 4), (3, 8), (64, 1), (0, 2)):
        parts = [DataParallelTrainStep.shard(nv, r, ws) for r in range(ws)]
        assert sorted(sum(parts, [])) == list(range(nv))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
