"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: helper API of gsplat/distributed.py
(reference tests/_test_distributed.py:13-114 checks the same helpers with full(rank) tensors) plus the
attribute all-gather / gradient reduce-scatter the view-parallel render path is built on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gscodec_studio_b200.distributed import (all_gather_attributes, all_gather_int32, all_gather_tensor_list,
                                                  all_to_all_int32, all_to_all_tensor_list)
    try:
        dev = torch.device("cpu")
        assert all_gather_int32(world, rank + 5, device=dev) == [5, 6]
        got = all_gather_int32(world, torch.tensor(rank, dtype=torch.int), device=dev)
        assert [int(t) for t in got] == [0, 1]
        assert all_to_all_int32(world, [rank * 10 + 0, rank * 10 + 1], device=dev) == [rank, 10 + rank]
        # equal-size gather of a tensor list, differentiable
        a = torch.full((3, 2), float(rank), requires_grad=True)
        b = torch.full((3,), float(rank) + 0.5, requires_grad=True)
        ga, gb = all_gather_tensor_list(world, [a, b])
        assert ga.shape == (6, 2) and gb.shape == (6,)
        assert torch.equal(ga[:3], torch.zeros(3, 2)) and torch.equal(gb[3:], torch.full((3,), 1.5))
        ((ga * (rank + 1)).sum() + (gb * 2).sum()).backward()
        # every rank contributed d/d(ga) = (its rank + 1): the owner receives the SUM over ranks = 3
        assert torch.equal(a.grad, torch.full((3, 2), 3.0)) and torch.equal(b.grad, torch.full((3,), 4.0))
        # ragged attribute gather (different number of Gaussians per rank) + reduce-scatter of grads
        n = 2 + rank
        means = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) + 100 * rank
        means.requires_grad_(True)
        opac = torch.full((n,), float(rank), requires_grad=True)
        gm, go = all_gather_attributes([means, opac])
        assert gm.shape == (5, 3) and go.shape == (5,)
        assert torch.equal(go, torch.tensor([0.0, 0.0, 1.0, 1.0, 1.0]))
        w = torch.arange(5, dtype=torch.float32)
        ((gm * w[:, None]).sum() * (rank + 1) + (go * w).sum()).backward()
        rows = w[:2] if rank == 0 else w[2:]
        assert torch.equal(means.grad, (rows * 3)[:, None].expand(-1, 3))
        assert torch.equal(opac.grad, rows * 2)
        # padded layout used by rasterization(distributed=True): world * max(n_r) rows, inert padding rows
        from gscodec_studio_b200.distributed import gather_counts, start_gather
        geo = (torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 100 * rank).requires_grad_(True)
        pad = torch.tensor([float("nan"), 1.0, 0.0, 0.0])
        gg = start_gather(geo, gather_counts(n, dev), keep_padding=True, pad_row=pad).finish()
        assert gg.shape == (6, 4)
        assert torch.equal(gg[3:], torch.arange(12, dtype=torch.float32).reshape(3, 4) + 100)
        assert torch.isnan(gg[2, 0]) and torch.equal(gg[2, 1:], pad[1:]) and torch.equal(gg[:2], geo.detach() if rank == 0 else gg[:2])
        w6 = torch.arange(6, dtype=torch.float32)
        (gg.nan_to_num(0.0) * w6[:, None]).sum().backward()
        rows = w6[:2] if rank == 0 else w6[3:]
        assert torch.equal(geo.grad, (rows * 2)[:, None].expand(-1, 4))
        # per-call count exchange (shared-memory board on one node): ranks drift in and out of step
        import random
        import time
        from gscodec_studio_b200 import distributed as D
        rnd = random.Random(rank)
        for i in range(3000):
            if rnd.random() < 0.01:
                time.sleep(rnd.random() * 1e-3)
            assert gather_counts(i * 3 + rank, dev) == [i * 3, i * 3 + 1]
        assert D._host_board is not None and D._host_board.ok
        # all-to-all with variable splits, differentiable
        t = torch.arange(4, dtype=torch.float32)[:, None].repeat(1, 2) + 10 * rank
        t.requires_grad_(True)
        splits = [1, 3] if rank == 0 else [2, 2]
        (r,) = all_to_all_tensor_list(world, [t], splits)
        exp = torch.tensor([[0.0], [10.0], [11.0]]) if rank == 0 else torch.tensor([[1.0], [2.0], [3.0], [12.0], [13.0]])
        assert torch.equal(r[:, :1], exp)
        r.sum().backward()
        assert torch.equal(t.grad, torch.ones(4, 2))
        out[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback

        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_distributed_helpers_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}, dict(out)


def test_single_process_passthrough():
    from gscodec_studio_b200.distributed import all_gather_attributes, all_gather_int32, all_gather_tensor_list

    t = torch.ones(3, 2)
    assert all_gather_int32(1, 7) == [7]
    assert all_gather_tensor_list(1, [t])[0] is t
    assert all_gather_attributes([t])[0] is t
