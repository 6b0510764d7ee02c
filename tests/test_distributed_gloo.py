"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: gradient bucket all-reduce, weight broadcast,
W/L/D reduction, game sharding.  The data path itself has no collective (games shard by rank)."""
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


def _worker(rank, world_size, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from caro_ai_b200 import distributed as D
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import Net
    g = ConnectFour()
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    net = Net(g.obs_shape, g.action_space)
    D.broadcast_state_dict(net, src=0)
    torch.manual_seed(100)
    ref = Net(g.obs_shape, g.action_space)
    same = all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), ref.state_dict().values()))
    # gradients: rank r holds grad = r + 1 everywhere -> mean 1.5
    for p in net.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    n = D.allreduce_gradients(net.parameters())
    ok_grad = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in net.parameters())
    # persistent bucket: .grad tensors are views, a real backward accumulates into them, one all-reduce averages
    torch.manual_seed(7)
    net2 = Net(g.obs_shape, g.action_space)
    bucket = D.FlatGradients(net2.parameters())
    bucket.zero()
    x = torch.full((4, 2, 6, 7), float(rank + 1))
    logits, val = net2(x)
    (logits.sum() + val.sum()).backward()
    views_kept = all(p.grad.data_ptr() >= bucket.flat.data_ptr() and
                     p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.flat.numel() * 4 for p in net2.parameters())
    local = bucket.flat.clone()
    n2 = bucket.allreduce()
    gathered = [torch.zeros_like(local) for _ in range(world_size)]
    dist.all_gather(gathered, local)
    ok_bucket = views_kept and n2 == n and torch.allclose(bucket.flat, sum(gathered) / world_size, rtol=1e-6, atol=1e-6) \
        and bool(local.abs().sum() > 0)
    ok_grad = ok_grad and ok_bucket
    tall = D.reduce_tallies(rank + 1, 10 * (rank + 1), 0)
    first, count = D.shard_games(4097, rank, world_size)
    # train.py:199 taken collectively: rank 0 holds 2,500 replay entries, rank 1 only 1,500 -> BOTH skip the SGD rounds
    # (a per-rank decision would leave rank 0 alone inside its gradient all-reduces)
    decision = D.all_min(2500 if rank == 0 else 1500) >= 2000
    # the SGD batch assembled from the ranks' replay samples: rank order, every rank gets the same 256 rows
    rows = 128
    mine = (torch.full((rows, 2, 6, 7), float(rank)), torch.full((rows, 7), float(rank) + 0.5), torch.full((rows,), float(-rank)))
    planes, pi, z = D.all_gather_rows(mine)
    ok_gather = planes.shape == (256, 2, 6, 7) and pi.shape == (256, 7) and z.shape == (256,) \
        and bool((planes[:rows] == 0).all()) and bool((planes[rows:] == 1).all()) \
        and bool((pi[:rows] == 0.5).all()) and bool((pi[rows:] == 1.5).all()) and bool((z[rows:] == -1).all())
    results[rank] = (same, ok_grad, n, tall, first, count, decision, ok_gather)
    dist.destroy_process_group()


def test_two_rank_collectives():
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
        r0, r1 = results[0], results[1]
    assert r0[0] and r1[0], "weights differ after broadcast_state_dict"
    assert r0[1] and r1[1], "gradient all-reduce did not average"
    assert r0[2] == r1[2] == 188301  # trainable parameters of the Connect4 network
    assert r0[3] == r1[3] == (3, 30, 0)
    assert (r0[4], r0[5], r1[4], r1[5]) == (0, 2049, 2049, 2048)
    assert r0[6] is False and r1[6] is False, "train-or-skip must be decided on the smallest ring"
    assert r0[7] and r1[7], "all_gather_rows did not assemble the batch in rank order"
