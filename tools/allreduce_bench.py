"""GPU measurement (torchrun, N ranks): the training loop's collectives -- one flattened NCCL all-reduce of the
Connect4 network's gradients per SGD round (train.py:82-111 + SURVEY.md section 8e) and the weight broadcast after a
promotion.  Device-timed with CUDA events, max over ranks.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.distributed as dist

from caro_ai_b200 import distributed as D
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import Net


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = D.world()
    game = ConnectFour()
    torch.manual_seed(rank)
    net = Net(game.obs_shape, game.action_space).cuda()
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    n = sum(p.numel() for p in net.parameters())

    def timed(fn, iters=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ar = timed(lambda: D.allreduce_gradients(net.parameters()))
    net2 = Net(game.obs_shape, game.action_space).cuda()
    bucket = D.FlatGradients(net2.parameters())
    bucket.flat.normal_()
    ar_flat = timed(bucket.allreduce)
    bc = timed(lambda: D.broadcast_state_dict(net))
    # the gradients really are averaged: every rank holds the same values afterwards
    D.allreduce_gradients(net.parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(flat, ref))
    if rank == 0:
        print(json.dumps({"ranks": ws, "gradient_elements": n, "gradient_bytes": 4 * n, "allreduce_us_flatten_per_step": ar, "allreduce_us_persistent_bucket": ar_flat,
                          "weight_broadcast_us": bc, "ranks_agree_bitwise": same}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
