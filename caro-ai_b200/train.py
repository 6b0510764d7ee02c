#!/usr/bin/env python3
"""AlphaZero training loop with the reference's CLI and behaviour (train.py:1-217), self-play and arena
evaluation running as lock-step batches on the GPU engine.

    python -m caro_ai_b200.train -g 0 -n run_name [--games 4096] [--max-steps N]
    torchrun --nproc-per-node 8 -m caro_ai_b200.train -g 0 -n run_name        # games shard by rank

Same pieces as the reference: replay deque (maxlen REPLAY_BUFFER x ranks' local share), 10 SGD rounds of 256
samples (MSE + soft-target cross-entropy, lr 0.1, momentum 0.9), arena of EVALUATION_ROUNDS games every
EVALUATE_EVERY_STEP steps with 20 x 16 searches at tau = 0, promotion above 0.60, checkpoints
``saves/<name>/best_%03d_%05d.dat`` written with ``torch.save(net.state_dict())``, TensorBoard tags
speed_steps / speed_nodes / loss_total / loss_value / loss_policy / eval_win_ratio.
The SGD step itself is plain PyTorch autograd (out of scope for hand kernels); with several ranks the gradients
are averaged by one flattened NCCL all-reduce per step and promoted weights are broadcast from rank 0.
"""
from __future__ import annotations

import argparse
import collections
import os
import random
import sys
import time

import torch
import torch.nn.functional as F
import torch.optim as optim

from . import config as cfg
from . import distributed as D
from .game import get_game
from .model import DeviceNet, Net, NetWrapper, save_checkpoint
from .utils import TBMeanTracker, play_games_batched


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass


def make_writer(name: str):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(comment="-" + name)
    except Exception:  # noqa: BLE001
        return _NullWriter()


def self_play(game, replay_buffer, best: DeviceNet, games: int, tb, step_idx: int, seed: int):
    """train.py:25-59 for `games` games at once; returns (plies, new nodes) like the reference's counters."""
    t = time.time()
    stats = play_games_batched(game, games, best, best, cfg.STEPS_BEFORE_TAU_0, cfg.MCTS_SEARCHES, cfg.MCTS_BATCH_SIZE,
                               replay_buffer=replay_buffer, trees_per_game=1, seed=seed)
    dt = time.time() - t
    tb.track("speed_steps", stats["plies"] / dt, step_idx)
    tb.track("speed_nodes", stats["leaf_evals"] / dt, step_idx)
    return stats, dt


def train_neural_net(game, net: Net, replay_buffer, optimizer, tb, step_idx: int, device, bucket=None):
    """train.py:62-117.  ``bucket``: a ``distributed.FlatGradients`` over the net's parameters (one all-reduce per
    round, no flatten / scatter); without it the gradients are flattened per round (``allreduce_gradients``)."""
    sums = [0.0, 0.0, 0.0]
    net.train()
    for _ in range(cfg.TRAIN_ROUNDS):
        batch = random.sample(replay_buffer, cfg.BATCH_SIZE)
        states, who, probs, values = zip(*batch)
        d_boards = torch.from_numpy(game.boards_from_states(states).view("int64")).to(device)
        d_who = torch.tensor(list(who), dtype=torch.uint8, device=device)
        states_t = game.planes_device(d_boards, d_who, len(states))
        if bucket is not None:
            bucket.zero()
        else:
            optimizer.zero_grad()
        probs_v = torch.tensor(probs, dtype=torch.float32, device=device)
        values_v = torch.tensor(values, dtype=torch.float32, device=device)
        out_logits, out_values = net(states_t)
        loss_value = F.mse_loss(out_values.squeeze(-1), values_v)
        loss_policy = (-F.log_softmax(out_logits, dim=1) * probs_v).sum(dim=1).mean()
        loss = loss_policy + loss_value
        loss.backward()
        if bucket is not None:
            bucket.allreduce()
        else:
            D.allreduce_gradients(net.parameters())
        optimizer.step()
        sums[0] += loss.item()
        sums[1] += loss_value.item()
        sums[2] += loss_policy.item()
    for tag, s in zip(("loss_total", "loss_value", "loss_policy"), sums):
        tb.track(tag, s / cfg.TRAIN_ROUNDS, step_idx)
    return [s / cfg.TRAIN_ROUNDS for s in sums]


def evaluate(game, challenger: Net, champion: DeviceNet, rounds: int, seed: int, device) -> float:
    """train.py:120-149: challenger vs champion, 20 x 16 searches, tau = 0; rounds shard over ranks."""
    rank, ws = D.world()
    _, mine = D.shard_games(rounds, rank, ws)
    challenger.eval()
    ch = DeviceNet(challenger, game)
    w = l = d = 0
    if mine:
        s = play_games_batched(game, mine, ch, champion, 0, 20, 16, trees_per_game=2, seed=seed + rank)
        w, l, d = s["wins"], s["losses"], s["draws"]
    ch.close()
    w, l, d = D.reduce_tallies(w, l, d, device=device)
    return w / max(1, w + l + d)


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("-n", "--name", required=True, help="Name of the run")
    p.add_argument("--cuda", default=False, action="store_true", help="accepted for compatibility (always CUDA)")
    p.add_argument("-g", "--game", required=True, choices=["0", "1"], help="0: Connect4, 1: TicTacToe")
    p.add_argument("--games", type=int, default=256, help="self-play games per step and per rank (reference: PLAY_EPISODES=1)")
    p.add_argument("--max-steps", type=int, default=0, help="stop after this many steps (0 = run forever, like the reference)")
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank, ws = D.world()
    device = torch.device("cuda", torch.cuda.current_device())
    saves_path = os.path.join("saves", args.name)
    if rank == 0:
        os.makedirs(saves_path, exist_ok=True)
    writer = make_writer(args.name) if rank == 0 else _NullWriter()
    game = get_game(args.game)
    net = Net(game.obs_shape, game.action_space).to(device)
    D.broadcast_state_dict(net)
    best_net = NetWrapper(net)
    best_dev = DeviceNet(best_net.target_model, game)
    if rank == 0:
        print(net)
    optimizer = optim.SGD(net.parameters(), lr=cfg.LEARNING_RATE, momentum=0.9)
    bucket = D.FlatGradients(net.parameters())
    replay_buffer = collections.deque(maxlen=cfg.REPLAY_BUFFER)
    step_idx = best_idx = 0
    with TBMeanTracker(writer, batch_size=10) as tb:
        while args.max_steps == 0 or step_idx < args.max_steps:
            stats, dt = self_play(game, replay_buffer, best_dev, args.games, tb, step_idx, seed=1000 * step_idx + rank)
            step_idx += 1
            if rank == 0:
                print("Step %d, steps %3d, leaves %4d, steps/s %5.2f, leaves/s %6.2f, best_idx %d, replay %d" % (
                    step_idx, stats["plies"], stats["leaf_evals"], stats["plies"] / dt, stats["leaf_evals"] / dt, best_idx,
                    len(replay_buffer)), end="\r")
                sys.stdout.flush()
            if len(replay_buffer) < cfg.MIN_REPLAY_TO_TRAIN:
                continue
            train_neural_net(game, net, replay_buffer, optimizer, tb, step_idx, device, bucket)
            if step_idx % cfg.EVALUATE_EVERY_STEP == 0:
                win_ratio = evaluate(game, net, best_dev, cfg.EVALUATION_ROUNDS, seed=step_idx, device=device)
                if rank == 0:
                    print("Net evaluated, win ratio = %.2f" % win_ratio)
                writer.add_scalar("eval_win_ratio", win_ratio, step_idx)
                if win_ratio > cfg.BEST_NET_WIN_RATIO:
                    if rank == 0:
                        print("Net is better than cur best, sync")
                    best_net.sync()
                    best_dev.update(best_net.target_model)
                    best_idx += 1
                    if rank == 0:
                        save_checkpoint(net, os.path.join(saves_path, "best_%03d_%05d.dat" % (best_idx, step_idx)))
    if ws > 1 and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
