"""GPU micro-benchmark of the network forward alone (CUDA events).
Usage: python tools/net_bench.py [leaves] [iters] [impl] [connect4|caro]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

from caro_ai_b200.game import ConnectFour, TicTacToe
from caro_ai_b200.model import DeviceNet, Net

FLOP_PER_LEAF = 15598672


def main():
    leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 10368
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    impl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    caro = len(sys.argv) > 4 and sys.argv[4] == "caro"
    game = TicTacToe(15, 5) if caro else ConnectFour()
    flop = 83760340 if caro else FLOP_PER_LEAF
    torch.manual_seed(0)
    net = Net(game.obs_shape, game.action_space).eval()
    dn = DeviceNet(net, game, precision="bf16")
    rng = np.random.default_rng(0)
    # random legal-looking boards: random heights, random colours
    boards = np.zeros((leaves, 8 if caro else 2), dtype=np.uint64)
    for i in range(0 if caro else leaves):
        mask = black = 0
        for c in range(7):
            h = int(rng.integers(0, 7))
            col = (1 << h) - 1
            mask |= col << (7 * c)
            black |= (int(rng.integers(0, 64)) & col) << (7 * c)
        boards[i] = (mask, black)
    d_boards = torch.from_numpy(boards.view(np.int64)).cuda()
    d_who = torch.from_numpy(rng.integers(0, 2, leaves).astype(np.uint8)).cuda()
    for _ in range(3):
        dn.forward_boards(d_boards, d_who, leaves, impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dn.forward_boards(d_boards, d_who, leaves, impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("leaves=%d impl=%d ms=%.4f leaves/s=%.3e TFLOP/s=%.1f" % (leaves, impl, ms, leaves / ms * 1e3, leaves * flop / ms / 1e9))


if __name__ == "__main__":
    main()
