"""CTA-pair form of the tap-per-MMA tower (impl 6) against the single-CTA form (impl 3 / 0 for large boards): bit identity on
several geometries and leaf counts, and timing on Caro 15x15.  Usage: python tools/tc_pair_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

from caro_ai_b200.game import ConnectFour, TicTacToe
from caro_ai_b200.model import DeviceNet, Net


def boards_for(game, leaves, rng):
    n = game.obs_shape[1]
    if isinstance(game, ConnectFour):
        boards = np.zeros((leaves, 2), dtype=np.uint64)
        for i in range(leaves):
            mask = black = 0
            for c in range(7):
                h = int(rng.integers(0, 7))
                col = (1 << h) - 1
                mask |= col << (7 * c)
                black |= (int(rng.integers(0, 64)) & col) << (7 * c)
            boards[i] = (mask, black)
        return boards
    boards = np.zeros((leaves, 8), dtype=np.uint64)
    for i in range(leaves):
        cells = rng.integers(0, 3, n * n)
        for cell, r in enumerate(cells):
            if r == 1:
                boards[i, cell // 64] |= np.uint64(1 << (cell % 64))
            elif r == 2:
                boards[i, 4 + cell // 64] |= np.uint64(1 << (cell % 64))
    return boards


def main():
    rng = np.random.default_rng(2)
    bad = 0
    for game, counts in ((TicTacToe(15, 5), (1, 2, 3, 5, 300, 1501, 4096)), (TicTacToe(9, 5), (7, 100, 2000)), (ConnectFour(), (17, 1000))):
        torch.manual_seed(0)
        dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
        boards = boards_for(game, max(counts), rng)
        d_boards = torch.from_numpy(boards.view(np.int64)).cuda()
        d_who = torch.from_numpy(rng.integers(0, 2, len(boards)).astype(np.uint8)).cuda()
        for n in counts:
            p3, v3 = dn.forward_boards(d_boards, d_who, n, 3)
            p6, v6 = dn.forward_boards(d_boards, d_who, n, 6)
            torch.cuda.synchronize()
            same = torch.equal(p3, p6) and torch.equal(v3, v6)
            print(game.obs_shape, n, "bit-identical" if same else "DIFF %.3e %.3e" % ((p3 - p6).abs().max().item(), (v3 - v6).abs().max().item()), flush=True)
            bad += not same
        if game.obs_shape[1] == 15:
            for n in (4096, 7963, 16384):
                if n > len(boards):
                    bb = torch.cat([d_boards] * (n // len(boards) + 1))[:n].contiguous()
                    ww = torch.cat([d_who] * (n // len(boards) + 1))[:n].contiguous()
                else:
                    bb, ww = d_boards, d_who
                for impl in (3, 6):
                    for _ in range(3):
                        dn.forward_boards(bb, ww, n, impl)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        dn.forward_boards(bb, ww, n, impl)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 20
                    print("caro leaves=%d impl=%d ms=%.4f leaves/s=%.3e TFLOP/s=%.1f" % (n, impl, ms, n / ms * 1e3, n * 83760340 / ms / 1e9), flush=True)
        dn.close()
    print("MISMATCHES:", bad)


if __name__ == "__main__":
    main()
