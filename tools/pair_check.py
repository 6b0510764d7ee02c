"""A/B of the CTA-pair tower (CARO_RT_PAIR=1) against the single-CTA tower: outputs for several leaf counts saved to a file
(the env var is read once per process, so run it twice and compare), and timing.
Usage: CARO_RT_PAIR=0|1 python tools/pair_check.py save FILE   |   python tools/pair_check.py compare FILE0 FILE1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

COUNTS = [1, 15, 16, 17, 33, 100, 1000, 4737, 9472, 19264, 33152]


def boards_for(game_name, leaves, rng):
    from caro_ai_b200.game import ConnectFour, TicTacToe
    if game_name == "c4":
        boards = np.zeros((leaves, 2), dtype=np.uint64)
        for i in range(leaves):
            mask = black = 0
            for c in range(7):
                h = int(rng.integers(0, 7))
                col = (1 << h) - 1
                mask |= col << (7 * c)
                black |= (int(rng.integers(0, 64)) & col) << (7 * c)
            boards[i] = (mask, black)
        return ConnectFour(), boards
    n = 3 if game_name == "ttt" else 6
    boards = np.zeros((leaves, 8), dtype=np.uint64)
    for i in range(leaves):
        w = b = 0
        for cell in range(n * n):
            r = int(rng.integers(0, 3))
            if r == 1:
                w |= 1 << cell
            elif r == 2:
                b |= 1 << cell
        boards[i, 0] = w
        boards[i, 4] = b
    return TicTacToe(n, 3), boards


def save(path):
    from caro_ai_b200.model import DeviceNet, Net
    out = {}
    for game_name in ("c4", "ttt", "mnk6"):
        rng = np.random.default_rng(1)
        game, boards = boards_for(game_name, max(COUNTS) if game_name == "c4" else 1000, rng)
        torch.manual_seed(0)
        dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
        d_boards = torch.from_numpy(boards.view(np.int64)).cuda()
        d_who = torch.from_numpy(rng.integers(0, 2, len(boards)).astype(np.uint8)).cuda()
        for n in COUNTS:
            if n > len(boards):
                continue
            p, v = dn.forward_boards(d_boards, d_who, n, 0)
            torch.cuda.synchronize()
            out[(game_name, n)] = (p[:n].cpu().clone(), v[:n].cpu().clone())
        if game_name == "c4":
            for n in (9472, 19264, 33152):
                for _ in range(3):
                    dn.forward_boards(d_boards, d_who, n, 0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    dn.forward_boards(d_boards, d_who, n, 0)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 50
                print("pair=%s leaves=%d ms=%.4f TFLOP/s=%.1f" % (os.environ.get("CARO_RT_PAIR", "0"), n, ms, n * 15598672 / ms / 1e9), flush=True)
    torch.save(out, path)


def compare(f0, f1):
    a, b = torch.load(f0), torch.load(f1)
    bad = 0
    for k in a:
        dp = (a[k][0] - b[k][0]).abs().max().item()
        dv = (a[k][1] - b[k][1]).abs().max().item()
        same = torch.equal(a[k][0], b[k][0]) and torch.equal(a[k][1], b[k][1])
        print(k, "bit-identical" if same else "DIFF max|dp|=%.3e max|dv|=%.3e" % (dp, dv))
        bad += not same
    print("MISMATCHES:", bad)


if __name__ == "__main__":
    if sys.argv[1] == "save":
        save(sys.argv[2])
    else:
        compare(sys.argv[2], sys.argv[3])
