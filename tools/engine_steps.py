"""GPU helper for profiling the tree kernels: Connect4, G games, a short pre-roll, then `n` search minibatches.
Usage: python tools/engine_steps.py [games] [preroll_plies] [minibatches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    preroll = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    eng = SelfPlayEngine(game, games, max_batch=8, node_capacity=24576, seed=7)
    eng.play(dn, dn, moves=preroll, count=100, batch=8, tau_plies=10, auto_restart=True)
    torch.cuda.synchronize()
    eng.profile(2)
    eng.search(dn, n, 8)
    p = eng.profile_read()
    print({k: (v / n if k.endswith("_ms") else v) for k, v in p.items()}, eng.counters())


if __name__ == "__main__":
    main()
