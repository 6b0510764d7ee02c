"""GPU helper: a digest of the trees a virtual-loss search grows (used to compare the one-thread-per-game and the
eight-lanes-per-game select kernels, CARO_VL_GROUP=0 / 1, which must make identical choices).  Prints one line."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def digest(games=96, moves=7, count=12):
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
    eng = SelfPlayEngine(game, games, max_batch=8, node_capacity=2048, seed=11, virtual_loss=True, mask_priors=True)
    eng.play(dn, dn, moves=moves, count=count, batch=8, tau_plies=10, auto_restart=True)
    c = eng.counters()
    n = eng.pool("N").to(torch.int64)
    w = eng.pool("W").double()
    idx = torch.arange(n.numel(), device=n.device, dtype=torch.int64).view_as(n) % 1000003
    out = (c["leaf_evals"], c["descents"], c["errors"], int(eng.region("node_count").sum().item()), int((n * idx).sum().item()),
           float(w.abs().sum().item()), hash(tuple(eng.roots()[0])) & 0xFFFFFFFF)
    eng.close()
    dn.close()
    return out


if __name__ == "__main__":
    print("DIGEST", digest())
