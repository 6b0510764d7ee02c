import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import TicTacToe
from caro_ai_b200.model import DeviceNet, Net
game = TicTacToe(15, 5)
torch.manual_seed(0)
dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = SelfPlayEngine(game, G, max_batch=8, node_capacity=16384, seed=1)
eng.play(dn, dn, moves=3, count=200, batch=8, tau_plies=10, auto_restart=True)
torch.cuda.synchronize()
eng.profile(2)
n = 50
eng.search(dn, n, 8)
p = eng.profile_read()
print({k: (round(v / n, 4) if k.endswith("_ms") else v) for k, v in p.items()}, eng.counters())
