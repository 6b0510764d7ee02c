"""GPU debug: how do consecutive tower launches of the self-play pipeline overlap?  The row-tiled tower stamps the
global timer at entry / exit of its first and last CTA (trace[7998] == 2); this tool runs a few plies of the 2-part
pipeline and prints, per launch, when CTA 0 and CTA 147 started and ended relative to the previous launch.
Usage: python tools/pipeline_trace.py [games_per_part] [parts]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200 import _cabi
from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    gpp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    engs = [SelfPlayEngine(game, gpp, max_batch=8, node_capacity=24576, seed=5 + h) for h in range(parts)]

    def play(n, count=100):
        SelfPlayEngine.play_multi(engs, dn, moves=n, count=count, batch=8, tau_plies=10, auto_restart=True)

    play(12, 8)
    trace = torch.zeros(8000 + 2 * 4096, dtype=torch.int64, device="cuda")
    trace[7998] = 2
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, trace.data_ptr()))  # before the ply graph is captured: the pointer is baked in
    play(3)
    torch.cuda.synchronize()
    trace[8000] = 0
    trace[8000 + 4096] = 0
    torch.cuda.synchronize()
    play(2)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, None))
    t = trace.cpu().numpy()
    a, b = t[8000:8000 + 4096], t[8000 + 4096:]
    n = int(min(a[0], b[0], 2040))
    rows = []
    for k in range(n):
        rows.append((a[1 + 2 * k], a[2 + 2 * k], b[1 + 2 * k], b[2 + 2 * k]))
    t0 = rows[0][0]
    print("launches recorded: first CTA %d, last CTA %d" % (a[0], b[0]))
    print("%5s %10s %10s %10s %10s %12s" % ("k", "cta0_in", "cta0_out", "ctaN_in", "ctaN_out", "period_us"))
    prev = None
    for k, (s0, e0, s1, e1) in enumerate(rows[100:140], start=100):
        period = (s0 - prev) / 1e3 if prev is not None else 0.0
        prev = s0
        print("%5d %10.1f %10.1f %10.1f %10.1f %12.1f" % (k, (s0 - t0) / 1e3, (e0 - t0) / 1e3, (s1 - t0) / 1e3, (e1 - t0) / 1e3, period))


if __name__ == "__main__":
    main()
