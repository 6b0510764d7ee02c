"""GPU debug: pipeline timeline of CTA 0 of the tcgen05 tower (clock64 stamps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

from caro_ai_b200 import _cabi
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 10368
    impl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    div = 8 if impl == 0 else 4  # the row-tiled kernel stamps gl * 8 + tile, the tap-per-MMA kernel gl * 4 + tile
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    boards = torch.zeros((leaves, 2), dtype=torch.int64, device="cuda")
    who = torch.zeros(leaves, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dn.forward_boards(boards, who, leaves, impl)
    trace = torch.zeros(8000, dtype=torch.int64, device="cuda")
    if len(sys.argv) > 4 and sys.argv[4] == "mma-only":
        trace[7999] = 1  # row-tiled kernel: epilogue warps only pass the barriers on
    trace[7996] = int(os.environ.get("TRACE_FLAGS", "0"))  # CTA-pair tower experiments (bit 0: accumulators are not zeroed)
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, trace.data_ptr()))
    dn.forward_boards(boards, who, leaves, impl)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, None))
    t = trace.cpu().numpy()
    names = ["mma_start", "mma_issued", "epi_start", "epi_done", "lastepi+inputs", "heads_done", "weights_ok"]
    ev = []
    for kind in (0, 1, 2, 3, 4, 6, 7):
        for idx in range(1000):
            v = int(t[kind * 1000 + idx])
            if v and not (kind == 7 and idx >= 4) and not (kind * 1000 + idx >= 7990):  # trace[7998], trace[7999] are mode flags, not stamps
                ev.append((v, kind, idx))
    ev.sort()
    t0 = ev[0][0]
    for gl in range(1, 8):  # first tile of a layer: before / after the waits for the two weight regions (local, peer)
        w = [int(x) - t0 for x in t[5000 + gl * 6:5006 + gl * 6]]
        if w[0] > 0:
            print("weights gl=%d: region0 at %d local +%d peer +%d | region1 at %d local +%d peer +%d" %
                  (gl, w[0], w[1] - w[0], w[2] - w[1], w[3], w[4] - w[3], w[5] - w[4]))
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 140
    for clk, kind, idx in ev[:limit]:
        if kind == 6:
            print("%8d  %-14s gl=%d %s" % (clk - t0, "layer_enter" if idx % 2 == 0 else "bars_passed", idx // 2, ""))
        elif kind == 7:
            print("%8d  %s" % (clk - t0, ["kernel_entered", "setup_done", "warp0_finished", "all_finished"][idx]))
        elif kind < 4:
            print("%8d  %-14s gl=%d t=%d" % (clk - t0, names[kind], idx // div, idx % div))
        else:
            print("%8d  %-14s group=%d" % (clk - t0, names[kind], idx))


if __name__ == "__main__":
    main()
