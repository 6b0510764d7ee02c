"""GPU debug of the split-precision row-tiled tower (net_rx.cu): time of a launch with the whole kernel, with the MMA stream
only (epilogue warps pass the barriers on) and with the epilogue chain only (no MMA issued), plus CTA 0's timeline of
one layer (weight-block starts, epilogue start / done per tile).  Usage: python tools/rx_probe.py [leaves]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200 import _cabi
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16x3")
    boards = torch.zeros((leaves, 2), dtype=torch.int64, device="cuda")
    who = torch.zeros(leaves, dtype=torch.uint8, device="cuda")
    trace = torch.zeros(8000, dtype=torch.int64, device="cuda")

    def timed(mode, iters=50):
        trace.zero_()
        trace[7999] = mode
        _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, trace.data_ptr() if mode >= 0 else None))
        for _ in range(3):
            dn.forward_boards(boards, who, leaves, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dn.forward_boards(boards, who, leaves, 2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for name, mode in (("full kernel (no trace)", -1), ("full kernel (traced)", 0), ("MMA stream only", 1), ("epilogue chain only", 2)):
        print("%-26s %.4f ms per %d leaves" % (name, timed(mode), leaves), flush=True)
    # timeline of CTA 0, layers 1 and 2 of its first group
    trace.zero_()
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, trace.data_ptr()))
    dn.forward_boards(boards, who, leaves, 2)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib().caro_net_set_trace(dn.handle, None))
    t = trace.cpu().numpy()
    ev = []
    for idx in range(1000):
        if t[idx]:
            gl, r = divmod(idx, 32)
            ev.append((int(t[idx]), "blk   gl=%d b=%2d %s" % (gl, r // 2, "lo" if r & 1 else "hi")))
        for kind, nm in ((2, "epi_start"), (3, "epi_done ")):
            if t[kind * 1000 + idx]:
                ev.append((int(t[kind * 1000 + idx]), "%s gl=%d y=%d" % (nm, idx // 8, idx % 8)))
    ev.sort()
    t0 = ev[0][0]
    for clk, what in ev:
        if " gl=1 " in what or " gl=2 " in what or " gl=0 " in what:
            print("%8d  %s" % (clk - t0, what))


if __name__ == "__main__":
    main()
