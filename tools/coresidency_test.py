"""GPU experiment: can the tree kernels become resident next to a running tensor-core tower CTA?
Times select+plan of one engine alone, and again while a long network pass runs on another stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    eng = SelfPlayEngine(game, 4096, max_batch=8, node_capacity=24576, seed=1)
    eng.play(dn, dn, moves=6, count=20, batch=8, tau_plies=10, auto_restart=True)  # grow some trees
    leaves = 2368 * 30  # ~1 ms of network
    boards = torch.zeros((leaves, 2), dtype=torch.int64, device="cuda")
    who = torch.zeros(leaves, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()

    x = torch.zeros(1024, device="cuda")

    def tree(mb):
        if len(sys.argv) > 1 and sys.argv[1] == "tiny":
            x.add_(1.0)  # a 1-block elementwise kernel: can ANYTHING become resident next to the tower?
            return
        eng.select(8, mb)
        eng.plan(8)

    for concurrent in (False, True, False, True):
        torch.cuda.synchronize()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if concurrent:
            with torch.cuda.stream(s1):
                n0.record()
                dn.forward_boards(boards, who, leaves, 0)
                n1.record()
            torch.cuda._sleep(400000)  # ~0.2 ms of host-visible delay on the default stream: let the tower occupy the SMs
        with torch.cuda.stream(s2):
            if concurrent:
                s2.wait_stream(torch.cuda.current_stream())
            t0.record()
            tree(3)
            t1.record()
        torch.cuda.synchronize()
        msg = "tree (noise+select+plan) %.1f us" % (1e3 * t0.elapsed_time(t1))
        if concurrent:
            msg += " | network %.1f us, tree finished %.1f us after the network started" % (
                1e3 * n0.elapsed_time(n1), 1e3 * n0.elapsed_time(t1))
        print("concurrent=%d  %s" % (concurrent, msg), flush=True)


if __name__ == "__main__":
    main()
