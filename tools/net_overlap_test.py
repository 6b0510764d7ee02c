"""GPU experiment: do two tower launches on two streams overlap (the tail of one with the head of the other)?
Usage: python tools/net_overlap_test.py [leaves]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def main():
    leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 9917
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    boards = torch.zeros((leaves, 2), dtype=torch.int64, device="cuda")
    who = torch.zeros(leaves, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    from caro_ai_b200 import _cabi
    probs = [torch.empty((leaves, 7), dtype=torch.float32, device="cuda") for _ in range(2)]
    values = [torch.empty((leaves,), dtype=torch.float32, device="cuda") for _ in range(2)]
    fwd = _cabi.lib().caro_net_forward

    def launch(st, j):
        _cabi.check(fwd(dn.handle, game.game_kind, game.n, game.k, boards.data_ptr(), who.data_ptr(), None, leaves,
                        probs[j].data_ptr(), values[j].data_ptr(), 0, st.cuda_stream))

    for _ in range(3):
        launch(s1, 0)
    torch.cuda.synchronize()
    for mode in ("one stream", "two streams"):
        iters = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for i in range(iters):
            j = 0 if (mode == "one stream" or i % 2 == 0) else 1
            launch(s1 if j == 0 else s2, j)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        print("%-12s %d leaves: %.1f us per launch" % (mode, leaves, 1e3 * e0.elapsed_time(e1) / iters), flush=True)


if __name__ == "__main__":
    main()
