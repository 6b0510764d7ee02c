"""GPU experiment: host enqueue time vs GPU time for single-stream and pair-pipelined self-play plies."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net


def run(label, fn, plies):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn(plies)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-28s host enqueue %.2f ms/ply, gpu %.2f ms/ply, wall %.2f ms/ply" % (
        label, 1e3 * (t1 - t0) / plies, e0.elapsed_time(e1) / plies, 1e3 * (t2 - t0) / plies), flush=True)


def main():
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    G = 4096
    single = SelfPlayEngine(game, G, max_batch=8, node_capacity=24576, seed=1)
    single.play(dn, dn, moves=20, count=8, batch=8, tau_plies=10, auto_restart=True)
    run("single stream", lambda n: single.play(dn, dn, moves=n, count=100, batch=8, tau_plies=10, auto_restart=True), 20)
    del single
    a = SelfPlayEngine(game, G // 2, max_batch=8, node_capacity=24576, seed=2)
    b = SelfPlayEngine(game, G // 2, max_batch=8, node_capacity=24576, seed=3)
    a.play_pair(b, dn, moves=20, count=8, batch=8, tau_plies=10, auto_restart=True)
    run("pair pipeline", lambda n: a.play_pair(b, dn, moves=n, count=100, batch=8, tau_plies=10, auto_restart=True), 20)
    if len(sys.argv) > 2:
        n = int(sys.argv[2])
        es = [SelfPlayEngine(game, G, max_batch=8, node_capacity=24576, seed=10 + i) for i in range(n)]
        SelfPlayEngine.play_multi(es, dn, moves=20, count=8, batch=8, tau_plies=10, auto_restart=True)
        run("multi pipeline %dx4096" % n, lambda k: SelfPlayEngine.play_multi(es, dn, moves=k, count=100, batch=8, tau_plies=10, auto_restart=True), 20)
        return
    if len(sys.argv) > 1:
        a2 = SelfPlayEngine(game, G, max_batch=8, node_capacity=24576, seed=4)
        b2 = SelfPlayEngine(game, G, max_batch=8, node_capacity=24576, seed=5)
        a2.play_pair(b2, dn, moves=20, count=8, batch=8, tau_plies=10, auto_restart=True)
        run("pair pipeline 2x4096", lambda n: a2.play_pair(b2, dn, moves=n, count=100, batch=8, tau_plies=10, auto_restart=True), 20)


if __name__ == "__main__":
    main()
