"""Caro 15,15,5 at search_batch(200,8) -- or, with GAME=c4trained, Connect4 at search_batch(100,8) with the reference's trained
checkpoint (split-precision tower) --: leaf evaluations/s against the number of pipeline parts and games per part.
Usage: [GAME=c4trained] [NET_SMS=n] [PLIES=n] [CAP=nodes per game] python tools/caro_sweep.py PARTSxGAMES[:flag] ...   (e.g. 2x1024 3x1024 2x1536:recycle)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour, TicTacToe
from caro_ai_b200.model import DeviceNet, Net, load_checkpoint


def main():
    trained = os.environ.get("GAME") == "c4trained"
    c4 = trained or os.environ.get("GAME") == "c4"
    game = ConnectFour() if c4 else TicTacToe(15, 5)
    count = 100 if c4 else 200
    torch.manual_seed(0)
    if trained:
        dn = DeviceNet(load_checkpoint(os.path.join(ROOT, "tests", "golden", "checkpoints", "connect4_best_026_12000.dat"), game).eval(), game)
    else:
        dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    if os.environ.get("NET_SMS"):
        dn.set_grid_limit(int(os.environ["NET_SMS"]))
    for spec in sys.argv[1:]:
        spec, _, flag = spec.partition(":")
        parts, games = (int(x) for x in spec.split("x"))
        plies = int(os.environ.get("PLIES", "3"))
        flags = {"recycle_tree": True} if flag == "recycle" else {"compact_tree": True} if flag == "compact" else {}
        engs = [SelfPlayEngine(game, games, max_batch=8, node_capacity=int(os.environ.get("CAP", "8192")), seed=7 * h, **flags) for h in range(parts)]
        SelfPlayEngine.play_multi(engs, dn, moves=1, count=count, batch=8, tau_plies=10, auto_restart=True)
        torch.cuda.synchronize()
        c0 = [e.counters() for e in engs]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        SelfPlayEngine.play_multi(engs, dn, moves=plies, count=count, batch=8, tau_plies=10, auto_restart=True)
        t1.record()
        torch.cuda.synchronize()
        c1 = [e.counters() for e in engs]
        sec = t0.elapsed_time(t1) / 1e3
        leaf = sum(b["leaf_evals"] - a["leaf_evals"] for a, b in zip(c0, c1))
        print(json.dumps({"parts": parts, "games_per_part": games, "flag": flag, "plies": plies, "ms_per_ply": 1e3 * sec / plies,
                          "leaf_evals_per_sec": leaf / sec, "leaves_per_launch": leaf / (plies * count * parts), "precision": dn.precision, "max_nodes": max(int(e.region("node_count").max().item()) for e in engs),
                          "errors": sum(c["errors"] for c in c1)}), flush=True)
        for e in engs:
            e.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
