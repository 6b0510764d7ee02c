"""GPU measurement of the tree kernels against the HBM roofline (SURVEY.md section 8d byte model).

For G concurrent Connect4 games with grown trees, times noise+select, plan and expand+backup per minibatch (CUDA events
around each kernel group, engine profile level 2, single stream, nothing else running) and converts the ALGORITHMIC
bytes -- select d(12A+20) per descent, backup 20 d per backed-up descent, expand 16A+12 per new node, plan 17 per
descent + 17 per unique leaf -- into GB/s.  Usage: python tools/tree_bench.py [games ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, Net

HBM_GBS = 6539.2  # MEASURED_PEAKS.json


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 16384]
    game = ConnectFour()
    A = game.action_space
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    for G in sizes:
        eng = SelfPlayEngine(game, G, max_batch=8, node_capacity=24576 if G <= 8192 else 8192, seed=7)
        eng.play(dn, dn, moves=6, count=100, batch=8, tau_plies=10, auto_restart=True)  # mid-game trees
        eng.search(dn, 40, 8)                                                             # grow the current roots' trees
        torch.cuda.synchronize()
        n = 30
        depth_sum = 0
        c0 = eng.counters()
        eng.profile(2)
        for i in range(n):
            eng.search(dn, 1, 8)
            depth_sum += int(eng.region("desc_path_len").reshape(-1)[: G * 8].to(torch.int64).sum().item())
        p = eng.profile_read()
        eng.profile(0)
        c1 = eng.counters()
        desc = c1["descents"] - c0["descents"]
        leaves = c1["leaf_evals"] - c0["leaf_evals"]
        d = depth_sum / max(1, desc)
        sel_b = depth_sum * (12 * A + 20)
        bak_b = depth_sum * 20 + leaves * (16 * A + 12)
        plan_b = desc * 17 + leaves * 17
        out = {"games": G, "descents_per_minibatch": desc / n, "avg_path_len": d, "unique_leaves_per_minibatch": leaves / n}
        for name, ms, b in (("noise+select", p["select_ms"], sel_b), ("plan", p["plan_ms"], plan_b), ("expand+backup", p["expand_backup_ms"], bak_b)):
            gbs = b / (ms / 1e3) / 1e9
            out[name] = {"us_per_minibatch": 1e3 * ms / n, "algorithmic_MB_per_minibatch": b / n / 1e6, "GB_per_s": gbs,
                         "frac_of_hbm_peak": gbs / HBM_GBS}
        print(json.dumps(out), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
