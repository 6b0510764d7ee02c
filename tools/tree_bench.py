"""GPU measurement of the tree kernels against the HBM roofline (SURVEY.md section 8d byte model).

For G concurrent Connect4 games with grown trees, times noise+select, plan and expand+backup per minibatch (CUDA events
around each kernel, engine profile level 2, single stream, nothing else running) and converts the ALGORITHMIC
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


def board_kernels(game):
    """Streaming board kernels (caro_boards_apply / legal_mask / encode_planes) on 16 M random Connect4 positions:
    algorithmic bytes = 16 B board in (+ 16 B out + 5 B action/player + 2 B flags for apply; + 4 B mask; + 336 B planes)."""
    import ctypes as C
    import numpy as np
    from caro_ai_b200 import _cabi
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(1)
    heights = torch.randint(0, 6, (n, 7), device="cuda", generator=g)           # never full: every column stays legal
    colbits = torch.randint(0, 64, (n, 7), device="cuda", generator=g)
    shifts = (7 * torch.arange(7, device="cuda")).view(1, 7)
    col_mask = (torch.ones_like(heights) << heights) - 1
    mask = (col_mask << shifts).sum(1)
    black = ((colbits & col_mask) << shifts).sum(1)
    boards = torch.stack([mask, black], 1).contiguous()
    acts = torch.randint(0, 7, (n,), device="cuda", dtype=torch.int32, generator=g)
    pl = torch.randint(0, 2, (n,), device="cuda", generator=g).to(torch.uint8)
    out = torch.empty_like(boards)
    won = torch.empty(n, dtype=torch.uint8, device="cuda")
    draw = torch.empty(n, dtype=torch.uint8, device="cuda")
    lm = torch.empty(n, dtype=torch.int32, device="cuda")
    np_ = 1 << 22
    planes = torch.empty((np_, 2, 6, 7), dtype=torch.float32, device="cuda")
    lib, st = _cabi.lib(), torch.cuda.current_stream().cuda_stream
    jobs = {
        "boards_apply": (lambda: lib.caro_boards_apply(game.game_kind, 0, 0, boards.data_ptr(), acts.data_ptr(), pl.data_ptr(), n,
                                                      out.data_ptr(), won.data_ptr(), draw.data_ptr(), st), n * (16 + 16 + 5 + 2)),
        "boards_legal_mask": (lambda: lib.caro_boards_legal_mask(game.game_kind, 0, 0, boards.data_ptr(), n, lm.data_ptr(), st), n * 20),
        "boards_encode_planes": (lambda: lib.caro_boards_encode_planes(game.game_kind, 0, 0, boards.data_ptr(), pl.data_ptr(), np_,
                                                                      planes.data_ptr(), st), np_ * (17 + 336)),
    }
    res = {"positions": n}
    for name, (fn, nbytes) in jobs.items():
        for _ in range(3):
            _cabi.check(fn())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _cabi.check(fn())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = {"ms": ms, "GB_per_s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / HBM_GBS}
    print(json.dumps({"board_kernels": res}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "boards":
        board_kernels(ConnectFour())
        return
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 16384]
    game = ConnectFour()
    A = game.action_space
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    for G in sizes:
        compact = os.environ.get("COMPACT") == "1"  # CARO_FLAG_COMPACT_TREE: packed arenas (the same search, bit for bit)
        cap = int(os.environ.get("CAP", "0")) or (24576 if G <= 8192 else 8192)
        eng = SelfPlayEngine(game, G, max_batch=8, node_capacity=cap, seed=7, compact_tree=compact)
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
        out = {"games": G, "compact_tree": compact, "node_capacity": cap, "descents_per_minibatch": desc / n, "avg_path_len": d, "unique_leaves_per_minibatch": leaves / n}
        noise_b = desc * A * 8  # float64 [descents][A] written by noise_kernel (read again by select: not in select's byte model)
        for name, ms, b in (("noise", p["noise_ms"], noise_b), ("select", p["select_ms"], sel_b), ("plan", p["plan_ms"], plan_b),
                            ("expand+backup", p["expand_backup_ms"], bak_b)):
            gbs = b / (ms / 1e3) / 1e9
            out[name] = {"us_per_minibatch": 1e3 * ms / n, "algorithmic_MB_per_minibatch": b / n / 1e6, "GB_per_s": gbs,
                         "frac_of_hbm_peak": gbs / HBM_GBS}
        print(json.dumps(out), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
