"""GPU diagnostic: 026 vs 025 (shipped Connect4 checkpoints) through play_games_batched with different tower precisions,
and a mirror match (026 vs 026 on two handles) that must come out even."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch

from caro_ai_b200.game import ConnectFour
from caro_ai_b200.model import DeviceNet, load_checkpoint
from caro_ai_b200.utils import play_games_batched

CK = os.path.join(ROOT, "tests", "golden", "checkpoints")


def main():
    game = ConnectFour()
    n026 = load_checkpoint(os.path.join(CK, "connect4_best_026_12000.dat"), game).eval()
    n025 = load_checkpoint(os.path.join(CK, "connect4_best_025_10600.dat"), game).eval()
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ("bf16x3",)  # e.g. bf16x3,fp32-simt: the tallies must agree
    settings = ((20, 16), (40, 8)) if len(precs) > 1 else ((20, 16), (40, 8), (20, 8), (10, 8), (80, 8))
    for prec in precs:
        a, a2, b = DeviceNet(n026, game, precision=prec), DeviceNet(n026, game, precision=prec), DeviceNet(n025, game, precision=prec)
        for tag, x, y in (("026 vs 026", a, a2), ("026 vs 025", a, b), ("025 vs 026", b, a)):
            for searches, batch in settings:
                s = play_games_batched(game, rounds, x, y, 0, searches, batch, trees_per_game=2, seed=7)
                print(prec, tag, "search_batch(%d,%d)" % (searches, batch), {k: s[k] for k in ("wins", "losses", "draws")}, flush=True)
        for d in (a, a2, b):
            d.close()


if __name__ == "__main__":
    main()
