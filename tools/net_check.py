"""GPU diagnostic: max |prior| / |value| error of both towers vs the fp32 PyTorch reference and vs a
bf16-emulated PyTorch reference, per test network.  Usage: python tools/net_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

import test_gpu_parity as T
from caro_ai_b200.model import DeviceNet
from harness import oracle_for, random_position


def main():
    rng = np.random.default_rng(3)
    for tag, game, net in T._net_cases():
        og = oracle_for(game)
        cells = game.obs_shape[1] * game.obs_shape[2]
        count = 300 if cells < 100 else 40
        pos = [random_position(og, rng, int(rng.integers(0, min(40, max(1, cells - 4))))) for _ in range(count)]
        states, players = [p[0] for p in pos], [p[1] for p in pos]
        ref_p, ref_v = T._reference_outputs(game, net, states, players)
        dn = DeviceNet(net, game)
        for impl in (1, 0, 7, 3, 2):
            try:
                p, v = dn.forward_states(states, players, impl=impl)
                torch.cuda.synchronize()
                p, v = p.cpu().numpy(), v.cpu().numpy()
                print("%-14s impl=%d finite=%s dp=%.3e dv=%.3e argmax_agree=%.3f" % (
                    tag, impl, bool(np.isfinite(p).all() and np.isfinite(v).all()), np.abs(p - ref_p).max(),
                    np.abs(v - ref_v).max(), float((p.argmax(1) == ref_p.argmax(1)).mean())), flush=True)
            except Exception as exc:  # noqa: BLE001
                print("%-14s impl=%d FAILED: %s" % (tag, impl, exc), flush=True)
                raise
        dn.close()


if __name__ == "__main__":
    main()
