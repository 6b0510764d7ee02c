#!/usr/bin/env python3
"""Round-2 fixtures, again produced by RUNNING THE UNMODIFIED REFERENCE (read-only, /root/reference).

    python tests/golden/make_golden2.py          # build container only (the GPU box has no /root/reference)

Outputs (committed):
  render.json  `game.render(state)` strings of ~20 positions per game (connect_four.py:267-281, tictactoe.py:237-259)
               and `Session.render()` strings with and without a position evaluation (play_session.py:38-49).
  train.json   one call of the reference's `train.train_neural_net` (train.py:62-117: 10 SGD rounds of 256 samples,
               MSE + soft-target cross-entropy, lr 0.1, momentum 0.9) on a seeded synthetic replay buffer with a
               seeded Net: the per-round losses and checksums of the weights afterwards.  `train.py` imports
               tensorboardX (absent here), so a stub module of that name is registered before the import; none of
               the arithmetic touches it.
The reference code is imported, never copied.  make_golden.py's four files are left untouched.
"""
import collections
import json
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import torch  # noqa: E402

sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))
import train as ref_train  # noqa: E402
import config as ref_cfg  # noqa: E402
from lib import model as ref_model, play_session as ref_session  # noqa: E402
from lib.game.connect_four.connect_four import ConnectFour  # noqa: E402
from lib.game.tictactoe.tictactoe import TicTacToe  # noqa: E402

from make_golden import random_position  # noqa: E402


def render_cases(rng):
    out = []
    for tag, game in [("connect4", ConnectFour()), ("mnk:3:3", TicTacToe(3, 3)), ("mnk:5:4", TicTacToe(5, 4))]:
        cells = game.obs_shape[1] * game.obs_shape[2]
        rows = []
        for i in range(20):
            s, _ = random_position(game, rng, rng.randrange(0, max(1, cells - 2)))
            rows.append({"state": s, "render": game.render(s)})
        out.append({"game": tag, "positions": rows})
    # Session.render: the object is built without __init__ (no checkpoint / tree needed for the string)
    sess = []
    for tag, game in [("connect4", ConnectFour()), ("mnk:3:3", TicTacToe(3, 3))]:
        for value in (None, 0.0, -0.256, 0.999, np.float32(0.125)):
            s, _ = random_position(game, rng, 5)
            obj = ref_session.Session.__new__(ref_session.Session)
            obj.game, obj.state, obj.value = game, s, value
            sess.append({"game": tag, "state": s, "value": None if value is None else float(value), "render": obj.render()})
    return {"games": out, "sessions": sess}


class LossRecorder:
    def __init__(self):
        self.rows = {}

    def track(self, name, value, step):
        self.rows[name] = float(value)


def synthetic_replay(game, rng, n):
    """(state, player, probs, z) tuples shaped like lib/utils.py:101-106 writes them."""
    A = game.action_space
    buf = collections.deque(maxlen=ref_cfg.REPLAY_BUFFER)
    cells = game.obs_shape[1] * game.obs_shape[2]
    for _ in range(n):
        s, who = random_position(game, rng, rng.randrange(0, max(1, cells - 2)))
        legal = game.possible_moves(s)
        w = [rng.random() if a in legal else 0.0 for a in range(A)]
        tot = sum(w)
        buf.append((s, who, [float(np.float32(x / tot)) for x in w], rng.choice([-1, 0, 1])))
    return buf


def train_cases(rng):
    out = []
    for tag, game in [("connect4", ConnectFour()), ("mnk:3:3", TicTacToe(3, 3))]:
        replay = synthetic_replay(game, rng, 600)
        torch.manual_seed(11)
        net = ref_model.Net(game.obs_shape, game.action_space)
        init = {k: v.clone() for k, v in net.state_dict().items()}
        opt = torch.optim.SGD(net.parameters(), lr=ref_cfg.LEARNING_RATE, momentum=0.9)
        # per-round losses: the reference only reports the means, so the forward is wrapped to see every round
        rounds = []
        fwd = net.forward

        def spy(x, fwd=fwd, rounds=rounds):
            lg, v = fwd(x)
            rounds.append((lg, v))
            return lg, v
        net.forward = spy
        ref_train.net, ref_train.step_idx = net, 1
        rec = LossRecorder()
        random.seed(4242)
        ref_train.train_neural_net(game, replay, opt, rec, "cpu")
        net.forward = fwd
        random.seed(4242)
        batches = [random.sample(replay, ref_cfg.BATCH_SIZE) for _ in range(ref_cfg.TRAIN_ROUNDS)]
        first = batches[0]
        pv = torch.FloatTensor([b[2] for b in first])
        zv = torch.FloatTensor([b[3] for b in first])
        lg0, v0 = rounds[0]
        l_val = torch.nn.functional.mse_loss(v0.squeeze(-1), zv).item()
        l_pol = (-torch.nn.functional.log_softmax(lg0, dim=1) * pv).sum(dim=1).mean().item()
        sd = net.state_dict()
        out.append({"game": tag, "net_seed": 11, "sample_seed": 4242,
                    "replay": [[s, int(p), [float(x) for x in pi], int(z)] for s, p, pi, z in replay],
                    "mean_losses": rec.rows, "round0": {"loss_value": l_val, "loss_policy": l_pol},
                    "first_batch_states": [b[0] for b in first[:8]],
                    "init_sum": {k: float(v.double().sum()) for k, v in init.items() if v.dtype.is_floating_point},
                    "final_sum": {k: float(v.double().sum()) for k, v in sd.items() if v.dtype.is_floating_point},
                    "final_abs_sum": {k: float(v.double().abs().sum()) for k, v in sd.items() if v.dtype.is_floating_point}})
    return out


def main():
    rng = random.Random(20261019)
    with open(os.path.join(HERE, "render.json"), "w") as f:
        json.dump(render_cases(rng), f, ensure_ascii=True, separators=(",", ":"))
    with open(os.path.join(HERE, "train.json"), "w") as f:
        json.dump(train_cases(rng), f, separators=(",", ":"))
    for n in ("render.json", "train.json"):
        print(n, os.path.getsize(os.path.join(HERE, n)))


if __name__ == "__main__":
    main()
