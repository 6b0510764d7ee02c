#!/usr/bin/env python3
"""Generate tests/golden/*.json by RUNNING THE UNMODIFIED REFERENCE (read-only, /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (committed):
  games.json     random legal play-outs through the reference's ConnectFour and TicTacToe(n,k):
                 every transition, legal-move list, win flag and an exact plane checksum.
  mcts.json      reference MCTS.search_batch on several roots with the integer stub network
                 (oracle/stubs.py) and np.random.seed-ed Dirichlet noise: the complete N/W/Q/P
                 dictionaries (floats as hex) and the scalar type of every W entry.
  play_game.json reference lib.utils.play_game transcripts (stub network subclassing the
                 reference's Net so its isinstance assert holds; seeded global RNG).
The reference code is imported, never copied.
"""
import json
import os
import random
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import torch  # noqa: E402
from lib import mcts as ref_mcts, model as ref_model, utils as ref_utils  # noqa: E402
from lib.game.connect_four.connect_four import ConnectFour  # noqa: E402
from lib.game.tictactoe.tictactoe import TicTacToe  # noqa: E402
from oracle.stubs import stub_forward  # noqa: E402


def fhex(x):
    return float(x).hex()


def plane_checksum(planes: np.ndarray):
    """Exact, order-sensitive checksum of a 0/1 plane tensor."""
    flat = planes.reshape(len(planes), -1).astype(np.int64)
    w = (np.arange(flat.shape[1], dtype=np.int64) * 7919 + 13) % 1000003
    return [int(v) for v in (flat * w).sum(axis=1)]


def playouts(game, n_games, rng, tag):
    out = []
    for _ in range(n_games):
        s = game.initial_state
        who = rng.randrange(2)
        steps = []
        while True:
            legal = game.possible_moves(s)
            if not legal:
                break
            a = rng.choice(legal)
            s2, won = game.move(s, a, who)
            planes = game.states_to_training_batch([s2, s2], [who, 1 - who])
            steps.append({"s": s, "a": int(a), "p": who, "s2": s2, "won": bool(won),
                          "legal2": [int(x) for x in game.possible_moves(s2)],
                          "planes": plane_checksum(planes)})
            s, who = s2, 1 - who
            if won:
                break
        out.append(steps)
    return {"game": tag, "games": out}


class StubNet(ref_model.Net):
    """Reference Net subclass (passes utils.py:52-53) whose forward is the integer stub."""

    def __init__(self, game):
        super().__init__(game.obs_shape, game.action_space)
        self.actions_n = game.action_space

    def forward(self, x):
        return stub_forward(x, self.actions_n)


def dump_tree(tree):
    nodes = {}
    for s in tree.probs:
        nodes[str(s)] = {
            "N": [int(n) for n in tree.visit_count[s]],
            "W": [fhex(w) for w in tree.value[s]],
            "Wt": "".join("n" if isinstance(w, np.floating) else "f" for w in tree.value[s]),
            "Q": [fhex(q) for q in tree.value_avg[s]],
            "Qt": "".join("n" if isinstance(q, np.floating) else "f" for q in tree.value_avg[s]),
            "P": [fhex(p) for p in tree.probs[s]],
        }
    return nodes


def random_position(game, rng, plies):
    """A non-terminal position reached by `plies` random legal moves."""
    while True:
        s, who, ok = game.initial_state, rng.randrange(2), True
        for _ in range(plies):
            legal = game.possible_moves(s)
            if not legal:
                ok = False
                break
            s, won = game.move(s, rng.choice(legal), who)
            who = 1 - who
            if won:
                ok = False
                break
        if ok and game.possible_moves(s):
            return s, who


def mcts_cases(rng):
    cases = []
    specs = [("connect4", ConnectFour(), None, [0, 6, 14, 30, 36], 12, 8),
             ("mnk", TicTacToe(3, 3), (3, 3), [0, 2, 5, 6], 10, 8),
             ("mnk", TicTacToe(5, 4), (5, 4), [0, 9, 18], 8, 8),
             ("mnk", TicTacToe(15, 5), (15, 5), [0, 40], 4, 8)]
    seed = 1000
    for tag, game, nk, plies_list, count, bs in specs:
        net = lambda x, g=game: stub_forward(x, g.action_space)
        for plies in plies_list:
            root, who = random_position(game, rng, plies)
            tree = ref_mcts.MCTS(game)
            seed += 1
            np.random.seed(seed)
            # two consecutive searches on the same tree, the second from a child: exercises
            # tree persistence across moves the way play_game uses it
            tree.search_batch(count, bs, root, who, net)
            pi1, q1 = tree.get_policy_value(root, tau=1)
            a = int(np.argmax(tree.visit_count[root]))
            child, won = game.move(root, a, who)
            second = None
            if not won and game.possible_moves(child):
                tree.search_batch(count // 2, bs, child, 1 - who, net)
                pi2, _ = tree.get_policy_value(child, tau=0)
                second = {"root": child, "player": 1 - who, "count": count // 2, "pi_tau0": pi2}
            cases.append({"game": tag, "nk": nk, "seed": seed, "root": root, "player": who,
                          "count": count, "batch": bs, "pi_tau1": [fhex(p) for p in pi1],
                          "q_root": [fhex(q) for q in q1], "second": second,
                          "len": len(tree), "tree": dump_tree(tree)})
    return cases


def play_game_cases():
    cases = []
    for tag, game, nk, searches, bs, tau_steps, n in [
            ("connect4", ConnectFour(), None, 6, 8, 4, 3),
            ("mnk", TicTacToe(3, 3), (3, 3), 5, 8, 3, 4),
            ("mnk", TicTacToe(5, 4), (5, 4), 3, 4, 5, 2)]:
        net = StubNet(game)
        for i in range(n):
            seed = 7000 + 17 * len(cases)
            np.random.seed(seed)
            import collections
            replay = collections.deque(maxlen=10000)
            # mcts_stores: None -> two private trees (play.py); MCTS -> one shared tree (train.py)
            shared = (i % 2 == 1)
            stores = ref_mcts.MCTS(game) if shared else None
            res, steps = ref_utils.play_game(game, stores, replay, net, net, tau_steps, searches, bs)
            cases.append({"game": tag, "nk": nk, "seed": seed, "shared_tree": shared,
                          "searches": searches, "batch": bs, "tau_steps": tau_steps,
                          "result": res, "steps": steps,
                          "replay": [[s, int(p), [fhex(x) for x in pi], int(z)] for s, p, pi, z in replay]})
    return cases


def net_cases():
    """Reference Net (.eval()) outputs on fixed boards, for the shipped checkpoints and a seeded
    random init: pins oracle/net.py and the checkpoint format."""
    out = []
    rng = random.Random(5)
    for tag, game, ck in [("connect4", ConnectFour(), "saves/trained_connect4/best_026_12000.dat"),
                          ("connect4", ConnectFour(), None),
                          ("mnk", TicTacToe(3, 3), "saves/trained_tictactoe/best_005_00900.dat")]:
        torch.manual_seed(0)
        net = ref_model.Net(game.obs_shape, game.action_space)
        if ck:
            net.load_state_dict(torch.load(os.path.join(REF, ck), map_location="cpu"))
        net.eval()
        states, players = [], []
        for plies in [0, 1, 3, 5, 7]:
            s, who = random_position(game, rng, plies)
            states.append(s)
            players.append(who)
        with torch.no_grad():
            logits, vals = net(torch.tensor(game.states_to_training_batch(states, players)))
        out.append({"game": tag, "checkpoint": ck, "seed": 0, "states": states, "players": players,
                    "logits": logits.numpy().astype(float).tolist(),
                    "values": vals.numpy()[:, 0].astype(float).tolist(),
                    "keys": {k: list(v.shape) for k, v in net.state_dict().items()}})
    return out


def main():
    rng = random.Random(20261018)
    games = [playouts(ConnectFour(), 40, rng, "connect4"),
             playouts(TicTacToe(3, 3), 40, rng, "mnk:3:3"),
             playouts(TicTacToe(5, 4), 24, rng, "mnk:5:4"),
             playouts(TicTacToe(15, 5), 3, rng, "mnk:15:5")]
    with open(os.path.join(HERE, "games.json"), "w") as f:
        json.dump(games, f, separators=(",", ":"))
    with open(os.path.join(HERE, "mcts.json"), "w") as f:
        json.dump(mcts_cases(rng), f, separators=(",", ":"))
    with open(os.path.join(HERE, "play_game.json"), "w") as f:
        json.dump(play_game_cases(), f, separators=(",", ":"))
    with open(os.path.join(HERE, "net.json"), "w") as f:
        json.dump(net_cases(), f, separators=(",", ":"))
    for n in ("games.json", "mcts.json", "play_game.json", "net.json"):
        print(n, os.path.getsize(os.path.join(HERE, n)))


if __name__ == "__main__":
    main()
