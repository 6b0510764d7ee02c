"""Pins the oracle (oracle/*.py) to the reference.

Part 1 restates every golden vector the reference's own unit tests hold for the hot path
(lib/game/connect_four/test_connect_four.py, lib/game/tictactoe/test_tictactoe.py,
lib/game/tictactoe/test_tictactoe_helpers.py, lib/test_mcts.py -- file:line cited per test).
Part 2 replays the fixtures produced by running the unmodified reference
(tests/golden/make_golden.py).
"""
import collections

import numpy as np
import pytest

from helpers import fhex, oracle_game, plane_checksum
from oracle.games import ConnectFourOracle, MNKOracle
from oracle.mcts import OracleMCTS, play_game
from oracle.stubs import stub_forward

C4_EMPTY = 0b000000000000000000000000000000000000000000110110110110110110110
C4_BLACK = 0b111111111111111111111111111111111111111111000000000000000000000


# ---------------------------------------------------------------- part 1: reference unit tests
class TestConnectFourReferenceVectors:
    g = ConnectFourOracle()

    def test_encode_decode(self):  # test_connect_four.py:27-44
        assert self.g.encode_lists([[]] * 7) == C4_EMPTY == 1797558 == self.g.initial_state
        assert self.g.encode_lists([[1] * 6] * 7) == C4_BLACK
        assert self.g.encode_lists([[0] * 6] * 7) == 0
        assert self.g.decode_binary(C4_EMPTY) == [[]] * 7
        assert self.g.decode_binary(C4_BLACK) == [[1] * 6] * 7
        assert self.g.decode_binary(0) == [[0] * 6] * 7

    def test_possible_moves(self):  # test_connect_four.py:47-56
        assert self.g.possible_moves(0) == []
        assert self.g.possible_moves(C4_BLACK) == []
        assert self.g.possible_moves(C4_EMPTY) == [0, 1, 2, 3, 4, 5, 6]

    def test_vertical_win(self):  # test_connect_four.py:58-80
        f = self.g.initial_state
        for i in range(4):
            f, won = self.g.move(f, 0, 1)
            assert won == (i == 3)
            assert self.g.decode_binary(f) == [[1] * (i + 1)] + [[]] * 6

    def test_horizontal_win(self):  # test_connect_four.py:82-104
        f = self.g.initial_state
        for col, expect in [(0, False), (1, False), (3, False), (2, True)]:
            f, won = self.g.move(f, col, 1)
            assert won == expect
        assert self.g.decode_binary(f) == [[1], [1], [1], [1], [], [], []]

    def test_diagonals(self):  # test_connect_four.py:106-127
        f = self.g.encode_lists([[0, 0, 0, 1], [0, 0, 1], [0], [1], [], [], []])
        assert self.g.move(f, 2, 1)[1] is True
        assert self.g.move(f, 2, 0)[1] is False
        f = self.g.encode_lists([[], [0, 1], [0, 0, 1], [1, 0, 0, 1], [], [], []])
        assert self.g.move(f, 0, 1)[1] is True
        assert self.g.move(f, 0, 0)[1] is False

    def test_tricky(self):  # test_connect_four.py:129-141
        f = self.g.encode_lists([[0, 1, 1], [1, 0], [0, 1], [0, 0, 1], [0, 0], [1, 1, 1, 0], []])
        s, won = self.g.move(f, 4, 0)
        assert won is True
        assert s == 3531389463375529686

    def test_planes(self):  # test_connect_four.py:143-191
        s = self.g.encode_lists([[0, 1, 0], [0], [1, 1, 1], [], [1], [], []])
        batch = self.g.states_to_training_batch([s, s], [1, 0])
        mine = [[0] * 7, [0] * 7, [0] * 7, [0, 0, 1, 0, 0, 0, 0], [1, 0, 1, 0, 0, 0, 0], [0, 0, 1, 0, 1, 0, 0]]
        other = [[0] * 7, [0] * 7, [0] * 7, [1, 0, 0, 0, 0, 0, 0], [0] * 7, [1, 1, 0, 0, 0, 0, 0]]
        np.testing.assert_equal(batch, [[mine, other], [other, mine]])
        assert batch.dtype == np.float32


class TestTicTacToeReferenceVectors:
    g = MNKOracle(3, 3)

    def test_encode_decode(self):  # test_tictactoe.py:11-37
        assert self.g.encode_game_state([[1, 2, 0], [0, 2, 0], [1, 0, 0]]) == int("120020100")
        assert self.g.encode_game_state([[0, 0, 0], [1, 2, 0], [1, 0, 0]]) == int("000120100")
        assert self.g.convert_mcts_state_to_list_state(int("010220011")) == [[0, 1, 0], [2, 2, 0], [0, 1, 1]]
        assert self.g.convert_mcts_state_to_list_state(0) == [[0, 0, 0]] * 3
        assert self.g.initial_state == 222222222

    def test_moves_lists(self):  # test_tictactoe.py:40-58
        s = self.g.encode_game_state([[0, 1, 0], [2, 2, 0], [0, 1, 1]])
        assert self.g.possible_moves(s) == [3, 4]
        assert self.g.invalid_moves(s) == [0, 1, 2, 5, 6, 7, 8]
        s = self.g.encode_game_state([[2, 1, 2], [2, 2, 0], [0, 1, 2]])
        assert self.g.possible_moves(s) == [0, 2, 3, 4, 8]
        assert self.g.invalid_moves(s) == [1, 5, 6, 7]

    def test_planes(self):  # test_tictactoe.py:61-98
        batch = self.g.states_to_training_batch([int("001010221"), int("101222001")], [1, 0])
        b1 = [[[0, 0, 1], [0, 1, 0], [0, 0, 1]], [[1, 1, 0], [1, 0, 1], [0, 0, 0]]]
        b2 = [[[0, 1, 0], [0, 0, 0], [1, 1, 0]], [[1, 0, 1], [0, 0, 0], [0, 0, 1]]]
        np.testing.assert_equal(batch, [b1, b2])

    def test_moves(self):  # test_tictactoe.py:101-119
        b = int("222222222")
        for mv, pl, expect in [(1, 0, "202222222"), (5, 1, "202221222"), (8, 0, "202221220"), (7, 1, "202221210")]:
            b, won = self.g.move(b, mv, pl)
            assert won is False and b == int(expect)

    def test_wins(self):  # test_tictactoe.py:121-144
        for board, mv, pl, expect in [("002112122", 2, 0, "000112122"), ("021012212", 6, 0, "021012012"),
                                      ("021102212", 8, 0, "021102210"), ("120122012", 4, 0, "120102012"),
                                      ("120102222", 6, 1, "120102122")]:
            nb, won = self.g.move(int(board), mv, pl)
            assert won is True and nb == int(expect)

    def test_lines_and_runs(self):  # test_tictactoe_helpers.py:14-53
        g = MNKOracle(3, 3)
        board = [1, -1, 1, -1, -1, 0, 0, -1, 1]
        lines = lambda r, c: g._lines_through(board, r, c)
        assert lines(0, 0)[1] == [1, -1, 0] and lines(2, 1)[1] == [-1, -1, -1] and lines(1, 2)[1] == [1, 0, 1]
        assert lines(0, 0)[2] == [1, -1, 1] and lines(1, 0)[2] == [-1, -1] and lines(1, 2)[2] == [-1, 0]
        assert lines(2, 1)[2] == [-1, -1] and lines(1, 1)[2] == [1, -1, 1]
        assert lines(0, 0)[3] == [1] and lines(1, 0)[3] == [-1, -1] and lines(2, 1)[3] == [-1, 0]
        assert lines(1, 2)[3] == [-1, 0] and lines(1, 1)[3] == [0, -1, 1]
        run = MNKOracle._has_run
        assert run([1, 1, 1], 3, 1) and run([-1, -1, -1], 3, -1)
        assert not run([1, 0, 1], 3, 1) and not run([-1, -1, 1], 3, -1)
        assert run([1, 1, 1, 0], 3, 1) and run([0, -1, -1, -1], 3, -1)
        assert not run([1, 0, 1, 1], 3, 1) and not run([-1, 1, -1, 1], 3, -1)


def test_backup_known_answer():  # lib/test_mcts.py:9-38
    t = OracleMCTS(game=None)
    t.visit_count = {1: [0, 1], 2: [1, 0], 3: [0, 0]}
    t.value = {1: [0.0, 0.5], 2: [0.6, 0.0], 3: [0.0, 0.0]}
    t.value_avg = {1: [0.0, 0.5], 2: [0.6, 0.0], 3: [0.0, 0.0]}
    t.probs = {1: [0.1, 0.9], 2: [0.8, 0.2], 3: [0.7, 0.3]}
    t._backup(0.2, [1, 2, 3], [1, 0, 0])
    assert t.visit_count == {1: [0, 2], 2: [2, 0], 3: [1, 0]}
    assert t.value == {1: [0.0, 0.3], 2: [0.8, 0.0], 3: [-0.2, 0.0]}
    assert t.value_avg == {1: [0.0, 0.15], 2: [0.4, 0.0], 3: [-0.2, 0.0]}


# ---------------------------------------------------------------- part 2: reference-generated fixtures
def test_playouts_match_reference(golden_games):
    n = 0
    for block in golden_games:
        g = oracle_game(block["game"])
        for steps in block["games"]:
            for st in steps:
                s2, won = g.move(st["s"], st["a"], st["p"])
                assert s2 == st["s2"] and bool(won) == st["won"]
                assert g.possible_moves(s2) == st["legal2"]
                assert sorted(g.invalid_moves(s2)) == [a for a in range(g.action_space) if a not in st["legal2"]]
                planes = g.states_to_training_batch([s2, s2], [st["p"], 1 - st["p"]])
                assert plane_checksum(planes) == st["planes"]
                n += 1
    assert n > 1500


def _dump(tree):
    out = {}
    for s in tree.probs:
        out[str(s)] = {
            "N": [int(x) for x in tree.visit_count[s]],
            "W": [fhex(x) for x in tree.value[s]],
            "Wt": "".join("n" if isinstance(w, np.floating) else "f" for w in tree.value[s]),
            "Q": [fhex(x) for x in tree.value_avg[s]],
            "Qt": "".join("n" if isinstance(w, np.floating) else "f" for w in tree.value_avg[s]),
            "P": [fhex(x) for x in tree.probs[s]],
        }
    return out


def test_search_matches_reference_trees(golden_mcts):
    """Same stub net + same np.random seed -> identical N (ints), W/Q/P (bit-exact, incl. the
    python-float vs float32 type of every W/Q entry)."""
    for case in golden_mcts:
        g = oracle_game(case["game"], case["nk"])
        net = lambda x, a=g.action_space: stub_forward(x, a)
        tree = OracleMCTS(g)
        np.random.seed(case["seed"])
        tree.search_batch(case["count"], case["batch"], case["root"], case["player"], net)
        pi1, q1 = tree.get_policy_value(case["root"], tau=1)
        assert [fhex(p) for p in pi1] == case["pi_tau1"]
        assert [fhex(q) for q in q1] == case["q_root"]
        sec = case["second"]
        if sec is not None:
            tree.search_batch(sec["count"], case["batch"], sec["root"], sec["player"], net)
            assert tree.get_policy_value(sec["root"], tau=0)[0] == sec["pi_tau0"]
        assert len(tree) == case["len"]
        assert _dump(tree) == case["tree"]


def test_play_game_matches_reference(golden_play):
    class Stub:
        def __init__(self, a):
            self.a = a

        def __call__(self, x):
            return stub_forward(x, self.a)

    for case in golden_play:
        g = oracle_game(case["game"], case["nk"])
        net = Stub(g.action_space)
        np.random.seed(case["seed"])
        replay = collections.deque(maxlen=10000)
        stores = OracleMCTS(g) if case["shared_tree"] else None
        res, steps = play_game(g, stores, replay, net, net, case["tau_steps"], case["searches"], case["batch"])
        assert (res, steps) == (case["result"], case["steps"])
        got = [[s, int(p), [fhex(x) for x in pi], int(z)] for s, p, pi, z in replay]
        assert got == case["replay"]


def test_net_matches_reference(golden_net):
    """oracle/net.py has the reference's state_dict layout and (for a seeded init) its outputs."""
    import torch
    from oracle.net import OracleNet
    for case in golden_net:
        g = oracle_game(case["game"], (3, 3))
        torch.manual_seed(case["seed"])
        net = OracleNet(g.obs_shape, g.action_space)
        assert {k: list(v.shape) for k, v in net.state_dict().items()} == case["keys"]
        if case["checkpoint"] is None:  # same seed, same construction order -> same init
            net.eval()
            with torch.no_grad():
                logits, vals = net(torch.tensor(g.states_to_training_batch(case["states"], case["players"])))
            np.testing.assert_allclose(logits.numpy(), np.array(case["logits"]), atol=1e-5)
            np.testing.assert_allclose(vals.numpy()[:, 0], np.array(case["values"]), atol=1e-5)
