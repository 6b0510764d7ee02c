"""GPU edge cases of the hot path (run with -m gpu): empty batches, boards played to the last cell (draws), the largest
board, a full arena, games that are all finished.  The oracle is only the checker."""
import numpy as np
import pytest

from harness import oracle_for

pytestmark = pytest.mark.gpu

# play-outs to a full board without a winner, found off-line with the oracle (first player, moves)
C4_DRAW = (0, [5, 0, 0, 4, 6, 3, 0, 5, 6, 2, 0, 2, 2, 2, 4, 6, 4, 5, 6, 2, 6, 0, 3, 3, 0, 6, 3, 5, 5, 4, 1, 4, 5, 2, 1, 4, 3, 1,
               3, 1, 1, 1])
MNK44_DRAW = (0, [13, 14, 4, 0, 1, 7, 9, 3, 6, 8, 15, 11, 12, 5, 10, 2])
MNK65_DRAW = (0, [33, 14, 24, 8, 2, 30, 34, 1, 18, 26, 16, 0, 25, 3, 20, 9, 15, 5, 31, 17, 27, 28, 11, 32, 7, 10, 6, 35, 19, 12,
                  29, 23, 21, 22, 4, 13])


def caro_draw_sequence():
    """15 x 15, five in a row: the colouring ((2 r + c + 1) // 2) % 2 has no run of five in any direction and 113 / 112
    cells per colour, so interleaving the two colours' cells is a legal 225-ply game that ends in a draw."""
    n = 15
    cells = {0: [], 1: []}
    for r in range(n):
        for c in range(n):
            cells[((2 * r + c + 1) // 2) % 2].append(r * n + c)
    assert len(cells[1]) == 113 and len(cells[0]) == 112
    seq = []
    for i in range(113):
        seq.append(cells[1][i])
        if i < 112:
            seq.append(cells[0][i])
    return 1, seq


def product(tag):
    from caro_ai_b200.game import ConnectFour, TicTacToe
    return ConnectFour() if tag == "c4" else TicTacToe(*tag)


@pytest.mark.parametrize("tag,draw", [("c4", C4_DRAW), ((4, 4), MNK44_DRAW), ((6, 5), MNK65_DRAW), ((15, 5), None)])
def test_games_played_to_the_last_cell(tag, draw):
    """Every ply of a drawn game (the longest game the board allows): states, win and draw flags, legal moves and planes
    of the CUDA kernels against the oracle; only the last ply raises the draw flag and leaves no legal move."""
    g = product(tag)
    og = oracle_for(g)
    first, seq = draw if draw is not None else caro_draw_sequence()
    states, players, nexts = [], [], []
    s, p = og.initial_state, first
    for a in seq:
        assert a in og.possible_moves(s)
        s2, won = og.move(s, a, p)
        assert not won
        states.append(s)
        players.append(p)
        nexts.append(s2)
        s, p = s2, 1 - p
    assert og.possible_moves(s) == [] and len(seq) == g.action_space * (6 if tag == "c4" else 1)
    new_states, won, draw_flag = g.apply_batch(states, seq, players)
    assert new_states == nexts
    assert not won.any()
    assert draw_flag.tolist() == [0] * (len(seq) - 1) + [1]
    masks = g.legal_masks(nexts)
    for m, st in zip(masks, nexts):
        assert [int(a) for a in np.nonzero(m)[0]] == og.possible_moves(st)
    assert not masks[-1].any()
    who = [1 - q for q in players]
    np.testing.assert_array_equal(g.states_to_training_batch(nexts, who), og.states_to_training_batch(nexts, who))
    # the facade's single-position calls on the final position
    assert g.possible_moves(nexts[-1]) == [] and sorted(g.invalid_moves(nexts[-1])) == list(range(g.action_space))


def test_empty_batches():
    """Zero positions / zero leaves: shaped empty results, no launch, no error (the reference's list comprehensions
    over empty lists)."""
    import torch
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    for g in (ConnectFour(), TicTacToe(3, 3), TicTacToe(15, 5)):
        new_states, won, draw = g.apply_batch([], [], [])
        assert new_states == [] and won.shape == (0,) and draw.shape == (0,)
        assert g.legal_masks([]).shape == (0, g.action_space)
        assert g.states_to_training_batch([], []).shape == (0,) + tuple(g.obs_shape)
    g = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(g.obs_shape, g.action_space).eval(), g)
    pri, val = dn.forward_states([], [])
    assert tuple(pri.shape) == (0, 7) and tuple(val.shape) == (0,)
    # a device-side leaf count of zero: the tower launches, finds nothing to do and leaves the outputs untouched
    boards = torch.zeros((64, 2), dtype=torch.int64, device="cuda")
    who = torch.zeros(64, dtype=torch.uint8, device="cuda")
    from caro_ai_b200 import _cabi
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    probs = torch.full((64, 7), -7.0, device="cuda")
    values = torch.full((64,), -7.0, device="cuda")
    _cabi.check(_cabi.lib().caro_net_forward(dn.handle, g.game_kind, g.n, g.k, boards.data_ptr(), who.data_ptr(), count.data_ptr(),
                                             64, probs.data_ptr(), values.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert (probs == -7.0).all() and (values == -7.0).all()
    dn.close()


def test_full_arena_is_reported_and_search_continues():
    """A node arena that is too small: the engine raises the arena-full error bit (include/caro_b200.h, counters[7]),
    never writes past the arena, and the visit counts still add up (a leaf that found no room is still backed up)."""
    import torch
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import DeviceNet, Net
    g = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(g.obs_shape, g.action_space).eval(), g)
    cap = 32
    eng = SelfPlayEngine(g, 16, max_batch=8, node_capacity=cap, seed=3)
    eng.search(dn, 20, 8)
    torch.cuda.synchronize()
    c = eng.counters()
    assert c["errors"] & 1, c
    counts = eng.region("node_count").cpu().numpy()
    assert (counts == cap).all()
    assert c["descents"] == 16 * 8 * 20
    # root visit counts: every descent that reached an expanded root added one visit to exactly one root edge
    pi, q, n = eng.root_policy(1)
    assert int(n.sum(dim=1).min().item()) > 0 and int(n.sum(dim=1).max().item()) <= 8 * 20
    # every record of every arena is a well-formed node: an overflowing arena did not spill into its neighbour
    assert eng.region("nodes").shape[0] == 16 * cap
    links = eng.pool("C")[:, :7]
    assert int(links.min().item()) >= -1 and int(links.max().item()) < cap
    np.testing.assert_allclose(eng.pool("P")[:, :7].sum(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    assert int(eng.pool("N").min().item()) >= 0 and int(eng.pool("N").sum().item()) > 0
    eng.close()
    dn.close()


def test_all_games_finished_is_a_no_op():
    """Lock-step search over games that have all ended (no re-seating): no descents are planned, no leaf reaches the
    network, counters stay put."""
    import torch
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    g = TicTacToe(3, 3)
    torch.manual_seed(0)
    dn = DeviceNet(Net(g.obs_shape, g.action_space).eval(), g)
    eng = SelfPlayEngine(g, 32, max_batch=8, node_capacity=512, seed=5)
    eng.play(dn, dn, moves=9, count=4, batch=8, tau_plies=2, auto_restart=False)  # nine plies end every 3x3 game
    torch.cuda.synchronize()
    assert (eng.region("status") == 1).all()
    c0 = eng.counters()
    assert c0["games"] == 32 and c0["wins_p0"] + c0["wins_p1"] + c0["draws"] == 32
    eng.search(dn, 3, 8)
    eng.play(dn, dn, moves=2, count=4, batch=8, tau_plies=2, auto_restart=False)
    torch.cuda.synchronize()
    c1 = eng.counters()
    assert c1 == c0 and eng.leaf_count() == 0
    eng.close()
    dn.close()


def test_large_launch_select_matches_small_launch():
    """More descents per launch than fit the GPU at once (> 94,720) take the 72-register build of the select kernel: the
    first 64 of 12,288 games must grow exactly the trees of a 64-game engine with the same seed (Philox streams are
    addressed by game, the network's result for a leaf does not depend on the batch it travels in)."""
    import torch
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import DeviceNet, Net
    g = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(g.obs_shape, g.action_space).eval(), g)
    cap = 128
    big = SelfPlayEngine(g, 12288, max_batch=8, node_capacity=cap, seed=21)
    small = SelfPlayEngine(g, 64, max_batch=8, node_capacity=cap, seed=21)
    for e in (big, small):
        e.search(dn, 6, 8)
    torch.cuda.synchronize()
    assert big.counters()["errors"] == 0 and small.counters()["errors"] == 0
    assert big.roots()[1][:64] == small.roots()[1]
    assert torch.equal(big.region("node_count")[:64], small.region("node_count")[:64])
    assert int(small.region("node_count").min().item()) > 6
    assert torch.equal(big.region("nodes")[: 64 * cap], small.region("nodes")[: 64 * cap])
    big.close()
    small.close()
    dn.close()
