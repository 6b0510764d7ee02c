"""CPU-side checks of the product's host logic and of the __host__ __device__ rule arithmetic.

* state-int <-> device-board conversions of caro_ai_b200.game (pure host code) round-trip on the
  reference-generated play-outs;
* csrc/rules.cuh compiled for the host (tests/native/hostcheck.cpp, TEST-ONLY, gcc) reproduces
  every golden transition: same bitboards, win flag, legal moves, planes;
* the Philox generator matches the published Random123 known-answer vectors, and the Gamma(0.3)
  sampler has the right mean/variance.
The kernels themselves are exercised by tests/test_gpu_parity.py on the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import oracle_game, plane_checksum

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "hostcheck.cpp")
SO = os.path.join(HERE, "native", "libhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    deps = [SRC, os.path.join(HERE, "..", "caro-ai_b200", "csrc", "rules.cuh"), os.path.join(HERE, "..", "caro-ai_b200", "csrc", "rng.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", SO])
    lib = C.CDLL(SO)
    lib.hc_c4_key.restype = C.c_uint64
    lib.hc_c4_key.argtypes = [C.c_uint64, C.c_uint64]
    lib.hc_c4_legal.argtypes = [C.c_uint64, C.c_uint64]
    lib.hc_c4_plane.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.hc_gamma.restype = C.c_float
    lib.hc_gamma.argtypes = [C.c_float] + [C.c_uint32] * 5
    return lib


def product_game(tag):
    from caro_ai_b200.game import ConnectFour, TicTacToe
    if tag == "connect4":
        return ConnectFour()
    _, n, k = tag.split(":")
    return TicTacToe(int(n), int(k))


def test_state_conversions_roundtrip(golden_games):
    for block in golden_games:
        g = product_game(block["game"])
        assert g.words_to_state(g.state_to_words(g.initial_state)) == g.initial_state
        for steps in block["games"]:
            for st in steps:
                assert g.words_to_state(g.state_to_words(st["s2"])) == st["s2"]


def test_facade_encodings_match_reference_vectors():
    from caro_ai_b200.game import ConnectFour, TicTacToe
    g = ConnectFour()
    assert g.initial_state == 0b110110110110110110110 == g.encode_lists([[]] * 7)  # test_connect_four.py:28-30
    assert g.encode_lists([[1] * 6] * 7) == 0b111111111111111111111111111111111111111111000000000000000000000
    assert g.encode_lists([[0] * 6] * 7) == 0
    assert g.decode_binary(0) == [[0] * 6] * 7 and g.decode_binary(g.initial_state) == [[]] * 7
    t = TicTacToe(3, 3)
    assert t._pad_mcts_state("120120") == "000120120" and t._pad_mcts_state("0") == "000000000"  # test_tictactoe.py:12-14
    assert t.encode_game_state([[1, 2, 0], [0, 2, 0], [1, 0, 0]]) == int("120020100")
    assert t.convert_mcts_state_to_list_state(int("010220011")) == [[0, 1, 0], [2, 2, 0], [0, 1, 1]]
    assert t.initial_state == 222222222
    assert "0123456" in g.render(g.initial_state) and t.render(t.initial_state).startswith("|0|1|2|")


def test_connect4_rules_on_host(hc, golden_games):
    block = [b for b in golden_games if b["game"] == "connect4"][0]
    g = product_game("connect4")
    seen_keys = {}
    for steps in block["games"]:
        for st in steps:
            mask, black = g.state_to_words(st["s"])
            m, b = C.c_uint64(mask), C.c_uint64(black)
            won = hc.hc_c4_apply(C.byref(m), C.byref(b), st["a"], st["p"])
            assert g.words_to_state([m.value, b.value]) == st["s2"]
            assert bool(won) == st["won"]
            legal = hc.hc_c4_legal(m.value, b.value)
            assert [c for c in range(7) if (legal >> c) & 1] == st["legal2"]
            assert bool((legal >> 8) & 1) == (len(st["legal2"]) > 0)
            key = hc.hc_c4_key(m.value, b.value)
            assert seen_keys.setdefault(key, st["s2"]) == st["s2"]  # injective on everything seen
            planes = np.zeros((2, 2, 6, 7), np.float32)
            for v, who in enumerate((st["p"], 1 - st["p"])):
                for pl in range(2):
                    for r in range(6):
                        for c in range(7):
                            planes[v, pl, r, c] = hc.hc_c4_plane(m.value, b.value, who, pl, r, c)
            assert plane_checksum(planes) == st["planes"]


@pytest.mark.parametrize("tag", ["mnk:3:3", "mnk:5:4", "mnk:15:5"])
def test_mnk_rules_on_host(hc, golden_games, tag):
    block = [b for b in golden_games if b["game"] == tag][0]
    g = product_game(tag)
    n, k = g.board_len, g.k_to_win
    keys = {}
    for gi, steps in enumerate(block["games"]):
        for si, st in enumerate(steps):
            words = (C.c_uint64 * 8)(*g.state_to_words(st["s"]))
            res = hc.hc_mnk_apply(n, k, words, st["a"], st["p"])
            assert g.words_to_state(list(words)) == st["s2"]
            assert bool(res & 1) == st["won"]
            assert bool(res & 2) == (len(st["legal2"]) > 0)
            legal = [a for a in range(n * n) if hc.hc_mnk_legal(n, k, words, a)]
            assert legal == st["legal2"]
            out = (C.c_uint64 * 2)()
            hc.hc_mnk_key(n, k, words, out)
            assert keys.setdefault((out[0], out[1]), st["s2"]) == st["s2"]
            if n <= 5 or si % 16 == 0:
                planes = np.zeros((2, 2, n, n), np.float32)
                for v, who in enumerate((st["p"], 1 - st["p"])):
                    for pl in range(2):
                        for r in range(n):
                            for c in range(n):
                                planes[v, pl, r, c] = hc.hc_mnk_plane(n, k, words, who, pl, r, c)
                assert plane_checksum(planes) == st["planes"]
    assert len(keys) > 50


def test_philox_known_answers(hc):
    """Random123 kat_vectors: philox4x32-10."""
    out = (C.c_uint32 * 4)()
    hc.hc_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    hc.hc_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    hc.hc_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_gamma_sampler_moments(hc):
    xs = np.array([hc.hc_gamma(0.3, 1, 2, i, 7, 9) for i in range(40000)], dtype=np.float64)
    assert np.all(xs > 0) and np.all(np.isfinite(xs))
    assert abs(xs.mean() - 0.3) < 0.01          # E = alpha
    assert abs(xs.var() - 0.3) < 0.02           # Var = alpha
    # Dirichlet(0.3 x 7) marginal mean 1/7
    g = np.array([[hc.hc_gamma(0.3, 3, 4, i, 0, a) for a in range(7)] for i in range(5000)])
    d = g / g.sum(axis=1, keepdims=True)
    assert np.allclose(d.mean(axis=0), 1 / 7, atol=0.01)
