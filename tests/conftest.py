import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_games():
    return load_golden("games.json")


@pytest.fixture(scope="session")
def golden_mcts():
    return load_golden("mcts.json")


@pytest.fixture(scope="session")
def golden_play():
    return load_golden("play_game.json")


@pytest.fixture(scope="session")
def golden_net():
    return load_golden("net.json")


@pytest.fixture(scope="session")
def golden_render():
    return load_golden("render.json")


@pytest.fixture(scope="session")
def golden_train():
    return load_golden("train.json")
