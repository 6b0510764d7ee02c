"""Shared helpers for the test-suite (oracle construction from fixture tags, float-hex utils)."""
import numpy as np

from oracle.games import ConnectFourOracle, MNKOracle


def oracle_game(tag, nk=None):
    if tag == "connect4":
        return ConnectFourOracle()
    if tag.startswith("mnk:"):
        _, n, k = tag.split(":")
        return MNKOracle(int(n), int(k))
    if tag == "mnk":
        return MNKOracle(int(nk[0]), int(nk[1]))
    raise ValueError(tag)


def fhex(x):
    return float(x).hex()


def plane_checksum(planes):
    flat = np.asarray(planes).reshape(len(planes), -1).astype(np.int64)
    w = (np.arange(flat.shape[1], dtype=np.int64) * 7919 + 13) % 1000003
    return [int(v) for v in (flat * w).sum(axis=1)]
