"""Deterministic stand-ins for the policy/value network (TEST-ONLY, part of the oracle).

The reference's search accepts any callable ``net(planes) -> (logits, values)`` (lib/mcts.py:215);
``play_game`` additionally type-checks ``isinstance(net, model.Net)`` (lib/utils.py:52-53), which the
golden generator satisfies by subclassing the reference's Net and overriding ``forward`` with
``stub_forward`` below.

Two flavours, both pure functions of ONE board (no cross-row arithmetic), built from integer
hashes so that results are bit-identical on every machine:
  * ``stub_forward``  -- torch, returns (logits, values); the caller (reference / oracle) applies
                         its own softmax.  Used for the golden fixtures.
  * ``stub_priors``   -- numpy, returns (priors, values) directly with priors = w / sum(w) from
                         integer weights: one IEEE float32 division per entry, no exp().  Used
                         when the CUDA engine and the oracle must see *identical* network outputs.
"""
from __future__ import annotations

import numpy as np


def _weights(n_in: int, n_out: int, salt: int) -> np.ndarray:
    j = np.arange(n_in, dtype=np.uint64)[:, None]
    a = np.arange(n_out, dtype=np.uint64)[None, :]
    h = (j * np.uint64(2654435761) + a * np.uint64(40503) + np.uint64(salt)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return (h >> np.uint64(20)).astype(np.int64)  # 0..4095


def _scores(planes: np.ndarray, actions_n: int):
    """Exact integer projections of the 0/1 planes: s[L,A] and t[L]."""
    x = np.asarray(planes).reshape(len(planes), -1).astype(np.int64)
    s = x @ _weights(x.shape[1], actions_n, 12345)
    t = (x @ _weights(x.shape[1], 1, 777))[:, 0]
    return s, t


def stub_forward(planes_tensor, actions_n: int):
    """(logits [L,A] in [-4,4), values [L,1] in [-1,1]) as float32 torch tensors."""
    import torch
    s, t = _scores(planes_tensor.detach().cpu().numpy(), actions_n)
    logits = torch.from_numpy(((s % 1024) - 512).astype(np.float32)) / 128.0
    values = torch.from_numpy(((t % 2001) - 1000).astype(np.float32)[:, None]) / 1000.0
    return logits, values


def stub_priors(planes: np.ndarray, actions_n: int):
    """(priors float32 [L,A] summing to ~1, values float32 [L])."""
    s, t = _scores(planes, actions_n)
    w = (1 + (s % 1024)).astype(np.int64)
    tot = w.sum(axis=1, keepdims=True)  # exact integer sum
    pri = w.astype(np.float32) / tot.astype(np.float32)
    val = ((t % 2001) - 1000).astype(np.float32) / np.float32(1000.0)
    return pri, val
