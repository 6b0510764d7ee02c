"""Oracle (CPU, test-only) restatement of the reference's policy/value network.

Follows lib/model.py:10-94.  The ``state_dict`` key set and tensor shapes are those of the
reference's checkpoints (saves/*/best_*.dat) so that a ``.dat`` written by
either side loads in the other:  conv_in.{0,1}, conv_1..conv_5.{0,1}, conv_val.{0,1},
value.{0,2}, conv_policy.{0,1}, policy.0.

fp32, plain torch ops, no fusion: this is the thing the CUDA tower is compared against
(eval-mode BatchNorm: see SURVEY.md section 0 quirk 5 for why parity is defined on ``.eval()``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

NUM_FILTERS = 64  # model.py:7


def _conv_bn_act(cin: int, cout: int, ksize: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=ksize, padding=ksize // 2),
                         nn.BatchNorm2d(cout), nn.LeakyReLU())


class OracleNet(nn.Module):
    def __init__(self, input_shape, actions_n: int):
        super().__init__()
        planes, height, width = input_shape
        self.conv_in = _conv_bn_act(planes, NUM_FILTERS, 3)            # model.py:14-18
        for i in range(1, 6):                                           # model.py:21-45
            setattr(self, "conv_%d" % i, _conv_bn_act(NUM_FILTERS, NUM_FILTERS, 3))
        self.conv_val = _conv_bn_act(NUM_FILTERS, 1, 1)                 # model.py:50-54
        # model.py:55,74-76 sizes the head by pushing zeros through conv_val in TRAIN mode, which
        # leaves a first BatchNorm running-stat update behind; reproduced so a fresh net matches.
        probe = torch.zeros(1, NUM_FILTERS, height, width)
        self.conv_val(probe)
        self.value = nn.Sequential(nn.Linear(height * width, 20), nn.LeakyReLU(),
                                   nn.Linear(20, 1), nn.Tanh())        # model.py:56-61
        self.conv_policy = _conv_bn_act(NUM_FILTERS, 2, 1)              # model.py:64-68
        self.conv_policy(probe)                                         # model.py:69,78-80 (same probe)
        self.policy = nn.Sequential(nn.Linear(2 * height * width, actions_n))  # model.py:70-72

    def forward(self, x):
        # model.py:82-94: v <- v + block(v), no activation after the add
        b = x.size(0)
        v = self.conv_in(x)
        for i in range(1, 6):
            v = v + getattr(self, "conv_%d" % i)(v)
        val = self.value(self.conv_val(v).view(b, -1))
        pol = self.policy(self.conv_policy(v).view(b, -1))
        return pol, val
