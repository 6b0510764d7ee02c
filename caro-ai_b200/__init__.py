"""B200-native AlphaZero self-play engine behind the nh273/caro-ai MCTS / game / Net API.

Host side (Python, mirrors the reference's lib/ modules) over a C-ABI CUDA library
(``libcaro_b200.so``, see include/caro_b200.h).  There is no CPU fallback: anything that computes
raises ``CaroError`` when the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
