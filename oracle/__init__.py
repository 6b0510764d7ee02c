"""CPU oracle for the caro-ai MCTS self-play hot path.

TEST INFRASTRUCTURE ONLY.  This package is a from-scratch CPU restatement of the
reference's algorithm (nh273/caro-ai: lib/game/**, lib/mcts.py, lib/model.py,
lib/utils.py:play_game).  It exists to check the CUDA engine, never to serve it:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from here.  Nothing under
``caro-ai_b200/`` imports it, and the product path raises when the CUDA library
is missing instead of falling back to this code.

Parity status: PINNED.  Every function here is checked
  * against the reference's own golden vectors (lib/**/test_*.py, restated in
    tests/test_oracle_golden.py), and
  * against fixtures produced by running the unmodified reference in the build
    container (tests/golden/*.json, generator: tests/golden/make_golden.py).

Numerics contract: "the reference as it executes under numpy 2.x / torch 2.x"
(NEP-50 promotion: interior PUCT scores and W/Q are float32, the root's noisy
scores are float64) -- see DESIGN.md section 3.
"""
