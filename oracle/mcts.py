"""Oracle (CPU, test-only) restatement of the reference's tree search.

Follows lib/mcts.py:21-313 (class MCTS) and lib/utils.py:25-108 (play_game).

Numerics are the reference's *as executed under numpy >= 2 (NEP 50)* -- see SURVEY.md A.4:
  * interior nodes: U = f32(Q) + f32(f32(P * f32(sqrt(sum N))) / (1 + N))        (float32)
  * root (noisy):   P' = f64(f32(0.75 * P)) + 0.25 * noise;  U = Q + P' * sqrt(sum N) / (1+N) (float64)
  * backup:         W <- W + v ; Q <- W / N  in whatever type W currently has (python float
                    until the first float32 net value reaches the edge, float32 afterwards).
To keep those promotions *identical* to the reference this file performs the arithmetic on the
same kinds of objects (python floats, np.float32 scalars taken from a float32 row, np.float64
noise) rather than on re-typed arrays.

Injection points (the reference has none, it uses global RNG and a torch net):
  * ``dirichlet``  -- callable(alpha_list) -> float64[A]; default np.random.dirichlet   (mcts.py:56)
  * ``evaluate``   -- method turning leaf states into (priors float32[L,A], values float32[L]);
                      default = the reference's torch path (mcts.py:212-218)
  * ``rng_choice`` in play_game -- default np.random.choice                    (utils.py:66,83)
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# config.py:26-28
C_PUCT = 1.0
ALPHA = 0.30
EXPLORE = 0.25


class OracleMCTS:
    def __init__(self, game, c_puct: float = C_PUCT, alpha: float = ALPHA, explore: float = EXPLORE,
                 dirichlet: Optional[Callable] = None):
        self.game = game
        self.c_puct = c_puct
        self.alpha = alpha
        self.explore = explore
        self.dirichlet = dirichlet if dirichlet is not None else np.random.dirichlet
        # mcts.py:29-36: four maps keyed by the state int
        self.visit_count: Dict[int, List[int]] = {}
        self.value: Dict[int, List[float]] = {}
        self.value_avg: Dict[int, List[float]] = {}
        self.probs: Dict[int, Sequence[float]] = {}
        # instrumentation for the parity tests (not in the reference)
        self.trace: Optional[List[dict]] = None

    def clear(self) -> None:  # mcts.py:39-43
        for d in (self.visit_count, self.value, self.value_avg, self.probs):
            d.clear()

    def __len__(self) -> int:  # mcts.py:45-46
        return len(self.value)

    def is_leaf(self, state: int) -> bool:  # mcts.py:150-160
        return state not in self.probs

    # -- scoring -----------------------------------------------------------------------
    def _noisy(self, priors):
        """mcts.py:48-62: fresh Dirichlet(alpha) over the whole action space, every call."""
        noise = self.dirichlet([self.alpha] * self.game.action_space)
        keep = 1 - self.explore
        return [keep * p + self.explore * z for p, z in zip(priors, noise)]

    def _ucb(self, q_row, p_row, n_row):
        """mcts.py:64-84, evaluation order ((c*P)*sqrt)/(1+N) then Q + ..."""
        root_n = math.sqrt(sum(n_row))
        return [q + self.c_puct * p * root_n / (1 + n) for q, p, n in zip(q_row, p_row, n_row)]

    # -- descent -----------------------------------------------------------------------
    def find_leaf(self, state_int: int, player: int):
        """mcts.py:97-148.  Returns (value|None, leaf_state, leaf_player, states, actions)."""
        path_states: List[int] = []
        path_actions: List[int] = []
        s, who, value = state_int, player, None
        while s in self.probs:
            path_states.append(s)
            pri = self.probs[s]
            if s == state_int:  # noise whenever the walk stands on the root state (mcts.py:131)
                pri = self._noisy(pri)
            scores = self._ucb(self.value_avg[s], pri, self.visit_count[s])
            for a in self.game.invalid_moves(s):  # mcts.py:86-95
                scores[a] = -np.inf
            a = int(np.argmax(scores))  # first maximum
            path_actions.append(a)
            s, won = self.game.move(s, a, who)
            if won:
                value = -1.0  # the player to move at the terminal state has lost (mcts.py:140-142)
            who = 1 - who
            if value is None and len(self.game.possible_moves(s)) == 0:
                value = 0.0  # draw (mcts.py:145-146)
        return value, s, who, path_states, path_actions

    # -- expansion ---------------------------------------------------------------------
    def evaluate(self, states: List[int], players: List[int], net, device: str = "cpu"):
        """mcts.py:212-218: planes -> net -> softmax(dim=1); values = column 0."""
        import torch
        import torch.nn.functional as F
        planes = self.game.states_to_training_batch(states, players)
        logits, vals = net(torch.tensor(planes).to(device))
        pri = F.softmax(logits, dim=1)
        return pri.data.cpu().numpy(), vals.data.cpu().numpy()[:, 0]

    def _new_node(self, state: int, prior_row) -> None:  # mcts.py:178-190
        a = self.game.action_space
        self.visit_count[state] = [0] * a
        self.value[state] = [0.0] * a
        self.value_avg[state] = [0.0] * a
        self.probs[state] = prior_row

    def _backup(self, value, states: Sequence[int], actions: Sequence[int]) -> None:
        """mcts.py:225-246."""
        v = -value
        for s, a in zip(reversed(states), reversed(actions)):
            self.visit_count[s][a] += 1
            self.value[s][a] += v
            self.value_avg[s][a] = self.value[s][a] / self.visit_count[s][a]
            v = -v

    def search_minibatch(self, batch_size: int, state_int: int, player: int, net, device: str = "cpu") -> None:
        """mcts.py:248-287: `batch_size` descents on a frozen tree, duplicates of a planned leaf
        dropped, one evaluation, then back-ups in queue order (terminals first)."""
        backups = []           # (value, states, actions)
        todo_states, todo_players, todo_paths = [], [], []
        seen = set()
        leaves_dbg = []
        for _ in range(batch_size):
            value, leaf, leaf_player, states, actions = self.find_leaf(state_int, player)
            if self.trace is not None:
                leaves_dbg.append((value, leaf, leaf_player, list(states), list(actions)))
            if value is not None:
                backups.append((value, states, actions))
            elif leaf not in seen:
                seen.add(leaf)
                todo_states.append(leaf)
                todo_players.append(leaf_player)
                todo_paths.append((states, actions))
        if todo_states:
            pri, vals = self.evaluate(todo_states, todo_players, net, device)
            for leaf, (states, actions), v, p in zip(todo_states, todo_paths, vals, pri):
                self._new_node(leaf, p)
                backups.append((v, states, actions))
        for v, states, actions in backups:
            self._backup(v, states, actions)
        if self.trace is not None:
            self.trace.append({"descents": leaves_dbg, "expanded": list(todo_states)})

    def search_batch(self, count: int, batch_size: int, state_int: int, player: int, net, device: str = "cpu") -> None:
        for _ in range(count):  # mcts.py:162-176
            self.search_minibatch(batch_size, state_int, player, net, device)

    def get_policy_value(self, state_int: int, tau: float = 1):
        """mcts.py:289-313."""
        counts = self.visit_count[state_int]
        if tau == 0:
            pi = [0.0] * self.game.action_space
            pi[int(np.argmax(counts))] = 1.0
        else:
            powered = [c ** (1.0 / tau) for c in counts]
            z = sum(powered)
            pi = [c / z for c in powered]
        return pi, self.value_avg[state_int]


def play_game(game, mcts_stores, replay_buffer, net1, net2, steps_before_tau_0: int,
              mcts_searches: int, mcts_batch_size: int, net1_plays_first: Optional[bool] = None,
              device: str = "cpu", rng_choice: Optional[Callable] = None, make_tree: Optional[Callable] = None,
              transcript: Optional[list] = None):
    """lib/utils.py:25-108 restated.  Returns (net1_result, step).

    Differences from the reference signature: the trailing keyword-only injection points
    (``rng_choice``, ``make_tree``, ``transcript``).  With the defaults the RNG consumption
    order is the reference's: choice(2) -> per move [dirichlet per descent] -> choice(A, p=pi).
    """
    choice = rng_choice if rng_choice is not None else np.random.choice
    if make_tree is None:
        make_tree = lambda: OracleMCTS(game)
    if mcts_stores is None:
        mcts_stores = [make_tree(), make_tree()]  # utils.py:58-59: one private tree per side
    elif isinstance(mcts_stores, OracleMCTS):
        mcts_stores = [mcts_stores, mcts_stores]
    state = game.initial_state
    nets = [net1, net2]
    if net1_plays_first is None:
        cur = int(choice(2))  # utils.py:66
    else:
        cur = 0 if net1_plays_first else 1
    step = 0
    tau = 1 if steps_before_tau_0 > 0 else 0
    history = []
    result = net1_result = None
    while result is None:
        tree = mcts_stores[cur]
        tree.search_batch(mcts_searches, mcts_batch_size, state, cur, nets[cur], device=device)
        pi, _ = tree.get_policy_value(state, tau=tau)
        history.append((state, cur, pi))
        action = int(choice(game.action_space, p=pi))  # utils.py:83
        if transcript is not None:
            transcript.append({"state": state, "player": cur, "pi": list(pi), "action": action,
                               "n": list(tree.visit_count[state])})
        if action not in game.possible_moves(state):
            print("Impossible action selected")  # utils.py:84-85
        state, won = game.move(state, action, cur)
        if won:
            result = 1
            net1_result = 1 if cur == 0 else -1
            break
        cur = 1 - cur
        if len(game.possible_moves(state)) == 0:  # draw (utils.py:92-96)
            result = 0
            net1_result = 0
            break
        step += 1
        if step >= steps_before_tau_0:
            tau = 0
    if replay_buffer is not None:  # utils.py:101-106: z alternates back from the last mover
        z = result
        for st, who, pi in reversed(history):
            replay_buffer.append((st, who, pi, z))
            z = -z
    return net1_result, step
