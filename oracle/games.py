"""Oracle (CPU, test-only) restatement of the reference's game rules.

Follows, function by function:
  * lib/game/game.py:9-120                       -- the BaseGame contract
  * lib/game/connect_four/connect_four.py:8-281  -- ConnectFour
  * lib/game/tictactoe/tictactoe.py:10-259       -- TicTacToe(n, k)  (the m,n,k family)
  * lib/game/tictactoe/tictactoe_helpers.py:7-179

The state integers are the reference's; the code below
works on them arithmetically instead of via intermediate bit/str lists, which is the
only liberty taken.  Pinned by tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


class ConnectFourOracle:
    """Connect-4 on 6 rows x 7 columns; player 1 = black ('X'), 0 = white ('O').

    State int (connect_four.py:16-55,110-147): 42 colour bits, column-major, each column
    bottom-up, most significant first, followed by seven 3-bit free-slot counters.
    """

    game_rows = 6
    game_cols = 7
    bits_in_len = 3
    player_black = 1
    player_white = 0
    count_to_win = 4

    # -- bit positions --------------------------------------------------------------
    @classmethod
    def _cell_shift(cls, col: int, row: int) -> int:
        # connect_four.py:120-127 lays cells out MSB-first; cell (col,row) is list index 6*col+row
        return (cls.game_cols * cls.game_rows + cls.game_cols * cls.bits_in_len - 1) - (cls.game_rows * col + row)

    @classmethod
    def _free_shift(cls, col: int) -> int:
        return cls.bits_in_len * (cls.game_cols - 1 - col)

    def _height(self, state: int, col: int) -> int:
        free = (state >> self._free_shift(col)) & 7
        return max(self.game_rows - free, 0)

    def _token(self, state: int, col: int, row: int) -> Optional[int]:
        """Colour at (col,row) or None when off-board / above the column's height."""
        if not (0 <= col < self.game_cols) or row < 0 or row >= self.game_rows:
            return None
        if row >= self._height(state, col):
            return None
        return (state >> self._cell_shift(col, row)) & 1

    # -- BaseGame surface (game.py:17-120) --------------------------------------------
    @property
    def initial_state(self) -> int:
        # connect_four.py:65-72: empty columns, every counter = 6 (0b110)
        s = 0
        for c in range(self.game_cols):
            s |= self.game_rows << self._free_shift(c)
        return s

    @property
    def obs_shape(self) -> Tuple[int, int, int]:
        return (2, self.game_rows, self.game_cols)  # connect_four.py:74-81

    @property
    def action_space(self) -> int:
        return self.game_cols  # connect_four.py:83-90

    def encode_lists(self, cols: Sequence[Sequence[int]]) -> int:
        """connect_four.py:108-127 (list-of-columns -> int)."""
        assert len(cols) == self.game_cols
        s = 0
        for c, col in enumerate(cols):
            for r, tok in enumerate(col):
                s |= (tok & 1) << self._cell_shift(c, r)
            s |= (self.game_rows - len(col)) << self._free_shift(c)
        return s

    def decode_binary(self, state: int) -> List[List[int]]:
        """connect_four.py:129-147 (int -> list-of-columns, bottom-up)."""
        assert isinstance(state, int)
        return [[(state >> self._cell_shift(c, r)) & 1 for r in range(self._height(state, c))]
                for c in range(self.game_cols)]

    def possible_moves(self, state: int) -> List[int]:
        # connect_four.py:157-165: columns that are not full, ascending
        assert isinstance(state, int)
        return [c for c in range(self.game_cols) if self._height(state, c) < self.game_rows]

    def invalid_moves(self, state: int) -> List[int]:
        # connect_four.py:167-173 (set difference; order is irrelevant to callers)
        ok = set(self.possible_moves(state))
        return [c for c in range(self.game_cols) if c not in ok]

    def _run(self, state: int, col: int, row: int, dcol: int, drow: int, player: int) -> int:
        """Contiguous `player` tokens strictly beyond (col,row) in direction (dcol,drow)."""
        n = 0
        c, r = col + dcol, row + drow
        while self._token(state, c, r) == player:
            n += 1
            c += dcol
            r += drow
        return n

    def move(self, state: int, col: int, player: int) -> Tuple[int, bool]:
        """connect_four.py:241-265 + _check_won :206-239."""
        assert isinstance(state, int)
        assert isinstance(col, (int, np.integer))
        assert 0 <= col < self.game_cols
        assert player == self.player_black or player == self.player_white
        h = self._height(state, col)
        assert h < self.game_rows
        new = state | (int(player) << self._cell_shift(col, h))
        new -= 1 << self._free_shift(col)  # one free slot fewer
        # vertical: the top count_to_win tokens of the column (connect_four.py:258-259)
        won = h + 1 >= self.count_to_win and all(
            self._token(new, col, h - i) == player for i in range(self.count_to_win))
        if not won:
            # horizontal, rising, falling lines through the new token (connect_four.py:260-263)
            for drow in (0, 1, -1):
                total = 1 + self._run(new, col, h, -1, -drow, player) + self._run(new, col, h, 1, drow, player)
                if total >= self.count_to_win:
                    won = True
                    break
        return new, won

    def states_to_training_batch(self, states: Sequence[int], who_moves: Sequence[int]) -> np.ndarray:
        """connect_four.py:175-204: plane 0 = mover's tokens, plane 1 = every other token,
        row 0 is the top of the board."""
        out = np.zeros((len(states),) + self.obs_shape, dtype=np.float32)
        for i, (s, who) in enumerate(zip(states, who_moves)):
            for c in range(self.game_cols):
                for r in range(self._height(s, c)):
                    tok = (s >> self._cell_shift(c, r)) & 1
                    out[i, 0 if tok == who else 1, self.game_rows - 1 - r, c] = 1.0
        return out

    def render(self, state: int) -> str:
        # connect_four.py:267-281
        grid = [[" "] * self.game_cols for _ in range(self.game_rows)]
        for c in range(self.game_cols):
            for r in range(self._height(state, c)):
                grid[self.game_rows - 1 - r][c] = "X" if (state >> self._cell_shift(c, r)) & 1 else "O"
        body = "\n".join("".join(row) for row in grid)
        return "0123456\n-------\n" + body + "\n-------\n0123456"


class MNKOracle:
    """The reference's TicTacToe(n, k): square n x n board, k in a row wins (overlines count).

    State int (tictactoe.py:89-135): n*n decimal digits, cell 0 (top-left) most significant,
    row-major; digit 0 = white, 1 = black, 2 = empty.
    """

    player_black = 1
    player_white = 0
    empty = 2

    def __init__(self, n: int = 3, k_to_win: int = 3):
        self.board_len = n
        self.k_to_win = k_to_win

    # -- digits ---------------------------------------------------------------------
    def _digits(self, state: int) -> List[int]:
        cells = self.board_len ** 2
        text = str(state).rjust(cells, "0")  # tictactoe.py:89-100 (leading zeros restored)
        return [ord(ch) - 48 for ch in text]

    @staticmethod
    def _from_digits(digits: Sequence[int]) -> int:
        v = 0
        for d in digits:
            v = v * 10 + int(d)
        return v

    # -- BaseGame surface -------------------------------------------------------------
    @property
    def initial_state(self) -> int:
        return self._from_digits([self.empty] * self.board_len ** 2)  # tictactoe.py:56-63

    @property
    def obs_shape(self) -> Tuple[int, int, int]:
        return (2, self.board_len, self.board_len)

    @property
    def action_space(self) -> int:
        return self.board_len ** 2

    def encode_game_state(self, rows: Sequence[Sequence[int]]) -> int:
        return self._from_digits([d for row in rows for d in row])  # tictactoe.py:102-114

    def convert_mcts_state_to_list_state(self, state: int) -> List[List[int]]:
        d = self._digits(state)
        n = self.board_len
        return [d[r * n:(r + 1) * n] for r in range(n)]  # tictactoe.py:116-135

    def possible_moves(self, state: int) -> List[int]:
        return [i for i, d in enumerate(self._digits(state)) if d == self.empty]  # :137-151

    def invalid_moves(self, state: int) -> List[int]:
        return [i for i, d in enumerate(self._digits(state)) if d != self.empty]  # :153-162

    @staticmethod
    def _has_run(line: Sequence[int], k: int, token: int) -> bool:
        """tictactoe_helpers.py:25-56: does `line` hold >= k consecutive `token`s?"""
        assert k > 1
        best = cur = 0
        for v in line:
            cur = cur + 1 if v == token else 0
            best = max(best, cur)
        return best >= k

    def _lines_through(self, d: Sequence[int], row: int, col: int) -> List[List[int]]:
        """Full row, column, diagonal and anti-diagonal through (row,col)
        (tictactoe_helpers.py:59-179)."""
        n = self.board_len
        at = lambda r, c: d[r * n + c]
        lines = [[at(row, c) for c in range(n)], [at(r, col) for r in range(n)]]
        off = min(row, col)
        r, c, diag = row - off, col - off, []
        while r < n and c < n:
            diag.append(at(r, c))
            r, c = r + 1, c + 1
        lines.append(diag)
        # anti-diagonal, collected bottom-left -> top-right like get_antidiag
        off = min(n - 1 - row, col)
        r, c, anti = row + off, col - off, []
        while r >= 0 and c < n:
            anti.append(at(r, c))
            r, c = r - 1, c + 1
        lines.append(anti)
        return lines

    def move(self, state: int, move: int, player: int) -> Tuple[int, bool]:
        """tictactoe.py:210-235.  Like the reference the target cell is overwritten without
        an emptiness check; callers are responsible for legality."""
        assert player == self.player_white or player == self.player_black
        assert 0 <= move < self.action_space
        d = self._digits(state)
        row, col = divmod(int(move), self.board_len)
        d[row * self.board_len + col] = int(player)
        won = any(self._has_run(line, self.k_to_win, player) for line in self._lines_through(d, row, col))
        return self._from_digits(d), won

    def states_to_training_batch(self, states: Sequence[int], who_moves: Sequence[int]) -> np.ndarray:
        """tictactoe.py:164-208: plane 0 = mover's cells, plane 1 = opponent's (non-empty) cells."""
        n = self.board_len
        out = np.zeros((len(states),) + self.obs_shape, dtype=np.float32)
        for i, (s, who) in enumerate(zip(states, who_moves)):
            for idx, dgt in enumerate(self._digits(s)):
                if dgt == who:
                    out[i, 0, idx // n, idx % n] = 1.0
                elif dgt != self.empty:
                    out[i, 1, idx // n, idx % n] = 1.0
        return out

    def render(self, state: int) -> str:
        # tictactoe.py:237-259
        n = self.board_len
        sym = {self.player_white: "❌", self.player_black: "⭕"}
        rows = []
        for r, row in enumerate(self.convert_mcts_state_to_list_state(state)):
            rows.append("|" + "|".join(str(r * n + c) if v == self.empty else sym[v]
                                       for c, v in enumerate(row)) + "|")
        return "\n".join(rows)


def make_game(kind: str, n: int = 3, k: int = 3):
    """'connect4' | 'mnk' factory used by the tests and the CPU baseline."""
    if kind == "connect4":
        return ConnectFourOracle()
    if kind == "mnk":
        return MNKOracle(n, k)
    raise ValueError(kind)
