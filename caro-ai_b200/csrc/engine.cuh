// Device-side data layout of the self-play engine: SoA node pools with per-game arenas, a
// per-arena open-addressing transposition table, per-descent records and the compact leaf batch.
//
// Replaces the four `state_int -> list[A]` dicts of lib/mcts.py:29-36 and the Python queues of
// lib/mcts.py:259-287.  See DESIGN.md section 4 for the byte budget.
#pragma once
#include <stdint.h>
#include "rules.cuh"

namespace caro {

enum : uint8_t { KIND_SKIP = 0, KIND_TERMINAL = 1, KIND_EXPAND = 2 };
enum : uint8_t { ST_ACTIVE = 0, ST_FINISHED = 1 };
enum { CTR_LEAVES = 0, CTR_GAMES, CTR_PLIES, CTR_WIN0, CTR_WIN1, CTR_DRAW, CTR_DESCENTS, CTR_ERRORS, CTR_COUNT };
enum : unsigned long long { ERR_ARENA_FULL = 1ull, ERR_REPLAY_OVERRUN = 2ull, ERR_ILLEGAL_ACTION = 4ull };

struct alignas(16) HashSlot {
  uint64_t key;   // Key128::lo
  int32_t node;   // arena-local node index
  uint32_t gen;   // slot is live iff gen == tree_gen[tree] (clear() == ++gen, no memset)
};

struct Dims {
  int G;          // games
  int tpg;        // trees per game
  int B;          // max descents per minibatch (stride of per-descent arrays)
  int A;          // actions
  int Apad;       // row stride of N/W/Q/P (A rounded up to 8)
  int FW;         // flag words per node = ceil(A/32)
  int node_cap;   // nodes per arena
  int hash_cap;   // slots per arena (power of two, 2x node_cap)
  int max_depth;  // path stride = max plies of the game
  int max_plies;  // history stride
  int replay_cap;
};

struct SearchParams {
  double c_puct, alpha, explore;
  uint32_t seed_lo, seed_hi;
};

template <class Board>
struct View {
  // ---- tree arenas (index = tree * node_cap + node) --------------------------------------
  int32_t* N;          // [trees*node_cap][Apad]  visit counts            (lib/mcts.py:30)
  float* W;            // [..][Apad]              total value             (lib/mcts.py:32)
  float* Q;            // [..][Apad]              mean value, f32(W/N)    (lib/mcts.py:34)
  float* P;            // [..][Apad]              priors                  (lib/mcts.py:36)
  int32_t* C;          // [..][Apad]  cached child node per edge (-1 = not linked yet): a pure cache of the
                       //             transposition lookup, so an interior step of a descent is ONE dependent access
  uint32_t* flags;     // [..][FW] bit a: W(s,a) has absorbed a float32 net value (numpy promotion state)
  uint64_t* key_hi;    // [..]  upper 64 fingerprint bits (m,n,k only)
  Board* node_board;   // [..]  position of the node (export / dict views)
  uint8_t* node_player;  // [..] side to move when the node was created
  HashSlot* ht;        // [trees][hash_cap]
  int32_t* node_count; // [trees]
  uint32_t* tree_gen;  // [trees]
  // ---- games -------------------------------------------------------------------------------
  Board* root_board;   // [G]
  uint8_t* root_player;  // [G]
  uint8_t* status;     // [G]
  int32_t* ply;        // [G] completed non-terminal plies (== `step` of lib/utils.py:69)
  int32_t* result;     // [G] last finished game: +1 player0 won, -1 player1 won, 0 draw
  uint64_t* uid;       // [G] RNG sub-stream id of the current game
  uint32_t* played;    // [G] games finished in this slot
  Board* hist_board;   // [G][max_plies]
  uint8_t* hist_player;  // [G][max_plies]
  float* hist_pi;      // [G][max_plies][A]
  // ---- per descent (index = g*B + j) -------------------------------------------------------
  uint8_t* d_kind;
  float* d_value;
  Board* d_board;
  uint8_t* d_player;
  uint64_t* d_key_lo;
  uint64_t* d_key_hi;
  int32_t* d_path_len;
  int32_t* d_path_node;    // [G*B][max_depth]
  uint8_t* d_path_action;  // [G*B][max_depth]
  int32_t* d_slot;         // compact leaf slot of an expand entry, -1 otherwise
  // ---- minibatch plan ----------------------------------------------------------------------
  int32_t* q_len;      // [G]
  uint8_t* q_order;    // [G][B] descent indices in back-up order
  // ---- compact leaf batch ------------------------------------------------------------------
  Board* leaf_board;   // [G*B]
  uint8_t* leaf_player;  // [G*B]
  int32_t* leaf_count; // [1]
  double* noise;       // [G][B][A] Dirichlet noise of the coming minibatch (Philox path)
  float* probs;        // [G*B][A]  network outputs (built-in net path)
  float* values;       // [G*B]
  // ---- replay ring -------------------------------------------------------------------------
  Board* rp_board;
  uint8_t* rp_player;
  float* rp_pi;        // [replay_cap][A]
  float* rp_z;
  unsigned long long* rp_cursor;  // [1] total entries ever written
  // ---- counters ------------------------------------------------------------------------------
  unsigned long long* ctr;  // [CTR_COUNT]
};

}  // namespace caro
