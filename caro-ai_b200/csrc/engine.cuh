// Device-side data layout of the self-play engine: node records with per-game arenas (structure of arrays inside
// a node: [N | W | P | child link] rows, one 128-byte line per Connect4 node), a per-arena open-addressing
// transposition table, per-descent records and the compact leaf batch.
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
constexpr int kNetBit = (int)0x80000000u;   // flag bit inside an N word
constexpr int kCountMask = 0x7fffffff;
enum : unsigned long long { ERR_ARENA_FULL = 1ull, ERR_REPLAY_OVERRUN = 2ull, ERR_ILLEGAL_ACTION = 4ull };

struct alignas(16) HashSlot {
  uint64_t key;   // Key128::lo
  int32_t node;   // arena-local node index
  uint32_t gen;   // slot is live iff gen == tree_gen[tree] (clear() == ++gen, no memset)
};

// One descent of the current minibatch (lib/mcts.py:97-148 returns value, leaf state, leaf player, states[], actions[]).
// ONE record per descent -- 64 bytes for Connect4 (two sectors of one line), 96 for m,n,k -- instead of eight parallel
// arrays: select writes it with four 16-byte stores, plan and expand+backup read what they need with one or two loads
// (the parallel arrays cost ~10 sectors per descent for <= 40 useful bytes: 4.8x the algorithmic DRAM traffic).
struct alignas(16) DescHead {
  uint8_t kind;      // KIND_*
  uint8_t player;    // side to move at the leaf
  uint16_t len;      // path length (edges from the root to the leaf)
  int32_t slot;      // compact leaf slot of an expand entry (plan), -1 otherwise
  float value;       // terminal value (-1 / 0), lib/mcts.py:140-146
  uint32_t pad_;
  uint64_t key_lo, key_hi;  // transposition key of the leaf
};
template <class Board>
struct alignas(32) DescRec {
  DescHead h;
  Board board;       // leaf position
};

// One entry of a game's back-up queue in back-up order (lib/mcts.py:269-278): everything expand+backup needs to start
// its loads at once (no order -> kind -> slot -> value chain).  `slot` holds the float bits of the value for a terminal.
struct alignas(8) QEntry {
  uint8_t j;         // descent index within the minibatch
  uint8_t kind;
  uint16_t len;
  int32_t slot;
};

struct Dims {
  int G;          // games
  int tpg;        // trees per game
  int B;          // max descents per minibatch (stride of per-descent arrays)
  int A;          // actions
  int Apad;       // entries per row of a node record (A rounded up to 8)
  int RS;         // 32-bit words per node record = 4 * Apad: [N | W | P | C] rows of Apad entries each
  int node_cap;   // nodes per arena
  int hash_cap;   // slots per arena (power of two, 2x node_cap)
  int max_depth;  // path stride = max plies of the game
  int max_plies;  // history stride
  int replay_cap;
  unsigned flags;  // CARO_FLAG_* (include/caro_b200.h): throughput-mode extensions, 0 = the reference's search
};
enum : unsigned { FLAG_VIRTUAL_LOSS = 1u, FLAG_MASK_PRIORS = 2u, FLAG_FRESH_TREE = 4u, FLAG_RECYCLE_TREE = 8u, FLAG_COMPACT_TREE = 16u };

struct SearchParams {
  double c_puct, alpha, explore;
  uint32_t seed_lo, seed_hi;
};

template <class Board>
struct View {
  // ---- tree arenas (index = tree * node_cap + node) --------------------------------------
  // One record of RS = 4 * Apad words per node, rows [N | W | P | C]; the four pointers below address row 0 / 1 / 2 / 3
  // of record 0, so X[(nb + node) * RS + a] is entry a of row X.  A Connect4 record is one aligned 128-byte line: a
  // descent step reads exactly that line, a back-up touches its first 64 bytes, an expansion writes it whole.
  int32_t* N;          // visit counts (lib/mcts.py:30) in bits 0..30; bit 31 (kNetBit): W(s,a) has absorbed a float32
                       // network value (numpy promotion state of the reference's python-float / np.float32 sums)
  float* W;            // total value (lib/mcts.py:32).  The mean value Q (lib/mcts.py:34) is not stored: it always
                       // equals f32(W / N) (0 while N == 0) and is recomputed where it is read
  float* P;            // priors (lib/mcts.py:36)
  int32_t* C;          // cached child node per edge (-1 = not linked yet): a pure cache of the
                       // transposition lookup, so an interior step of a descent is ONE dependent access
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
  DescRec<Board>* desc;    // [G*B]
  uint32_t* d_path;        // [G*B][max_depth]  (arena-local node index << 8) | action, root first
  // ---- minibatch plan ----------------------------------------------------------------------
  int32_t* q_len;      // [G]
  QEntry* q_entry;     // [G][B] back-up queue
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
