// TEST-ONLY host build of the __host__ __device__ game rules (caro-ai_b200/csrc/rules.cuh) so that
// the bitboard arithmetic can be checked against the oracle / golden fixtures on a machine without
// a GPU.  Not linked into libcaro_b200.so and not reachable from the product package.
#include <stdint.h>
#include "../../caro-ai_b200/csrc/rules.cuh"
#include "../../caro-ai_b200/csrc/rng.cuh"
using namespace caro;

extern "C" {

int hc_c4_apply(uint64_t* mask, uint64_t* black, int col, int player) {
  C4Board s{*mask, *black};
  const bool won = C4Rules().apply(s, col, player);
  *mask = s.mask;
  *black = s.black;
  return won ? 1 : 0;
}
int hc_c4_legal(uint64_t mask, uint64_t black) {
  C4Board s{mask, black};
  int m = 0;
  for (int c = 0; c < 7; ++c) m |= (C4Rules().legal(s, c) ? 1 : 0) << c;
  return m | ((C4Rules().any_legal(s) ? 1 : 0) << 8);
}
uint64_t hc_c4_key(uint64_t mask, uint64_t black) { return C4Rules().key(C4Board{mask, black}).lo; }
int hc_c4_plane(uint64_t mask, uint64_t black, int who, int plane, int row, int col) {
  return C4Rules().plane_value(C4Board{mask, black}, who, plane, row, col);
}
int hc_mnk_apply(int n, int k, uint64_t* words, int a, int player) {
  MnkBoard s;
  for (int i = 0; i < 4; ++i) { s.w[i] = words[i]; s.b[i] = words[4 + i]; }
  const bool won = MnkRules{n, k}.apply(s, a, player);
  for (int i = 0; i < 4; ++i) { words[i] = s.w[i]; words[4 + i] = s.b[i]; }
  return (won ? 1 : 0) | ((MnkRules{n, k}.any_legal(s) ? 1 : 0) << 1);
}
int hc_mnk_legal(int n, int k, const uint64_t* words, int a) {
  MnkBoard s;
  for (int i = 0; i < 4; ++i) { s.w[i] = words[i]; s.b[i] = words[4 + i]; }
  return MnkRules{n, k}.legal(s, a) ? 1 : 0;
}
void hc_mnk_key(int n, int k, const uint64_t* words, uint64_t* out) {
  MnkBoard s;
  for (int i = 0; i < 4; ++i) { s.w[i] = words[i]; s.b[i] = words[4 + i]; }
  const Key128 key = MnkRules{n, k}.key(s);
  out[0] = key.lo;
  out[1] = key.hi;
}
int hc_mnk_plane(int n, int k, const uint64_t* words, int who, int plane, int row, int col) {
  MnkBoard s;
  for (int i = 0; i < 4; ++i) { s.w[i] = words[i]; s.b[i] = words[4 + i]; }
  return MnkRules{n, k}.plane_value(s, who, plane, row, col);
}
float hc_gamma(float alpha, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2) {
  return gamma_small(alpha, k0, k1, c0, c1, c2);
}
void hc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  const Philox4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
  for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}
}
