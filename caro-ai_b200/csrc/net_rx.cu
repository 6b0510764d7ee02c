// Split-precision row-tiled policy/value tower (tcgen05 + TMEM, sm_100a only): the accuracy mode of net_rt.cu.
//
// lib/model.py:82-94 (eval-mode BatchNorm folded on the host) for boards with H <= 6, W <= 7, for networks the
// one-pass bf16 tower cannot evaluate within the 1e-3 contract of lib/mcts.py:212-218 -- trained checkpoints whose
// policy logits span +-100 need ~17 mantissa bits on BOTH convolution operands.
//
// Arithmetic: activations and weights are fp16 hi + fp16 lo pairs (22 mantissa bits) and every product is evaluated as
//   hi(a) hi(w) + lo(a) hi(w) + hi(a) lo(w)            (fp32 accumulation in TMEM, the lo lo term is below 2^-22),
// three MMAs per product on the same operand layout, tile mapping and stacked-tap trick as net_rt.cu (M-tile = one
// board row of 16 boards, B = [w(dy=+1) | w(dy=0) | w(dy=-1)], 128 x 192 x 16 MMAs into three adjacent accumulators).
// The residual stream v <- v + lrelu(conv(v)) IS the hi + lo pair in shared memory (no TMEM copy, no e5m2 tail).
// Measured against PyTorch fp32 on the shipped Connect4 checkpoint: priors 1.9e-4, values 1.7e-4 (tools/net_check.py).
//
// What differs from net_rt.cu is the schedule.  Two activation images (hi, lo: 2 x 98,816 B) leave 33 KB of the SM's
// shared memory, so a layer's weights (12 blocks x {hi, lo} x 6 KB = 144 KB) cannot be resident: they STREAM through a
// ring of five 6 KB slots (the head features and the FC scratch live in global memory to make room), and the MMA order
// is block-major -- for every (horizontal tap, k-step) block all H tiles issue their MMAs while the block is in the ring
// (hi block: a_hi w_hi and a_lo w_hi for every tile, then lo block: a_hi w_lo), so every weight byte is fetched from L2
// once per group and layer.  All H accumulators are live for the whole layer; the epilogue of tile y starts when the last
// block has passed tile y + 1, the next layer's first blocks chase the epilogues through the tiles: the first two and the
// last two blocks of a layer are issued tile-major for that reason (one and one with fewer than five slots or in conv_in).
// The kernel is templated on the board height: with run-time H the issuing thread, not the tensor pipe, set the pace
// (a block of 18 MMAs took ~2,300 cycles against a tensor floor of 1,536 whatever the issue order, the ring depth or the
// number of commits; with H = 6 folded into the code it takes 1,544).  Measured (tools/net_bench.py, tools/rx_probe.py,
// 9,472 leaves): 0.300 ms = 2.3 x the bf16 tower (0.130 ms) and 0.46 x the tap-per-MMA split kernel (0.648 ms); 0.385 ms
// before the specialisation.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "../../include/caro_b200.h"
#include "common_host.h"
#include "net.h"
#include "rules.cuh"

#include "tc_common.cuh"
#include "rt_common.cuh"

namespace caro {

#ifndef CARO_RX_SLOTS
#define CARO_RX_SLOTS 5
#endif
constexpr int kRxSlots = CARO_RX_SLOTS;                                   // weight ring: 6 KB blocks
constexpr int kRxLayerBlocks = 24;                            // 12 (dx, k-step) blocks x {hi, lo}
constexpr int kRxInBlocks = 6;                                // conv_in: 3 dx blocks x {hi, lo}
constexpr uint32_t kRxLoUnits = kRtActBytes / 16;             // descriptor offset of the lo activation image

__host__ __device__ constexpr uint32_t rx_idesc(uint32_t n) {  // D = f32, A = B = f16 (format 0), K-major, M = 128
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct RxCfg {
  static constexpr int kEpiWarps = 16;                             // two sets of 8: set s owns the tiles with y % 2 == s
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kSetThreads = 256;
  static constexpr int kMmaWarp = kEpiWarps;
  static constexpr int kLoadWarp = kEpiWarps + 1;
  static constexpr int kHeadWarp = kEpiWarps + 2;
  static constexpr int kHeadWarps = 2;
  static constexpr int kHeadThreads = 32 * kHeadWarps;
  static constexpr int kThreads = kEpiThreads + 64 + kHeadThreads;
  static constexpr int kActHi = 0;
  static constexpr int kActLo = kRtActBytes;
  static constexpr int kWgt = 2 * kRtActBytes;
  // the head features and the FC scratch live in global memory (caro_net::d_rt_scratch, one slot per launch in flight):
  // their 11.5 KB buy two more ring slots -- with three the MMA warp waited for weights ~8 % of the time
  static constexpr int kFcV = kWgt + kRxSlots * kRtBlockBytes;     // small head vectors (FC1 bias, FC2, policy bias)
  static constexpr int kBars = kFcV + 512;
  static constexpr int kNumBars = 2 * kRxSlots + 2 * kRtMaxH + 2;
  static constexpr int kTotal = kBars + kNumBars * 8 + 32;
  static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// Warp roles: 0..15 epilogue -- set = warp >> 3 owns the tiles with y % 2 == set, TMEM lane quarter = warp & 3, channel
// half = (warp >> 2) & 1, so two tiles are rewritten at the same time in the window between two layers where nothing
// else runs --, 16 = TMEM owner + MMA issue (one elected lane), 17 = weight producer, 18..19 = FC heads of the previous group.
// HC: the number of board rows as a compile-time constant (0 = run-time gm.H).  The MMA warp's loop is issue-bound: with H
// known every tile's position test folds away (one UTCHMMA per MMA instead of a predicated pair, no per-tile branch)
// -- 0.385 -> 0.31 ms per 9,472 Connect4 leaves.
template <class R, int HC>
__global__ void __launch_bounds__(RxCfg::kThreads, 1)
net_rx_kernel(R rules, RtGeom gm, const typename R::Board* __restrict__ boards, const uint8_t* __restrict__ who,
              const int32_t* __restrict__ d_count, long long max_count, const uint8_t* __restrict__ wimg,
              const __grid_constant__ RtConsts consts, const float* __restrict__ blob, BlobLayout L,
              const float* __restrict__ pol_fc_t, const float* __restrict__ val_fc1_t, float* __restrict__ probs,
              float* __restrict__ values, long long* __restrict__ trace, float* __restrict__ scratch) {
  using K = RxCfg;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* act = smem + K::kActHi;
  uint8_t* wgt = smem + K::kWgt;
  float* headf_s = scratch + (size_t)blockIdx.x * (kRtHeadFloats + kRtFcFloats);
  float* fc_s = headf_s + kRtHeadFloats;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + K::kBars);  // [slots] weight block landed
  uint64_t* bar_empty = bar_full + kRxSlots;                            // [slots] block consumed by every tile
  uint64_t* bar_acc = bar_empty + kRxSlots;                             // [H] all MMAs of source tiles <= y complete (last block)
  uint64_t* bar_act = bar_acc + kRtMaxH;                                // [H] activation tile rewritten + accumulator drained
  uint64_t* bar_feat = bar_act + kRtMaxH;                               // [0] head features complete, [1] consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_feat + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const long long count = d_count ? min((long long)*d_count, max_count) : max_count;
  const int nb = gm.nb;
  const int H = HC > 0 ? HC : gm.H;
  const long long n_groups = (count + nb - 1) / nb;
  if ((long long)blockIdx.x >= n_groups) return;  // uniform per CTA, before any barrier / TMEM use
  const int my_groups = (int)((n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x);

  // ---- one-time setup ---------------------------------------------------------------------
  for (int i = tid; i < 2 * kRtActBytes / 16; i += K::kThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < kRxSlots; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int t = 0; t < kRtMaxH; ++t) {
      mbar_init(bar_acc + t, 1);
      mbar_init(bar_act + t, K::kSetThreads);
    }
    mbar_init(bar_feat + 0, K::kEpiThreads);
    mbar_init(bar_feat + 1, K::kHeadThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == K::kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kRtTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= K::kHeadWarp) {
    // ===================== head warps: FC heads of group g while the tower already runs group g + 1 ==================
    const int htid = tid - K::kHeadWarp * 32;
    float* fcv = reinterpret_cast<float*>(smem + K::kFcV);
    for (int i = htid; i < 41 + gm.A; i += K::kHeadThreads)
      fcv[i] = i < 20 ? blob[L.val_fc1_b + i] : i < 40 ? blob[L.val_fc2_w + i - 20] : i == 40 ? blob[L.val_fc2_b] : blob[L.pol_fc_b + i - 41];
    asm volatile("bar.sync 2, %0;" ::"n"(K::kHeadThreads) : "memory");
    for (int gi = 0; gi + 1 < my_groups; ++gi) {
      const long long leaf0 = (blockIdx.x + (long long)gi * gridDim.x) * nb;
      const int nvalid = (int)min((long long)nb, count - leaf0);
      mbar_wait_relaxed(bar_feat + 0, (uint32_t)gi & 1u);
      rt_heads<K::kHeadThreads, 2>(gm, nvalid, leaf0, htid, headf_s, fc_s, consts.headb[0], consts.headb[1], consts.headb[2], fcv,
                                   pol_fc_t, val_fc1_t, probs, values, consts);
      mbar_arrive(bar_feat + 1);
    }
    {  // the last group of this CTA: the (idle) epilogue warps join in
      const int gi = my_groups - 1;
      const long long leaf0 = (blockIdx.x + (long long)gi * gridDim.x) * nb;
      const int nvalid = (int)min((long long)nb, count - leaf0);
      mbar_wait(bar_feat + 0, (uint32_t)gi & 1u);
      rt_heads<K::kEpiThreads + K::kHeadThreads, 3>(gm, nvalid, leaf0, K::kEpiThreads + htid, headf_s, fc_s, consts.headb[0],
                                                    consts.headb[1], consts.headb[2], fcv, pol_fc_t, val_fc1_t, probs, values, consts);
    }
  } else if (warp == K::kLoadWarp) {
    // ===================== weight producer: the blocks of the whole network, over and over, into the ring ===========
    if ((tid & 31) == 0) {
      const int blocks_net = kRxInBlocks + (gm.layers - 1) * kRxLayerBlocks;  // 126 blocks for the reference's five residual layers
      const long long total = (long long)my_groups * blocks_net;
      int slot = 0, src = 0;
      uint32_t round = 0;
      for (long long n = 0; n < total; ++n) {
        if (round > 0) mbar_wait(bar_empty + slot, (round - 1u) & 1u);
        mbar_expect_tx(bar_full + slot, (uint32_t)kRtBlockBytes);
        bulk_g2s(wgt + slot * kRtBlockBytes, wimg + (size_t)src * kRtBlockBytes, (uint32_t)kRtBlockBytes, bar_full + slot);
        if (++slot == kRxSlots) { slot = 0; ++round; }
        if (++src == blocks_net) src = 0;
      }
    }
  } else if (warp == K::kMmaWarp) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) ===========================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    const uint64_t a_desc0 = make_desc(smem_u32(act) + (uint32_t)kRtHalo * 16u, (uint32_t)kRtChunkBytes, 128u);
    const uint64_t b_desc0 = make_desc(smem_u32(wgt), 192u * 16u, 128u);
    const uint32_t full_a = smem_u32(bar_full), empty_a = smem_u32(bar_empty), acc_a = smem_u32(bar_acc), act_a = smem_u32(bar_act);
    const int n_layers = gm.layers;
    const int total_layers = my_groups * n_layers;
    int slot = 0;
    uint32_t sphase = 0;
    int layer = 0;
    // debug (tools/net_trace.py): trace[7999] == 2 -> no MMA is issued (commits only): the epilogue chain alone
    const bool dbg_skip_mma = trace != nullptr && trace[7999] == 2;
    const bool tr = elected && trace != nullptr && blockIdx.x == 0;
    // one MMA of source tile Y with the B block at `bd`: D columns out[y-1] | out[y] | out[y+1] (the first tile has no
    // out[-1]: B slot 0 is skipped; the last no out[H]).  `fresh`: the accumulators have not been written in this layer yet.
#define RX_ISSUE(Y, ad, bd, fresh)                                                                                       \
  do {                                                                                                                   \
    const uint32_t d_main_ = tmem_base + (uint32_t)((Y) == 0 ? 0 : ((Y)-1) * 64);                                        \
    if ((Y) == 0) {                                                                                                      \
      umma_f16(d_main_, ad, (bd) + 64ull, rx_idesc(128), (fresh) ? 0u : 1u);                                             \
    } else if ((Y) == H - 1) {                                                                                           \
      umma_f16(d_main_, ad, bd, rx_idesc(128), 1u); /* out[H-1] was overwritten by tile H-2 (tile 0 when H == 2) */     \
    } else if (fresh) {                                                                                                  \
      umma_f16(d_main_, ad, bd, rx_idesc(128), 1u);                                                                      \
      umma_f16(tmem_base + (uint32_t)(((Y) + 1) * 64), ad, (bd) + 128ull, rx_idesc(64), 0u);                             \
    } else {                                                                                                             \
      umma_f16(d_main_, ad, bd, rx_idesc(192), 1u);                                                                      \
    }                                                                                                                    \
  } while (0)
    // Schedule of a layer (nblk = 12 weight blocks, 3 for conv_in; block b = {hi, lo} in two ring slots):
    //   blocks 0, 1      tile-major (block 0 only in conv_in / with a short ring): tile y issues all its MMAs as soon as the previous
    //                    layer's epilogues of tiles y-1 .. y+1 are done (rows rewritten, accumulators drained) -- it chases the
    //                    epilogues through the tiles;
    //   middle blocks    block-major: the hi part (a_hi w_hi, a_lo w_hi for every tile) releases the hi slot before the lo
    //                    part (a_hi w_lo) runs, so that the ring always has the next blocks in flight;
    //   blocks n-2, n-1  tile-major again, with a commit per tile: the epilogue of tile y starts while tiles y+2 .. are issued.
    for (int gl = 0; gl < total_layers; ++gl) {
      const uint32_t par = (uint32_t)gl & 1u;
      const bool first = layer == 0;
      const int nblk = first ? 3 : 12;
#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        const int dx = first ? b - 1 : (b >> 2) - 1, kk = first ? 0 : (b & 3);
        const uint64_t a_blk = a_desc0 + (uint64_t)(int64_t)(dx + kk * 2 * kRtActRows);
        if (kRxSlots >= 5 && !first && (b == 0 || b == nblk - 2)) {
          // ---- TWO blocks tile-major (four ring slots): at the head of a layer tile y issues the MMAs of blocks 0 and 1 as soon
          // as the previous layer's epilogues of tiles y-1 .. y+1 are done, at the tail blocks n-2 and n-1 finish tile by tile with a
          // commit each -- twice the MMA work that can run underneath the epilogue chain between two layers
          const bool head = b == 0;
          const int b1 = b + 1;
          const uint64_t a_blk1 = a_desc0 + (uint64_t)(int64_t)(((b1 >> 2) - 1) + (b1 & 3) * 2 * kRtActRows);
          int sl[4];
          uint32_t sp[4];
          uint64_t bd[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sl[k] = slot + k;
            sp[k] = sphase;
            if (sl[k] >= kRxSlots) { sl[k] -= kRxSlots; sp[k] ^= 1u; }
            bd[k] = b_desc0 + (uint64_t)(uint32_t)(sl[k] * kRtBlockUnits);
          }
          mbar_wait_a(full_a + 8u * (uint32_t)sl[0], sp[0]);
          if (tr && gl * 32 + b * 2 < 1000) trace[gl * 32 + b * 2] = clock64();
#pragma unroll
          for (int y = 0; y < kRtMaxH; ++y) {
            if (y < H) {
              if (head) {
                if (y == 0) mbar_wait_a(act_a, par);
                if (y + 1 < H) mbar_wait_a(act_a + 8u * (uint32_t)(y + 1), par);
              }
              if (y == 0) mbar_wait_a(full_a + 8u * (uint32_t)sl[1], sp[1]);  // the later slots are waited for at their first use
              tc_fence_after();
              const uint64_t ad0 = a_blk + (uint64_t)(uint32_t)(y * 128), ad1 = a_blk1 + (uint64_t)(uint32_t)(y * 128);
              if (elected && !dbg_skip_mma) {
                RX_ISSUE(y, ad0, bd[0], head);
                RX_ISSUE(y, ad0 + (uint64_t)kRxLoUnits, bd[0], false);
                RX_ISSUE(y, ad0, bd[1], false);
              }
              if (y == 0) {
                mbar_wait_a(full_a + 8u * (uint32_t)sl[2], sp[2]);
                mbar_wait_a(full_a + 8u * (uint32_t)sl[3], sp[3]);
                tc_fence_after();
              }
              if (elected) {
                if (!dbg_skip_mma) {
                  RX_ISSUE(y, ad1, bd[2], false);
                  RX_ISSUE(y, ad1 + (uint64_t)kRxLoUnits, bd[2], false);
                  RX_ISSUE(y, ad1, bd[3], false);
                }
                if (!head) umma_commit_a(acc_a + 8u * (uint32_t)y);
              }
            }
          }
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_commit_a(empty_a + 8u * (uint32_t)sl[k]);
          }
          __syncwarp();
          for (int k = 0; k < 4; ++k)
            if (++slot == kRxSlots) { slot = 0; sphase ^= 1u; }
          ++b;
          continue;
        }
        const int slot_lo = slot + 1 == kRxSlots ? 0 : slot + 1;
        const uint32_t sphase_lo = slot + 1 == kRxSlots ? sphase ^ 1u : sphase;
        const uint64_t bd_hi = b_desc0 + (uint64_t)(uint32_t)(slot * kRtBlockUnits);
        const uint64_t bd_lo = b_desc0 + (uint64_t)(uint32_t)(slot_lo * kRtBlockUnits);
        mbar_wait_a(full_a + 8u * (uint32_t)slot, sphase);
        if (tr && gl * 32 + b * 2 < 1000) trace[gl * 32 + b * 2] = clock64();
        if (b == 0 || b == nblk - 1) {
          // ---- tile-major block
          mbar_wait_a(full_a + 8u * (uint32_t)slot_lo, sphase_lo);
          const bool last_blk = b == nblk - 1;
#pragma unroll
          for (int y = 0; y < kRtMaxH; ++y) {
            if (y < H) {
              if (b == 0) {
                if (y == 0) mbar_wait_a(act_a, par);
                if (y + 1 < H) mbar_wait_a(act_a + 8u * (uint32_t)(y + 1), par);
              }
              tc_fence_after();
              if (elected) {
                const uint64_t ad = a_blk + (uint64_t)(uint32_t)(y * 128);
                if (!dbg_skip_mma) {
                  RX_ISSUE(y, ad, bd_hi, b == 0);
                  if (!first) RX_ISSUE(y, ad + (uint64_t)kRxLoUnits, bd_hi, false);
                  RX_ISSUE(y, ad, bd_lo, false);
                }
                if (last_blk) umma_commit_a(acc_a + 8u * (uint32_t)y);
              }
            }
          }
          if (elected) {
            umma_commit_a(empty_a + 8u * (uint32_t)slot);
            umma_commit_a(empty_a + 8u * (uint32_t)slot_lo);
          }
        } else {
          // ---- block-major block: hi part
          tc_fence_after();
          if (elected) {
            if (!dbg_skip_mma) {
#pragma unroll
              for (int y = 0; y < kRtMaxH; ++y) {
                if (y < H) {
                  const uint64_t ad = a_blk + (uint64_t)(uint32_t)(y * 128);
                  RX_ISSUE(y, ad, bd_hi, false);
                  if (!first) RX_ISSUE(y, ad + (uint64_t)kRxLoUnits, bd_hi, false);
                }
              }
            }
            umma_commit_a(empty_a + 8u * (uint32_t)slot);
          }
          // lo part
          mbar_wait_a(full_a + 8u * (uint32_t)slot_lo, sphase_lo);
          tc_fence_after();
          if (tr && gl * 32 + b * 2 + 1 < 1000) trace[gl * 32 + b * 2 + 1] = clock64();
          if (elected) {
            if (!dbg_skip_mma) {
#pragma unroll
              for (int y = 0; y < kRtMaxH; ++y) {
                if (y < H) RX_ISSUE(y, a_blk + (uint64_t)(uint32_t)(y * 128), bd_lo, false);
              }
            }
            umma_commit_a(empty_a + 8u * (uint32_t)slot_lo);
          }
        }
        __syncwarp();
        for (int k = 0; k < 2; ++k)
          if (++slot == kRxSlots) { slot = 0; sphase ^= 1u; }
      }
      if (++layer == n_layers) layer = 0;
    }
#undef RX_ISSUE
  } else {
    // ========================================= epilogue warps =========================================
    const int quarter = warp & 3, cp = (warp >> 2) & 1, set = warp >> 3;
    const int row = quarter * 32 + (tid & 31);              // lane of the tile = TMEM lane
    const int bidx = row >> gm.pshift, col = row & (gm.pitch - 1);
    const bool real = col < gm.W;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int HW = gm.H * gm.W;
    const bool dbg_skip_epilogue = trace != nullptr && trace[7999] == 1;

    auto write_inputs = [&](int y, long long leaf0, int writer) {
      if (cp == writer) {
        uint32_t lo = 0u;
        const long long leaf = leaf0 + bidx;
        if (real && leaf < count) {
          const typename R::Board s = boards[leaf];
          const int wm = who[leaf];
          const uint32_t mine = rules.plane_value(s, wm, 0, y, col) ? 0x3C00u : 0u;  // fp16(1.0)
          const uint32_t other = rules.plane_value(s, wm, 1, y, col) ? 0x3C00u : 0u;
          lo = mine | (other << 16);
        }
        uint8_t* dst = act + (size_t)(kRtHalo + y * 128 + row) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(lo, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + kRtChunkBytes) = make_uint4(0u, 0u, 0u, 0u);  // K is padded to 16
        fence_async_smem();
      }
      mbar_arrive(bar_act + y);
    };

    // accumulator + bias, LeakyReLU and (HAS_RES) the residual hi + lo of 32 channels [c0, c0 + 32) of tile y -> v[16] (pairs)
    auto load_values = [&](auto has_res_c, int layer, int y, int c0, float2* v, const uint8_t* arow) {
      constexpr bool HAS_RES = decltype(has_res_c)::value;
      const uint32_t a_acc = tmem_base + lane_base + (uint32_t)(y * 64 + c0);
      const float4* bl4 = reinterpret_cast<const float4*>(consts.bias + layer * 64 + c0);
      uint32_t ra[32];
      uint4 hv[4], lv[4];
      TMEM_LD16(a_acc, ra);
      TMEM_LD16(a_acc + 16u, (ra + 16));
      if (HAS_RES) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          hv[c8] = *reinterpret_cast<const uint4*>(arow + (size_t)c8 * kRtChunkBytes);
          lv[c8] = *reinterpret_cast<const uint4*>(arow + (size_t)c8 * kRtChunkBytes + kRtActBytes);
        }
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float2 slope = make_float2(kLeaky, kLeaky);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bq = bl4[q];
        const float2 x01 = __fadd2_rn(make_float2(__uint_as_float(ra[q * 4]), __uint_as_float(ra[q * 4 + 1])), make_float2(bq.x, bq.y));
        const float2 x23 = __fadd2_rn(make_float2(__uint_as_float(ra[q * 4 + 2]), __uint_as_float(ra[q * 4 + 3])), make_float2(bq.z, bq.w));
        const float2 t01 = __fmul2_rn(x01, slope), t23 = __fmul2_rn(x23, slope);
        float2 m01 = make_float2(fmaxf(x01.x, t01.x), fmaxf(x01.y, t01.y));
        float2 m23 = make_float2(fmaxf(x23.x, t23.x), fmaxf(x23.y, t23.y));
        if (HAS_RES) {
          const uint4 h4 = hv[q >> 1], l4 = lv[q >> 1];
          const uint32_t h0 = (q & 1) ? h4.z : h4.x, h1 = (q & 1) ? h4.w : h4.y;
          const uint32_t l0 = (q & 1) ? l4.z : l4.x, l1 = (q & 1) ? l4.w : l4.y;
          m01 = __fadd2_rn(m01, __fadd2_rn(h2_to_f2(h0), h2_to_f2(l0)));
          m23 = __fadd2_rn(m23, __fadd2_rn(h2_to_f2(h1), h2_to_f2(l1)));
        }
        v[q * 2] = m01;
        v[q * 2 + 1] = m23;
      }
    };

    // one tile of one layer but the last: the warp's 32 channels of the tile's 128 rows, rewritten in place as hi + lo
    auto epilogue_tile = [&](auto has_res_c, int layer, int gl, int y) {
      mbar_wait(bar_acc + min(y + 1, H - 1), (uint32_t)gl & 1u);
      __syncwarp();
      tc_fence_after();
      if ((tid & 255) == 0) TC_TRACE(2, gl * 8 + y);
      if (dbg_skip_epilogue) {  // debug: the MMA stream without the epilogue's work
        tc_fence_before();
        mbar_arrive(bar_act + y);
        return;
      }
      uint8_t* arow = act + (size_t)(cp * 4 * kRtActRows + kRtHalo + y * 128 + row) * 16;
      float2 v[16];
      load_values(has_res_c, layer, y, cp * 32, v, arow);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 vv = v[c8 * 4 + j];
          const __half2 h = __floats2half2_rn(vv.x, vv.y);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(vv.x - hf.x, vv.y - hf.y);
          ph[j] = real ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
          pl[j] = real ? *reinterpret_cast<const uint32_t*>(&l) : 0u;
        }
        *reinterpret_cast<uint4*>(arow + (size_t)c8 * kRtChunkBytes) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(arow + (size_t)c8 * kRtChunkBytes + kRtActBytes) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_act + y);
      if ((tid & 255) == 0) TC_TRACE(3, gl * 8 + y);
    };

    // one tile of the last layer: its output only feeds the 1x1 head convolutions.  The cp == 0 warps of the owning set take
    // all 64 channels of their rows (one plain store per feature slot, a fixed summation order) and then write the next
    // group's input planes into the tile they have just read; the cp == 1 warps only pass the barrier on.
    auto last_tile = [&](int gl, int y, bool more, long long next_leaf0) {
      mbar_wait(bar_acc + min(y + 1, H - 1), (uint32_t)gl & 1u);
      __syncwarp();
      tc_fence_after();
      if ((tid & 255) == 0) TC_TRACE(2, gl * 8 + y);
      if (!dbg_skip_epilogue && cp == 0) {
        float av = 0.0f, ap0 = 0.0f, ap1 = 0.0f;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const uint8_t* arow = act + (size_t)(hh * 4 * kRtActRows + kRtHalo + y * 128 + row) * 16;
          float2 v[16];
          load_values(std::true_type{}, gm.layers - 1, y, hh * 32, v, arow);
          const float4* hw4 = reinterpret_cast<const float4*>(consts.headw + hh * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w0 = hw4[q], w1 = hw4[16 + q], w2 = hw4[32 + q];
            const float2 a = v[q * 2], b = v[q * 2 + 1];
            av = fmaf(a.x, w0.x, fmaf(a.y, w0.y, fmaf(b.x, w0.z, fmaf(b.y, w0.w, av))));
            ap0 = fmaf(a.x, w1.x, fmaf(a.y, w1.y, fmaf(b.x, w1.z, fmaf(b.y, w1.w, ap0))));
            ap1 = fmaf(a.x, w2.x, fmaf(a.y, w2.y, fmaf(b.x, w2.z, fmaf(b.y, w2.w, ap1))));
          }
        }
        if (real) {
          const int cell = y * gm.W + col;
          headf_s[(bidx * 3 + 0) * HW + cell] = av;
          headf_s[(bidx * 3 + 1) * HW + cell] = ap0;
          headf_s[(bidx * 3 + 2) * HW + cell] = ap1;
        }
      }
      tc_fence_before();
      if (more) write_inputs(y, next_leaf0, 0);
      if ((tid & 255) == 0) TC_TRACE(3, gl * 8 + y);
    };

    for (int y = set; y < H; y += 2) write_inputs(y, (long long)blockIdx.x * nb, 0);
    for (int gi = 0; gi < my_groups; ++gi) {
      const long long leaf0 = (blockIdx.x + (long long)gi * gridDim.x) * nb;
      const bool more = gi + 1 < my_groups;
      const long long next_leaf0 = (blockIdx.x + (long long)(gi + 1) * gridDim.x) * nb;
      const int gl0 = gi * gm.layers;
#pragma unroll 1
      for (int y = set; y < H; y += 2) epilogue_tile(std::false_type{}, 0, gl0, y);
#pragma unroll 1
      for (int layer = 1; layer < gm.layers - 1; ++layer) {
#pragma unroll 1
        for (int y = set; y < H; y += 2) epilogue_tile(std::true_type{}, layer, gl0 + layer, y);
      }
      if (gi > 0) mbar_wait(bar_feat + 1, (uint32_t)(gi - 1) & 1u);  // the previous group's features have been consumed
#pragma unroll 1
      for (int y = set; y < H; y += 2) last_tile(gl0 + gm.layers - 1, y, more, next_leaf0);
      mbar_arrive(bar_feat + 0);
      if (!more) {  // join the head warps for the heads of the last group
        mbar_wait(bar_feat + 0, (uint32_t)gi & 1u);
        const int nvalid = (int)min((long long)nb, count - leaf0);
        const float* fcv = reinterpret_cast<const float*>(smem + K::kFcV);
        rt_heads<K::kEpiThreads + K::kHeadThreads, 3>(gm, nvalid, leaf0, tid, headf_s, fc_s, consts.headb[0], consts.headb[1],
                                                      consts.headb[2], fcv, pol_fc_t, val_fc1_t, probs, values, consts);
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == K::kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kRtTmemCols) : "memory");
  }
}

}  // namespace caro

using namespace caro;

// Weight image: 6 + 24 x blocks (126 for the reference's depth) blocks of 6,144 B in consumption order -- conv_in: hi(dx=-1), lo(dx=-1), hi(0), lo(0), hi(+1), lo(+1);
// every residual layer: for (dx, k-step) in order: hi block, lo block.  A block is the B operand [2 k-chunks][n = 192][8 in-channels]
// of net_rt.cu in fp16: n = 64 j + out-channel with j = 0, 1, 2 <-> vertical tap ky = 2, 1, 0.
int caro_net_rx_pack(caro_net* net, const float* h) {
  const BlobLayout& L = net->layout;
  const int blocks = L.blocks;
  const size_t img_bytes = (size_t)(kRxInBlocks + blocks * kRxLayerBlocks) * kRtBlockBytes;
  std::vector<uint16_t> img(img_bytes / 2, 0);
  auto put = [&](int block, int j, int co, int c, float w) {  // `block` = index of the hi block; its lo block follows
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const size_t off = (size_t)block * kRtBlockBytes + (size_t)(c / 8) * 3072 + (size_t)(j * 64 + co) * 16 + (size_t)(c % 8) * 2;
    memcpy(&img[off / 2], &hi, 2);
    memcpy(&img[(off + kRtBlockBytes) / 2], &lo, 2);
  };
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 3; ++j)
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 2; ++ci) put(2 * kx, j, co, ci, h[L.conv_in_w + ((size_t)(co * 2 + ci) * 9 + (2 - j) * 3 + kx)]);
  for (int l = 0; l < blocks; ++l)
    for (int kx = 0; kx < 3; ++kx)
      for (int kk = 0; kk < 4; ++kk)
        for (int j = 0; j < 3; ++j)
          for (int co = 0; co < 64; ++co)
            for (int c = 0; c < 16; ++c)
              put(kRxInBlocks + l * kRxLayerBlocks + 2 * (kx * 4 + kk), j, co, c,
                  h[L.conv_w[l] + ((size_t)(co * 64 + kk * 16 + c) * 9 + (2 - j) * 3 + kx)]);
  cudaError_t ce = cudaSuccess;
  if (!net->d_rx_weights) ce = cudaMalloc(&net->d_rx_weights, img_bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_rx_weights, img.data(), img_bytes, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

void caro_net_rx_free(caro_net* net) {
  if (net->d_rx_weights) cudaFree(net->d_rx_weights);
  net->d_rx_weights = nullptr;
}

template <class R>
static int launch_rx(const R& rules, caro_net* net, const void* boards, const uint8_t* who, const int32_t* d_count,
                     int64_t max_count, float* probs, float* values, cudaStream_t st) {
  RtGeom gm;
  gm.H = net->H;
  gm.W = net->W;
  gm.A = net->A;
  gm.pshift = net->W < 4 ? 2 : 3;
  gm.pitch = 1 << gm.pshift;
  gm.nb = 128 / gm.pitch;
  gm.layers = 1 + net->layout.blocks;
  static const int generic_env = getenv("CARO_RX_GENERIC") ? atoi(getenv("CARO_RX_GENERIC")) : 0;  // A/B: 1 = run-time H
  if (gm.nb * 3 * gm.H * gm.W > kRtHeadFloats || gm.nb * (20 + gm.A) > kRtFcFloats || 41 + gm.A > 128)
    return caro_fail(CARO_E_ARG, "board does not fit the row-tiled tensor-core geometry");
  const long long max_groups = (max_count + gm.nb - 1) / gm.nb;
  const int lim = net->grid_limit > 0 ? net->grid_limit : net->pipeline_limit;
  const int ctas = lim > 0 && lim < net->sm_count ? lim : net->sm_count;
  const unsigned grid = (unsigned)(max_groups < ctas ? max_groups : ctas);
  // every launch in flight has its own slot of the global head scratch (launches of different pipeline parts overlap)
  float* scratch = (float*)net->d_rt_scratch +
                   (size_t)(net->rt_scratch_seq++ % kRtScratchSlots) * net->sm_count * (kRtHeadFloats + kRtFcFloats);
  auto kern = gm.H == 6 && !generic_env ? net_rx_kernel<R, 6> : net_rx_kernel<R, 0>;
  kern<<<grid, RxCfg::kThreads, RxCfg::kTotal, st>>>(
      rules, gm, (const typename R::Board*)boards, who, d_count, (long long)max_count, (const uint8_t*)net->d_rx_weights,
      *reinterpret_cast<const RtConsts*>(net->h_rt_consts), net->d_blob, net->layout, net->d_pol_fc_t,
      net->d_pol_fc_t + (size_t)2 * net->H * net->W * net->A, probs, values, (long long*)net->d_trace, scratch);
  return caro_check_launch("net_rx_kernel");
}

int caro_net_rx_prepare() {
  cudaError_t ce = cudaFuncSetAttribute(net_rx_kernel<C4Rules, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, RxCfg::kTotal);
  if (ce == cudaSuccess) ce = cudaFuncSetAttribute(net_rx_kernel<C4Rules, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RxCfg::kTotal);
  if (ce == cudaSuccess) ce = cudaFuncSetAttribute(net_rx_kernel<MnkRules, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, RxCfg::kTotal);
  if (ce == cudaSuccess) ce = cudaFuncSetAttribute(net_rx_kernel<MnkRules, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RxCfg::kTotal);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

int caro_net_rx_forward(caro_net* net, int game, int n, int k, const void* d_boards, const uint8_t* d_who,
                        const int32_t* d_count, int64_t max_count, float* d_probs, float* d_values, cudaStream_t st) {
  if (game == CARO_GAME_CONNECT4) return launch_rx<C4Rules>(C4Rules(), net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
  return launch_rx<MnkRules>(MnkRules{n, k}, net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
}
