// Row-tiled policy/value tower on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// lib/model.py:82-94 (eval-mode BatchNorm folded on the host) for boards with H <= 6, W <= 7 (Connect4, 3x3 ..
// 6x6 m,n,k) -- the second-generation kernel; net_tc.cu keeps serving larger boards and the bf16x3 mode.
//
// What limited net_tc.cu: a 128x64x16 MMA reads 4 KB of A and 2 KB of B from shared memory, 48 cycles of the
// 128 B/clk operand path against a 32-cycle tensor floor, and 26 % of its rows were padding.  Here
//   * an M-tile is ONE BOARD ROW of 128/pitch boards (lane = board * pitch + column, pitch = 8 for Connect4),
//     so the vertical taps of a 3x3 convolution are TILE offsets, not row offsets: no zero row between boards
//     (87.5 % useful rows), and rows outside the board are simply not issued;
//   * the three vertical taps are stacked along N: source tile y multiplies its activations ONCE per
//     (horizontal tap, k-step) with B = [w(dy=+1) | w(dy=0) | w(dy=-1)] (192 columns) and the MMA accumulates
//     into the three neighbouring output tiles, which are adjacent column ranges of TMEM (out[y] at 64y):
//     12 MMAs of 128x192x16 per tile and layer instead of 36 of 128x64x16, 10 KB of operands per 96-cycle MMA
//     (in isolation these MMAs issue at the 96-cycle tensor floor with 80 operand wavefronts each, tools/cta2_probe.cu;
//     inside the tower ~103 cycles with the epilogue idle -- what the issuing warp does between two tiles is only
//     hidden while MMAs are queued, ~95 cycles next to a commit -- and ~115 with it running: the epilogue chain
//     bounds the layer, see DESIGN.md sections 4 and 8);
//   * the fp32 accumulators of all H output tiles fill 384 of the 512 TMEM columns, so the residual stream
//     v <- v + lrelu(conv(v)) is kept as bf16 hi (the shared-memory activations themselves) + an e5m2 lo part
//     (4 channels per TMEM column, 96 columns): ~11 mantissa bits, indistinguishable from the fp32 stream at
//     the 1e-3 contract;
//   * weights stream through three 36 KB regions (half a layer = six 6 KB blocks of one (horizontal tap, k-step)
//     each): layer L occupies two, the first half of L+1 is prefetched into the third; cp.async.bulk + one "full"
//     mbarrier per region by a producer warp, one "empty" mbarrier per region arrived by tcgen05.commit when the
//     layer's last tile has used it;
//   * biases and 1x1 head weights come through the constant cache (__grid_constant__ struct), the last layer
//     stores every head feature exactly once (tiles split by parity between the two channel halves): no
//     shared-memory atomics, bit-identical results from run to run.
//
// What bounds it (round 2): the EPILOGUE CHAIN -- the same eight warps rewrite the six tiles of a layer
// back to back, ~1,050 cycles of work (344 SASS instructions per thread) + ~230 cycles of barrier / fence / loop per tile =
// ~7,660 cycles per layer against a tensor floor of 6,144, with the MMA warp far ahead (bf16 form).  Three template switches:
//   HC    the board height as a compile-time constant (6): the MMA warp's tile loop unrolled, every position test folded;
//   F16   fp16 instead of bf16 operands (weights packed as fp16 too): the activations carry 11 mantissa bits by themselves and ARE
//         the residual stream -- no e5m2 tail, 40 % fewer epilogue instructions: +8 %, MMA-bound at 99 % of the sustained cuBLAS
//         rate, and 3-6x closer to fp32.  impl 7; what precision="auto" picks whenever its device check passes;
//   PAIR  two CTAs of a cluster run ONE cta_group::2 MMA stream (M = 256, each CTA stores and fetches half of every B
//         operand: 30 % fewer operand wavefronts, half the bank conflicts, bit-identical results) -- 7 % slower, because the
//         pair couples two epilogue chains and shared memory was not the limit; kept as impl 5 / CARO_RT_PAIR=1 for A/B runs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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

#ifndef CARO_RT_MAXREG
#define CARO_RT_MAXREG 96  // registers are allocated per SM sub-partition (16 K each): 14 warps x 96 registers leave
                           // >= 4 K per partition, enough for the warps of a tree-kernel block to become resident
                           // NEXT TO a tower CTA (tools/coresidency_test.py; at 112 they have to wait for it to end)
#endif
#ifndef CARO_RT_CP
#define CARO_RT_CP 2  // channel parts of the epilogue: 2 -> 8 epilogue warps of 32 channels, 4 -> 16 warps of 16 (needs
                      // -DCARO_RT_MAXREG=80; measured 2-4 % slower: the per-warp fixed cost of a tile is paid twice as often)
#endif

namespace caro {

// CP = channel parts of the epilogue (2 -> 8 epilogue warps x 32 channels, 4 -> 16 warps x 16 channels)
// PAIR: two CTAs of a cluster (one TPC) run the tower as ONE cta_group::2 MMA stream of M = 256: each CTA keeps its own 16
// boards (activations, accumulators, epilogue, heads) but stores and fetches only HALF of every B operand, 4 + 3 KB per
// N = 192 MMA instead of 4 + 6.  The blocks grow from 6 to 7 KB (three windows, rt_common.cuh), which the head features and
// the FC scratch pay for by moving to global memory.
// F16: activations AND weights in fp16 (11 mantissa bits on both MMA operands instead of bf16's 8), fp32 accumulation.  The
// residual stream is then the fp16 activations themselves -- as precise as the bf16 + e5m2 pair -- so the e5m2 tail in TMEM and
// 40 % of the epilogue's instructions go away.  What fp16 gives up is range (|x| < 65,504): the precision selection measures
// the result against the fp32 tower after every weight upload and falls back when that ever bites (model.py).
template <int CP, bool PAIR_ = false, bool F16_ = false>
struct RtCfg {
  static constexpr bool kPair = PAIR_;
  static constexpr bool kF16 = F16_;
  static_assert(!(PAIR_ && F16_), "the pair form exists for the bf16 mode only");
  static constexpr int kBlockBytes = PAIR_ ? kRtPairBlockBytes : kRtBlockBytes;
  static constexpr int kBlockUnits = kBlockBytes / 16;
  static constexpr int kRegionBytes = kRtRegionBlocks * kBlockBytes;
  static constexpr int kCP = CP;
  static constexpr int kCH = 64 / CP;
  static constexpr int kEpiWarps = 4 * CP;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kMmaWarp = kEpiWarps;
  static constexpr int kLoadWarp = kEpiWarps + 1;
  static constexpr int kHeadWarp = kEpiWarps + 2;
  static constexpr int kHeadWarps = 4;
  static constexpr int kHeadThreads = 32 * kHeadWarps;
  static constexpr int kThreads = kEpiThreads + 64 + kHeadThreads;
  static_assert(CP == 2 || CP == 4, "two or four channel parts");
  static_assert(CP == 2 || !PAIR_, "the pair form is built for two channel parts");
  static constexpr int kAct = 0;
  static constexpr int kWgt = kAct + kRtActBytes;
  static constexpr int kHeadF = kWgt + kRtRegions * kRegionBytes;  // float [nb][3][HW] head features (PAIR: in global memory)
  static constexpr int kFc = kHeadF + (PAIR_ ? 0 : kRtHeadFloats * 4);
  static constexpr int kFcW = kFc + (PAIR_ ? 0 : kRtFcFloats * 4);  // transposed FC weights (policy, value FC1) when they fit
  static constexpr int kBars0 = kFcW;
  static constexpr int kNumBars = kRtRegions * kRtRegionBlocks + 2 * kRtRegions + 2 * kRtMaxH + 2;
  static constexpr int kFixed = kBars0 + kNumBars * 8 + 32;         // everything but the FC weights
  static constexpr int kFcWFloats = (CARO_RT_SMEM_LIMIT - kFixed) / 4;  // what is left of the budget
  static constexpr int kBars = kFcW + kFcWFloats * 4;
  static constexpr int kTotal = kBars + kNumBars * 8 + 32;
  static_assert(kFixed <= CARO_RT_SMEM_LIMIT && kTotal <= CARO_RT_SMEM_LIMIT, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

template <int V>
struct RtInt {};
template <int V>
__device__ __forceinline__ constexpr int rt_value(RtInt<V>) { return V; }
__device__ __forceinline__ constexpr int rt_value(int v) { return v; }

// The MMAs of one source tile.  POS: 0 = first board row (no out[-1]: B slot 0 is skipped, the accumulators of
// out[0], out[1] are overwritten by the first MMA; it is also the tile that waits for the weight blocks to land),
// 1 = interior row (N = 192; the very first MMA is split so that out[y+1] is overwritten while out[y-1], out[y]
// accumulate), 2 = last row (no out[H]; it releases the weight regions).  Everything but the two region bases, the
// tile base and the TMEM addresses is a compile-time constant, so that the MMAs issue back to back.
// `mid` (run when do_mid, in the middle of the tile): work of the issuing warp that would otherwise sit between two layers; in the
// middle of an N = 192 tile the MMAs already queued hide ~265 cycles of it, next to a commit only ~95 (tools/cta2_probe.cu).
// PAIR (cta_group::2, issued by the leader CTA for both): every MMA accumulates -- the epilogues leave the accumulators they
// have read zeroed --, B comes from the window of the tile's position, `fullp` = the barriers the peer's weights are reported on.
template <int POS, bool FIRST, bool PAIR, bool F16, class Mid>
__device__ __forceinline__ void rt_issue_tile(bool do_mid, Mid&& mid, uint32_t elected, uint64_t a_tile, uint64_t rb0, uint64_t rb1, uint32_t d_main,
                                              uint32_t d_new, uint32_t full0, uint32_t full1, uint32_t ph0, uint32_t ph1,
                                              uint32_t empty0, uint32_t empty1, uint32_t next_bar, uint32_t next_bar2, uint32_t next_par,
                                              uint32_t fullp0, uint32_t fullp1, long long* tstamp = nullptr) {
  // all barriers are shared-memory addresses; next_bar / next_bar2 == 0: nothing to poll
  constexpr int NB = FIRST ? 3 : 12;
  constexpr int kUnits = PAIR ? kRtPairBlockBytes / 16 : kRtBlockUnits;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int dx = FIRST ? i - 1 : i / 4 - 1, kk = FIRST ? 0 : i % 4;
    if (POS == 0 && (i == 0 || i == 6)) {  // first use of the region in this layer
      if (tstamp && elected) tstamp[i < 6 ? 0 : 3] = clock64();
      mbar_wait_a(i < 6 ? full0 : full1, i < 6 ? ph0 : ph1);
      if (tstamp && elected) tstamp[i < 6 ? 1 : 4] = clock64();
      if (PAIR) mbar_wait_cluster_a(i < 6 ? fullp0 : fullp1, i < 6 ? ph0 : ph1);
      if (tstamp && elected) tstamp[i < 6 ? 2 : 5] = clock64();
    }
    // the barrier the NEXT tile needs is polled while this tile's MMAs are still queued in the tensor pipe
    if (i == (FIRST ? 1 : 4) && do_mid) mid();
    if (i == (FIRST ? 1 : 8) && next_bar != 0u) {
      if (PAIR) {
        mbar_wait_cluster_a(next_bar, next_par);
        if (next_bar2 != 0u) mbar_wait_cluster_a(next_bar2, next_par);
      } else {
        mbar_wait_a(next_bar, next_par);
        if (next_bar2 != 0u) mbar_wait_a(next_bar2, next_par);
      }
    }
    if (elected) {
      const uint64_t ad = a_tile + (uint64_t)(int64_t)(dx + kk * 2 * kRtActRows);
      if (PAIR) {
        const uint64_t bd = (i < 6 ? rb0 + (uint64_t)(i * kUnits) : rb1 + (uint64_t)((i - 6) * kUnits)) +
                            (POS == 0 ? (uint64_t)kRtPairTop : POS == 2 ? (uint64_t)kRtPairBottom : 0ull);
        umma_bf16_pair(d_main, ad, bd, POS == 1 ? rt_idesc_pair(192) : rt_idesc_pair(128), 1u);
        if (POS == 2 && (i == 5 || i == NB - 1)) umma_commit_pair_a(i < 6 ? empty0 : empty1);
      } else {
        const uint64_t bd = (i < 6 ? rb0 + (uint64_t)(i * kUnits) : rb1 + (uint64_t)((i - 6) * kUnits)) + (POS == 0 ? 64ull : 0ull);
        if (POS == 1 && i == 0) {
          umma_bf16(d_main, ad, bd, rt_idesc_fmt<F16>(128), 1u);
          umma_bf16(d_new, ad, bd + 128ull, rt_idesc_fmt<F16>(64), 0u);
        } else {
          umma_bf16(d_main, ad, bd, POS == 1 ? rt_idesc_fmt<F16>(192) : rt_idesc_fmt<F16>(128), (POS == 0 && i == 0) ? 0u : 1u);
        }
        if (POS == 2 && (i == 5 || i == NB - 1)) umma_commit_a(i < 6 ? empty0 : empty1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------- kernel
// Warp roles (CP = 4: 20 warps): 0 .. 4CP-1 epilogue (TMEM lane quarter = warp & 3, channel part = warp >> 2),
// 4CP = TMEM owner + MMA issue (one elected lane: the accumulate flags make the order of the MMAs significant, so
// there is exactly one issuing thread), 4CP+1 = weight producer, 4CP+2, 4CP+3 = FC heads of the previous group.
// Per layer `gl` (stage), tile y:
//   MMA(gl, y)  waits act[y+1] of stage gl (act[y-1], act[y] were waited for by the previous tiles): rows rewritten
//               AND the accumulators out[y-1..y+1] drained by the previous layer's epilogues;
//   EPI(gl, y)  waits the commit after source tile min(y+1, H-1): out[y] has received all of its contributions;
//               it rewrites act[y] in place (source tile y has been consumed by then).
// HC: the number of board rows as a compile-time constant (0 = run-time gm.H): the MMA warp's tile loop is then unrolled
// with every position test, barrier choice and TMEM address folded -- the issuing thread is the tower's critical resource.
template <class R, class K, int HC>
__global__ void __maxnreg__(CARO_RT_MAXREG)
net_rt_kernel(R rules, RtGeom gm, const typename R::Board* __restrict__ boards, const uint8_t* __restrict__ who,
              const int32_t* __restrict__ d_count, long long max_count, const uint8_t* __restrict__ wimg,
              const __grid_constant__ RtConsts consts, const float* __restrict__ blob, BlobLayout L,
              const float* __restrict__ pol_fc_t, const float* __restrict__ val_fc1_t, float* __restrict__ probs,
              float* __restrict__ values, long long* __restrict__ trace, float* __restrict__ scratch) {
  constexpr bool PAIR = K::kPair;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* act = smem + K::kAct;
  uint8_t* wgt = smem + K::kWgt;
  // PAIR: head features and FC scratch of this CTA live in global memory (`scratch`, one slot per launch in flight)
  float* headf_s = PAIR ? scratch + (size_t)blockIdx.x * (kRtHeadFloats + kRtFcFloats) : reinterpret_cast<float*>(smem + K::kHeadF);
  float* fc_s = PAIR ? headf_s + kRtHeadFloats : reinterpret_cast<float*>(smem + K::kFc);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + K::kBars);  // [regions][6], slot 0 used: the region's weight blocks landed
  uint64_t* bar_empty = bar_full + kRtRegions * kRtRegionBlocks;          // [regions] region consumed by the last tile
  uint64_t* bar_fullp = bar_empty + kRtRegions;                           // [regions] PAIR, leader: the peer's blocks landed too
  uint64_t* bar_acc = bar_fullp + kRtRegions;                             // [H] MMAs of source tile y complete
  uint64_t* bar_act = bar_acc + kRtMaxH;                                  // [H] activation tile rewritten + accumulator drained
  uint64_t* bar_feat = bar_act + kRtMaxH;                                 // [0] head features complete, [1] consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_feat + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) TC_TRACE(7, 0);  // kernel entered
  // debug (tools/pipeline_trace.py, trace[7998] == 2): global-timer entry / exit stamps of the first and the last CTA of
  // every launch, to see how consecutive launches of the self-play pipeline overlap on the GPU
  long long* gt_slot = nullptr;
  if (trace != nullptr && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && trace[7998] == 2) {
    long long* base = trace + 8000 + (blockIdx.x == 0 ? 0 : 4096);
    const long long k = base[0];
    base[0] = k + 1;
    if (k < 2040) {
      gt_slot = base + 1 + 2 * k;
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      gt_slot[0] = t;
    }
  }
  const long long count = d_count ? min((long long)*d_count, max_count) : max_count;
  const int nb = gm.nb;
  const int H = HC > 0 ? HC : gm.H;
  const long long n_groups = (count + nb - 1) / nb;
  // PAIR: the unit of work is a PAIR of groups, the CTA of rank r takes the r-th; an odd last group leaves the peer an empty one
  // (all of its leaves >= count: it feeds zeros through the same protocol and writes nothing)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const long long unit = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long n_units = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  const long long work = PAIR ? (n_groups + 1) / 2 : n_groups;
  if (unit >= work) return;  // uniform per CTA (pair), before any barrier / TMEM use
  const int my_groups = (int)((work - unit + n_units - 1) / n_units);
  auto group_leaf0 = [&](int gi) { return ((unit + (long long)gi * n_units) * (PAIR ? 2 : 1) + rank) * nb; };
  auto group_valid = [&](long long leaf0) { return (int)max(0ll, min((long long)nb, count - leaf0)); };

  // ---- one-time setup ---------------------------------------------------------------------
  for (int i = tid; i < kRtActBytes / 16; i += K::kThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < kRtRegions * kRtRegionBlocks; ++s) mbar_init(bar_full + s, 1);
    for (int s = 0; s < kRtRegions; ++s) mbar_init(bar_empty + s, 1);
    for (int s = 0; s < kRtRegions; ++s) mbar_init(bar_fullp + s, 1);
    for (int t = 0; t < kRtMaxH; ++t) {
      mbar_init(bar_acc + t, 1);
      mbar_init(bar_act + t, PAIR ? 2 * K::kEpiWarps : K::kEpiThreads);  // PAIR: one arrival per epilogue warp of either CTA
    }
    mbar_init(bar_feat + 0, K::kEpiThreads);
    mbar_init(bar_feat + 1, K::kHeadThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == K::kMmaWarp) {
    __syncwarp();
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kRtTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kRtTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised and its TMEM allocated before anything reaches across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) TC_TRACE(7, 1);  // setup done

  if (warp >= K::kHeadWarp) {
    // ===================== head warps: FC heads of group g while the pipeline already runs group g+1; the last
    // group of a CTA is left to the epilogue warps (idle by then, and many more threads) ==========================
    const int htid = tid - K::kHeadWarp * 32;
    const int HW = gm.H * gm.W;
    float* fcv = reinterpret_cast<float*>(smem + K::kFcW);  // small head vectors: FC1 bias, FC2 weights + bias, policy bias
    for (int i = htid; i < 41 + gm.A; i += K::kHeadThreads)
      fcv[i] = i < 20 ? blob[L.val_fc1_b + i] : i < 40 ? blob[L.val_fc2_w + i - 20] : i == 40 ? blob[L.val_fc2_b] : blob[L.pol_fc_b + i - 41];
    const int fcv_floats = (41 + gm.A + 3) & ~3;
    const int fcw_floats = HW * (2 * gm.A + 20);
    const float* polw = pol_fc_t;
    const float* valw = val_fc1_t;
    if (consts.fc_in_const && nb == 16) {
      // the FC weights come through the constant cache (rt_heads): nothing to stage
    } else if (fcv_floats + fcw_floats <= K::kFcWFloats) {  // both transposed matrices are contiguous in global memory (net_tc.cu pack)
      float* fcw_s = fcv + fcv_floats;
      for (int i = htid; i < fcw_floats; i += K::kHeadThreads) fcw_s[i] = pol_fc_t[i];
      polw = fcw_s;
      valw = fcw_s + 2 * HW * gm.A;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(K::kHeadThreads) : "memory");
    for (int gi = 0; gi + 1 < my_groups; ++gi) {
      const long long leaf0 = group_leaf0(gi);
      const int nvalid = group_valid(leaf0);
      mbar_wait_relaxed(bar_feat + 0, (uint32_t)gi & 1u);
      rt_heads<K::kHeadThreads, 2>(gm, nvalid, leaf0, htid, headf_s, fc_s, consts.headb[0], consts.headb[1], consts.headb[2], fcv, polw,
                                   valw, probs, values, consts);
      mbar_arrive(bar_feat + 1);  // features consumed, slots re-zeroed, scratch free
      if (htid == 0) TC_TRACE(5, gi);
    }
    {  // the last group of this CTA: nothing is left to overlap with, the (idle) epilogue warps join in
      const int gi = my_groups - 1;
      const long long leaf0 = group_leaf0(gi);
      const int nvalid = group_valid(leaf0);
      mbar_wait(bar_feat + 0, (uint32_t)gi & 1u);
      rt_heads<K::kEpiThreads + K::kHeadThreads, 3>(gm, nvalid, leaf0, K::kEpiThreads + htid, headf_s, fc_s, consts.headb[0],
                                                    consts.headb[1], consts.headb[2], fcv, polw, valw, probs, values, consts);
      if (htid == 0) TC_TRACE(5, gi);
    }
  } else if (warp == K::kLoadWarp) {
    // ===================== weight producer: the 11 regions of the network, over and over, round-robin into the
    // three resident regions; a region is refilled as soon as the last tile of its layer has released it ==========
    if ((tid & 31) == 0) {
      const int regions_net = 2 * gm.layers - 1;  // conv_in (3 blocks + 3 unused) + 2 per residual block
      const int total = my_groups * regions_net;
      const uint8_t* wrank = wimg + (size_t)rank * regions_net * K::kRegionBytes;  // PAIR: one image per rank
      int reg = 0, src = 0;
      uint32_t round = 0;
      for (int n = 0; n < total; ++n) {
        if (round > 0) mbar_wait_relaxed(bar_empty + reg, (round - 1u) & 1u);  // spinning instead: no gain (profiles/r1_net_bench_issue_loop.txt)
        const int nblk = src == 0 ? 3 : kRtRegionBlocks;  // conv_in: three real blocks
        // ONE barrier per region (slot 0 of its six): the first tile of a layer polls two barriers instead of twelve -- its
        // N = 128 MMAs issue at the tensor pipe's own rate, so every poll between two of them is exposed (tools/cta2_probe.cu)
        uint64_t* bar = bar_full + reg * kRtRegionBlocks;
        mbar_expect_tx(bar, (uint32_t)(nblk * K::kBlockBytes));
        for (int b = 0; b < nblk; ++b)
          bulk_g2s(wgt + (reg * kRtRegionBlocks + b) * K::kBlockBytes, wrank + (size_t)(src * kRtRegionBlocks + b) * K::kBlockBytes,
                   (uint32_t)K::kBlockBytes, bar);
        if (++reg == kRtRegions) { reg = 0; ++round; }
        if (++src == regions_net) src = 0;
      }
    }
  } else if (warp == K::kMmaWarp && PAIR && rank != 0) {
    // ===================== PAIR, peer CTA: the leader issues the MMAs of both; this warp only reports the arrival of this
    // CTA's weight regions to the leader (a local wait, then one remote arrive per region) ==============================
    if ((tid & 31) == 0) {
      const int total = my_groups * (2 * gm.layers - 1);
      const uint32_t fullp_leader = mapa_a(smem_u32(bar_fullp), 0u);
      int reg = 0;
      uint32_t ph = 0;
      for (int n = 0; n < total; ++n) {
        mbar_wait(bar_full + reg * kRtRegionBlocks, ph);  // spinning: the leader's first tile of a layer waits for this report
        mbar_arrive_cluster_a(fullp_leader + 8u * (uint32_t)reg);
        if (++reg == kRtRegions) { reg = 0; ph ^= 1u; }
      }
    }
  } else if (warp == K::kMmaWarp) {
    // ===================== MMA issuer: the whole warp runs the loop (warp-uniform descriptor arithmetic), one
    // elected lane issues MMAs and commits ============================================================================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    const uint64_t a_desc0 = make_desc(smem_u32(act) + (uint32_t)kRtHalo * 16u, (uint32_t)kRtChunkBytes, 128u);
    const uint64_t b_desc0 = make_desc(smem_u32(wgt), (PAIR ? (uint32_t)kRtPairRows : 192u) * 16u, 128u);
    uint32_t full_a = smem_u32(bar_full), empty_a = smem_u32(bar_empty), acc_a = smem_u32(bar_acc), act_a = smem_u32(bar_act);
    uint32_t fullp_a = smem_u32(bar_fullp);
    asm volatile("" : "+r"(full_a), "+r"(empty_a), "+r"(acc_a), "+r"(act_a), "+r"(fullp_a));  // opaque: not re-derived from the CTA's shared window per tile
    // the timeline's condition is evaluated once: between two tiles every instruction of this warp is exposed
    const bool tr = elected && trace != nullptr && blockIdx.x == 0;
#define RT_MMA_TRACE(kind, idx)                                                 \
  do {                                                                          \
    if (tr && (idx) < 1000) trace[(kind) * 1000 + (idx)] = clock64();           \
  } while (0)
    const int n_layers = gm.layers;
    const int total_layers = my_groups * n_layers;
    int reg = 0;
    uint32_t rphase = 0;
    int layer = 0;
    bool pre_waited = false;
    // weight regions of a layer: descriptor bases, "landed" / "consumed" barriers and the phases to wait for
    struct LayerW {
      uint64_t rb0, rb1;
      uint32_t full0, full1, empty0, empty1, ph0, ph1, fullp0, fullp1;
    };
    auto take_regions = [&](bool first_l) {
      LayerW w;
      w.rb0 = b_desc0 + (uint64_t)(uint32_t)(reg * kRtRegionBlocks * K::kBlockUnits);
      w.full0 = full_a + 8u * (uint32_t)(reg * kRtRegionBlocks);
      w.empty0 = empty_a + 8u * (uint32_t)reg;
      w.fullp0 = fullp_a + 8u * (uint32_t)reg;
      w.ph0 = rphase;
      if (++reg == kRtRegions) { reg = 0; rphase ^= 1u; }
      w.rb1 = w.rb0;
      w.full1 = w.full0;
      w.empty1 = w.empty0;
      w.fullp1 = w.fullp0;
      w.ph1 = w.ph0;
      if (!first_l) {
        w.rb1 = b_desc0 + (uint64_t)(uint32_t)(reg * kRtRegionBlocks * K::kBlockUnits);
        w.full1 = full_a + 8u * (uint32_t)(reg * kRtRegionBlocks);
        w.empty1 = empty_a + 8u * (uint32_t)reg;
        w.fullp1 = fullp_a + 8u * (uint32_t)reg;
        w.ph1 = rphase;
        if (++reg == kRtRegions) { reg = 0; rphase ^= 1u; }
      }
      return w;
    };
    LayerW nxt = take_regions(true);
    for (int gl = 0; gl < total_layers; ++gl) {
      const uint32_t par = (uint32_t)gl & 1u;
      RT_MMA_TRACE(6, gl * 2);      // layer iteration entered
      const bool first = layer == 0;
      const LayerW cur = nxt;
      const uint64_t rb0 = cur.rb0, rb1 = cur.rb1;
      const uint32_t full0 = cur.full0, full1 = cur.full1, empty0 = cur.empty0, empty1 = cur.empty1, ph0 = cur.ph0, ph1 = cur.ph1;
      const uint32_t fullp0 = cur.fullp0, fullp1 = cur.fullp1;
      // the next layer's regions are worked out in the middle of tile 1 (H >= 2), not between two layers, and pinned there
      auto mid = [&]() {
        nxt = take_regions(layer + 1 == n_layers);
        asm volatile("" : "+l"(nxt.rb0), "+l"(nxt.rb1), "+r"(nxt.full0), "+r"(nxt.full1), "+r"(nxt.empty0), "+r"(nxt.empty1), "+r"(nxt.ph0),
                     "+r"(nxt.ph1));
        if (PAIR) asm volatile("" : "+r"(nxt.fullp0), "+r"(nxt.fullp1));
      };
      auto tile = [&](auto yv) {
        const int y = rt_value(yv);
        // tile y reads act[y] and writes out[y-1..y+1]: it needs the barriers of tiles y-1, y, y+1 at this stage.
        // All but the first tile's were already polled while the previous tile's MMAs were being issued.
        if (y == 0 && !pre_waited) {
          if (PAIR) {
            mbar_wait_cluster_a(act_a, par);
            mbar_wait_cluster_a(act_a + 8u, par);
          } else {
            mbar_wait_a(act_a, par);
            mbar_wait_a(act_a + 8u, par);
          }
        }
        if (y == 0) RT_MMA_TRACE(6, gl * 2 + 1);  // first tile: barriers passed
        tc_fence_after();  // orders this tile's MMAs after the barrier observations (also the early ones)
        RT_MMA_TRACE(0, gl * 8 + y);
        const uint64_t a_tile = a_desc0 + (uint64_t)(uint32_t)(y * 128);
        // keep the 24 weight descriptors of the layer from being hoisted in front of the tile loop (110 uniform-datapath
        // instructions during which the tensor pipe ran dry at every layer start): recomputed per tile, they interleave
        // with the MMA issue, which has slack
        uint64_t rb0t = rb0, rb1t = rb1;
        asm volatile("" : "+l"(rb0t), "+l"(rb1t));
        // D columns: out[y-1] | out[y] | out[y+1]; the first tile has no out[-1], the last no out[H]
        const uint32_t d_main = tmem_base + (uint32_t)(y == 0 ? 0 : (y - 1) * 64);
        const uint32_t d_new = tmem_base + (uint32_t)((y + 1) * 64);
        uint32_t nb0 = 0u, nb1 = 0u;
        uint32_t npar = par;
        if (y + 2 < H) {
          nb0 = act_a + 8u * (uint32_t)(y + 2);  // for tile y+1
        } else if (y == H - 1 && H >= 4 && gl + 1 < total_layers) {
          // next layer's first tile: its barriers depend on commits issued two or more tiles ago (H >= 4)
          nb0 = act_a;
          nb1 = act_a + 8u;
          npar = par ^ 1u;
        }
        if (first) {
          if (y == 0) rt_issue_tile<0, true, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1);
          else if (y == H - 1) rt_issue_tile<2, true, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1);
          else rt_issue_tile<1, true, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1);
        } else {
          if (y == 0) rt_issue_tile<0, false, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1, tr && gl < 100 ? trace + 5000 + gl * 6 : nullptr);
          else if (y == H - 1) rt_issue_tile<2, false, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1);
          else rt_issue_tile<1, false, PAIR, K::kF16>(y == 1, mid, elected, a_tile, rb0t, rb1t, d_main, d_new, full0, full1, ph0, ph1, empty0, empty1, nb0, nb1, npar, fullp0, fullp1);
        }
        if (elected) {
          if (PAIR) umma_commit_pair_a(acc_a + 8u * (uint32_t)y);
          else umma_commit_a(acc_a + 8u * (uint32_t)y);
        }
        RT_MMA_TRACE(1, gl * 8 + y);
        __syncwarp();
      };
      if constexpr (HC == 6) {
        tile(RtInt<0>{});
        tile(RtInt<1>{});
        tile(RtInt<2>{});
        tile(RtInt<3>{});
        tile(RtInt<4>{});
        tile(RtInt<5>{});
      } else {
#pragma unroll 1
        for (int y = 0; y < H; ++y) tile(y);
      }
      pre_waited = H >= 4 && gl + 1 < total_layers;
      if (++layer == n_layers) layer = 0;
    }
#undef RT_MMA_TRACE
  } else {
    // ========================================= epilogue warps =========================================
    const int quarter = warp & 3, cp = warp >> 2;
    constexpr int CH = K::kCH, QN = CH / 4, C8N = CH / 8;   // channels per warp, groups of four, 16-byte chunks
    const int row = quarter * 32 + (tid & 31);              // lane of the tile = TMEM lane
    const int bidx = row >> gm.pshift, col = row & (gm.pitch - 1);
    const bool real = col < gm.W;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int HW = gm.H * gm.W;
    const bool dbg_skip_epilogue = trace != nullptr && trace[7999] == 1;
    // "activation tile y rewritten, accumulator drained".  PAIR: the barrier lives in the leader CTA and counts WARPS of both
    // CTAs (every lane has fenced its own writes; the warp barrier orders them before lane 0's cluster-scope release)
    const uint32_t act_leader = PAIR ? mapa_a(smem_u32(bar_act), 0u) : 0u;
    auto act_arrive = [&](int y) {
      if (PAIR) {
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive_cluster_a(act_leader + 8u * (uint32_t)y);
      } else {
        mbar_arrive(bar_act + y);
      }
    };
    // PAIR: every MMA accumulates, so whoever reads an accumulator leaves it zeroed for the next layer
    const bool dbg_no_zero = trace != nullptr && (trace[7996] & 1);
    auto zero_acc = [&](int y, int c0) {
      if (dbg_no_zero) return;
      const uint32_t a = tmem_base + lane_base + (uint32_t)(y * 64 + c0);
      TMEM_ST16Z(a, 0u);
      TMEM_ST16Z(a + 16u, 0u);
    };

    // `writer`: the warp set that stores the planes (the one that last read this tile's activations)
    auto write_inputs = [&](int y, long long leaf0, int writer) {
      if (cp == writer) {
        uint32_t lo = 0u;
        const long long leaf = leaf0 + bidx;
        if (real && leaf < count) {
          const typename R::Board s = boards[leaf];
          const int wm = who[leaf];
          constexpr uint32_t kOne = K::kF16 ? 0x3C00u : 0x3F80u;  // 1.0 in fp16 / bf16
          const uint32_t mine = rules.plane_value(s, wm, 0, y, col) ? kOne : 0u;
          const uint32_t other = rules.plane_value(s, wm, 1, y, col) ? kOne : 0u;
          lo = mine | (other << 16);
        }
        uint8_t* dst = act + (size_t)(kRtHalo + y * 128 + row) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(lo, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + kRtChunkBytes) = make_uint4(0u, 0u, 0u, 0u);  // K is padded to 16
        fence_async_smem();
      }
      act_arrive(y);
    };

    // Loads the accumulator, bias, LeakyReLU and (HAS_RES) the residual hi (bf16, shared memory) + lo (e5m2, TMEM)
    // of 32 channels [c0, c0+32) of tile y into v[16] (pairs).
    auto load_values = [&](auto has_res_c, int layer, int y, int c0, float2* v, uint8_t* arow) {
      constexpr bool HAS_RES = decltype(has_res_c)::value;
      const uint32_t a_acc = tmem_base + lane_base + (uint32_t)(y * 64 + c0);
      const uint32_t a_lo = tmem_base + lane_base + kRtLoCol + (uint32_t)(y * 16 + c0 / 4);
      const float4* bl4 = reinterpret_cast<const float4*>(consts.bias + layer * 64 + c0);
      uint32_t ra[CH], rl[QN];
      uint4 hv[C8N];
      TMEM_LD16(a_acc, ra);
      if (CH == 32) TMEM_LD16(a_acc + 16u, (ra + CH - 16));
      if (HAS_RES) {
        if (!K::kF16) {
          if (CH == 32) TMEM_LD8(a_lo, rl);
          else TMEM_LD4(a_lo, rl);
        }
#pragma unroll
        for (int c8 = 0; c8 < C8N; ++c8) hv[c8] = *reinterpret_cast<const uint4*>(arow + (size_t)c8 * kRtChunkBytes);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float2 slope = make_float2(kLeaky, kLeaky);
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const float4 bq = bl4[q];
        const float2 x01 = __fadd2_rn(make_float2(__uint_as_float(ra[q * 4]), __uint_as_float(ra[q * 4 + 1])), make_float2(bq.x, bq.y));
        const float2 x23 = __fadd2_rn(make_float2(__uint_as_float(ra[q * 4 + 2]), __uint_as_float(ra[q * 4 + 3])), make_float2(bq.z, bq.w));
        const float2 t01 = __fmul2_rn(x01, slope), t23 = __fmul2_rn(x23, slope);
        float2 m01 = make_float2(fmaxf(x01.x, t01.x), fmaxf(x01.y, t01.y));
        float2 m23 = make_float2(fmaxf(x23.x, t23.x), fmaxf(x23.y, t23.y));
        if (HAS_RES) {
          const uint4 h4 = hv[q >> 1];
          const uint32_t w0 = (q & 1) ? h4.z : h4.x, w1 = (q & 1) ? h4.w : h4.y;
          if (K::kF16) {
            m01 = __fadd2_rn(m01, __half22float2(*reinterpret_cast<const __half2*>(&w0)));
            m23 = __fadd2_rn(m23, __half22float2(*reinterpret_cast<const __half2*>(&w1)));
          } else {
            float2 l01, l23;
            e5m2x4_to_float(rl[q], l01, l23);
            m01 = __fadd2_rn(m01, __fadd2_rn(bf16x2_to_float2(w0), l01));
            m23 = __fadd2_rn(m23, __fadd2_rn(bf16x2_to_float2(w1), l23));
          }
        }
        v[q * 2] = m01;
        v[q * 2 + 1] = m23;
      }
    };

    // One tile of one layer but the last: every warp handles its 32 channels of the tile's 128 rows and rewrites
    // the activations in place (bf16 hi -> shared memory, e5m2 lo -> TMEM).
    auto epilogue_tile = [&](auto has_res_c, int layer, int gl, int y) {
      mbar_wait(bar_acc + min(y + 1, H - 1), (uint32_t)gl & 1u);
      __syncwarp();
      tc_fence_after();
      if (tid == 0) TC_TRACE(2, gl * 8 + y);
      if (dbg_skip_epilogue) {  // debug (tools/net_trace.py): measure the MMA stream without the epilogue's traffic
        tc_fence_before();
        act_arrive(y);
        return;
      }
      uint8_t* arow = act + (size_t)(cp * C8N * kRtActRows + kRtHalo + y * 128 + row) * 16;
      float2 v[CH / 2];
      load_values(has_res_c, layer, y, cp * CH, v, arow);
      if (PAIR) zero_acc(y, cp * CH);
      const float2 minus1 = make_float2(-1.0f, -1.0f);
#pragma unroll
      for (int c8 = 0; c8 < C8N; ++c8) {
        uint32_t packed[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 vv = v[c8 * 4 + j];
          if (K::kF16) {
            const __half2 h = __floats2half2_rn(vv.x, vv.y);
            packed[j] = real ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
          } else {
            const __nv_bfloat162 h = __floats2bfloat162_rn(vv.x, vv.y);
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
            packed[j] = real ? hw : 0u;
            v[c8 * 4 + j] = __ffma2_rn(bf16x2_to_float2(hw), minus1, vv);  // lo part
          }
        }
        *reinterpret_cast<uint4*>(arow + (size_t)c8 * kRtChunkBytes) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      if (!K::kF16) {
        uint32_t rl[QN];
#pragma unroll
        for (int q = 0; q < QN; ++q) rl[q] = float_to_e5m2x4(v[q * 2], v[q * 2 + 1]);
        const uint32_t a_lo = tmem_base + lane_base + kRtLoCol + (uint32_t)(y * 16 + cp * QN);
        if (CH == 32) TMEM_ST8(a_lo, rl);
        else TMEM_ST4(a_lo, rl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      fence_async_smem();
      tc_fence_before();
      act_arrive(y);
      if (tid == 0) TC_TRACE(3, gl * 8 + y);
    };

    // One tile of the last layer: its output only feeds the 1x1 head convolutions (64 channels -> 3 features per
    // cell).  Warp set cp takes the tiles with y % 2 == cp and all 64 channels of its rows, so that every feature
    // slot receives ONE plain store (no atomics, no zeroing, a fixed summation order); the other set only passes
    // the barrier on.  The same set then writes the next group's input planes into the tile it has just read.
    auto last_tile = [&](int gi, int gl, int y, bool more, long long next_leaf0) {
      mbar_wait(bar_acc + min(y + 1, H - 1), (uint32_t)gl & 1u);
      __syncwarp();
      tc_fence_after();
      if (tid == 0) TC_TRACE(2, gl * 8 + y);
      if (!dbg_skip_epilogue && (y % K::kCP) == cp) {
        float av = 0.0f, ap0 = 0.0f, ap1 = 0.0f;
#pragma unroll 1
        for (int hh = 0; hh < K::kCP; ++hh) {
          uint8_t* arow = act + (size_t)(hh * C8N * kRtActRows + kRtHalo + y * 128 + row) * 16;
          float2 v[CH / 2];
          load_values(std::true_type{}, gm.layers - 1, y, hh * CH, v, arow);
          if (PAIR) zero_acc(y, hh * CH);
          const float4* hw4 = reinterpret_cast<const float4*>(consts.headw + hh * CH);
#pragma unroll
          for (int q = 0; q < QN; ++q) {
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
        if (PAIR) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      if (more) write_inputs(y, next_leaf0, y % K::kCP);  // by the set that just read this tile's residual
      if (tid == 0) TC_TRACE(3, gl * 8 + y);
    };

    if (PAIR) {  // all accumulators start out zero (the first MMA of a layer accumulates like every other)
      for (int c = cp * 192; c < cp * 192 + 192; c += 32) zero_acc(0, c);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
    }
    for (int y = 0; y < H; ++y) write_inputs(y, group_leaf0(0), 0);
    for (int gi = 0; gi < my_groups; ++gi) {
      const long long leaf0 = group_leaf0(gi);
      const bool more = gi + 1 < my_groups;
      const long long next_leaf0 = group_leaf0(gi + 1);
      const int gl0 = gi * gm.layers;
#pragma unroll 1
      for (int y = 0; y < H; ++y) epilogue_tile(std::false_type{}, 0, gl0, y);
#pragma unroll 1
      for (int layer = 1; layer < gm.layers - 1; ++layer) {
#pragma unroll 1
        for (int y = 0; y < H; ++y) epilogue_tile(std::true_type{}, layer, gl0 + layer, y);
      }
      if (gi > 0) mbar_wait(bar_feat + 1, (uint32_t)(gi - 1) & 1u);  // the previous group's features have been consumed
#pragma unroll 1
      for (int y = 0; y < H; ++y) last_tile(gi, gl0 + gm.layers - 1, y, more, next_leaf0);
      mbar_arrive(bar_feat + 0);  // this thread's head features are in place (release) -> head warps
      if (tid == 0) TC_TRACE(4, gi);
      if (!more) {  // join the head warps for the heads of the last group
        mbar_wait(bar_feat + 0, (uint32_t)gi & 1u);
        const int nvalid = group_valid(leaf0);
        const float* fcv = reinterpret_cast<const float*>(smem + K::kFcW);  // staged by the head warps (same rule as there)
        const int fcv_floats = (41 + gm.A + 3) & ~3;
        const bool staged = fcv_floats + HW * (2 * gm.A + 20) <= K::kFcWFloats;
        const float* polw = staged ? fcv + fcv_floats : pol_fc_t;
        const float* valw = staged ? fcv + fcv_floats + 2 * HW * gm.A : val_fc1_t;
        rt_heads<K::kEpiThreads + K::kHeadThreads, 3>(gm, nvalid, leaf0, tid, headf_s, fc_s, consts.headb[0], consts.headb[1],
                                                      consts.headb[2], fcv, polw, valw, probs, values, consts);
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  if (tid == 0) TC_TRACE(7, 2);  // epilogue warp 0 finished (incl. the last group's heads)
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA of a pair leaves (or frees TMEM) while the other may still be signalled or read
  if (tid == 0) TC_TRACE(7, 3);  // all roles finished
  if (gt_slot != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    gt_slot[1] = t;
  }
  if (warp == K::kMmaWarp) {
    __syncwarp();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kRtTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kRtTmemCols) : "memory");
  }
}

// float -> IEEE half, round to nearest even (host side of the fp16 weight image)
static uint16_t rt_f32_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t absu = u & 0x7fffffffu;
  if (absu >= 0x7f800000u) return (uint16_t)(sign | (absu > 0x7f800000u ? 0x7e00u : 0x7c00u));  // NaN / inf
  if (absu >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                   // rounds to >= 65,520: inf
  if (absu < 0x33000001u) return (uint16_t)sign;                                                // below half the smallest subnormal
  int e = (int)(absu >> 23) - 127;
  uint32_t m = (absu & 0x7fffffu) | 0x800000u;
  int shift = e >= -14 ? 13 : 13 + (-14 - e);  // bits of the 24-bit significand that do not fit
  uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  const uint32_t out = e >= -14 ? (((uint32_t)(e + 15) << 10) + (q - 0x400u)) : q;  // a carry out of the mantissa bumps the exponent
  return (uint16_t)(sign | out);
}

static uint16_t rt_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

}  // namespace caro

using namespace caro;

// Weight image: 11 regions of 6 blocks of 6,144 B (conv_in: 3 blocks used).  Block (layer, dx, k-step) = B operand
// [2 k-chunks][n = 192][8 in-channels], n = 64 j + out-channel with j = 0,1,2 <-> vertical tap ky = 2,1,0 (the
// output tile above / at / below the source tile).
int caro_net_rt_pack(caro_net* net, const float* h) {
  const BlobLayout& L = net->layout;
  const int blocks = L.blocks;
  const size_t img_bytes = (size_t)(1 + 2 * blocks) * kRtRegionBlocks * kRtBlockBytes;
  std::vector<uint16_t> img(img_bytes / 2, 0), himg(img_bytes / 2, 0);  // bf16 image, fp16 image (F16 mode)
  auto put = [&](int block, int j, int co, int c, float w) {
    const size_t off = (size_t)block * kRtBlockBytes + (size_t)(c / 8) * 3072 + (size_t)(j * 64 + co) * 16 + (size_t)(c % 8) * 2;
    img[off / 2] = rt_f32_to_bf16(w);
    himg[off / 2] = rt_f32_to_f16(w);
  };
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 3; ++j)
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 2; ++ci) put(kx, j, co, ci, h[L.conv_in_w + ((size_t)(co * 2 + ci) * 9 + (2 - j) * 3 + kx)]);
  for (int l = 0; l < blocks; ++l)
    for (int kx = 0; kx < 3; ++kx)
      for (int kk = 0; kk < 4; ++kk)
        for (int j = 0; j < 3; ++j)
          for (int co = 0; co < 64; ++co)
            for (int c = 0; c < 16; ++c)
              put(kRtRegionBlocks + l * 12 + kx * 4 + kk, j, co, c,
                  h[L.conv_w[l] + ((size_t)(co * 64 + kk * 16 + c) * 9 + (2 - j) * 3 + kx)]);
  static_assert(sizeof(RtConsts) <= sizeof(net->h_rt_consts), "caro_net::h_rt_consts is too small");
  RtConsts* hc = reinterpret_cast<RtConsts*>(net->h_rt_consts);
  for (int co = 0; co < 64; ++co) hc->bias[co] = h[L.conv_in_b + co];
  for (int l = 0; l < blocks; ++l)
    for (int co = 0; co < 64; ++co) hc->bias[(l + 1) * 64 + co] = h[L.conv_b[l] + co];
  for (int c = 0; c < 64; ++c) {
    hc->headw[c] = h[L.val_conv_w + c];
    hc->headw[64 + c] = h[L.pol_conv_w + c];
    hc->headw[128 + c] = h[L.pol_conv_w + 64 + c];
  }
  hc->headb[0] = h[L.val_conv_b];
  hc->headb[1] = h[L.pol_conv_b];
  hc->headb[2] = h[L.pol_conv_b + 1];
  hc->headb[3] = 0.0f;
  {
    const int HW = net->H * net->W, A = net->A;
    const int n = HW * (2 * A + 20);
    hc->fc_in_const = n <= kRtConstFcFloats ? 1 : 0;
    for (int i = 0; i < kRtConstFcFloats; ++i) hc->fcw[i] = 0.0f;
    if (hc->fc_in_const) {  // the layout of caro_net::d_pol_fc_t (net_tc.cu): policy [2 HW][A], then value FC1 [HW][20]
      for (int a = 0; a < A; ++a)
        for (int i = 0; i < 2 * HW; ++i) hc->fcw[(size_t)i * A + a] = h[L.pol_fc_w + (size_t)a * 2 * HW + i];
      for (int i = 0; i < 20; ++i)
        for (int c = 0; c < HW; ++c) hc->fcw[(size_t)2 * HW * A + (size_t)c * 20 + i] = h[L.val_fc1_w + (size_t)i * HW + c];
    }
  }
  // CTA-pair form: one image per cluster rank, 7 KB blocks of three windows cut out of the same 192 stacked columns
  const size_t n_blocks = (size_t)(1 + 2 * blocks) * kRtRegionBlocks;
  const size_t pair_bytes = n_blocks * kRtPairBlockBytes;
  std::vector<uint16_t> pimg(2 * pair_bytes / 2, 0);
  for (int r = 0; r < 2; ++r)
    for (size_t b = 0; b < n_blocks; ++b)
      for (int chunk = 0; chunk < 2; ++chunk)
        for (int row = 0; row < kRtPairRows; ++row) {
          const int n = row < 96 ? r * 96 + row : row < 160 ? 64 + r * 64 + (row - 96) : r * 64 + (row - 160);
          const uint16_t* src = &img[(b * kRtBlockBytes + (size_t)chunk * 3072 + (size_t)n * 16) / 2];
          uint16_t* dst = &pimg[((size_t)r * pair_bytes + b * kRtPairBlockBytes + (size_t)chunk * kRtPairRows * 16 + (size_t)row * 16) / 2];
          for (int e = 0; e < 8; ++e) dst[e] = src[e];
        }
  cudaError_t ce = cudaSuccess;
  if (!net->d_rt_weights) ce = cudaMalloc(&net->d_rt_weights, img_bytes);  // the depth of a handle never changes (caro_net_update checks the blob size)
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_rt_weights, img.data(), img_bytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !net->d_rt_f16_weights) ce = cudaMalloc(&net->d_rt_f16_weights, img_bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_rt_f16_weights, himg.data(), img_bytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !net->d_rt_pair_weights) ce = cudaMalloc(&net->d_rt_pair_weights, 2 * pair_bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_rt_pair_weights, pimg.data(), 2 * pair_bytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !net->d_rt_scratch)
    ce = cudaMalloc(&net->d_rt_scratch, (size_t)kRtScratchSlots * net->sm_count * (kRtHeadFloats + kRtFcFloats) * sizeof(float));
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

void caro_net_rt_free(caro_net* net) {
  if (net->d_rt_weights) cudaFree(net->d_rt_weights);
  if (net->d_rt_pair_weights) cudaFree(net->d_rt_pair_weights);
  if (net->d_rt_f16_weights) cudaFree(net->d_rt_f16_weights);
  net->d_rt_f16_weights = nullptr;
  if (net->d_rt_scratch) cudaFree(net->d_rt_scratch);
  net->d_rt_weights = nullptr;
  net->d_rt_pair_weights = nullptr;
  net->d_rt_scratch = nullptr;
}

bool caro_net_rt_supports(const caro_net* net) { return net->H >= 2 && net->H <= kRtMaxH && net->W >= 2 && net->W <= kRtMaxW; }

using RtK = RtCfg<CARO_RT_CP>;
using RtKPair = RtCfg<2, true>;  // the pair form keeps two channel parts
using RtKF16 = RtCfg<CARO_RT_CP, false, true>;

// CARO_RT_PAIR=1/0 switches the CTA-pair form of the tower on / off (read once)
static bool rt_pair_enabled() {
  static const bool on = getenv("CARO_RT_PAIR") ? atoi(getenv("CARO_RT_PAIR")) != 0 : false;
  return on;
}

template <class R>
static int launch_rt(const R& rules, caro_net* net, const void* boards, const uint8_t* who, const int32_t* d_count,
                     int64_t max_count, float* probs, float* values, int mode, cudaStream_t st) {
  const bool pair = mode == 1;  // mode: 0 = bf16, 1 = bf16 as CTA pairs, 2 = fp16
  RtGeom gm;
  gm.H = net->H;
  gm.W = net->W;
  gm.A = net->A;
  gm.pshift = net->W < 4 ? 2 : 3;
  gm.pitch = 1 << gm.pshift;
  gm.nb = 128 / gm.pitch;
  gm.layers = 1 + net->layout.blocks;
  if (gm.nb * 3 * gm.H * gm.W > kRtHeadFloats || gm.nb * (20 + gm.A) > kRtFcFloats)
    return caro_fail(CARO_E_ARG, "board does not fit the row-tiled tensor-core geometry");
  const long long max_groups = (max_count + gm.nb - 1) / gm.nb;
  static const int generic_env = getenv("CARO_RT_GENERIC") ? atoi(getenv("CARO_RT_GENERIC")) : 0;  // A/B: 1 = run-time H
  // SMs for the persistent tower: the user's limit, else (inside the parts pipeline) all but a ninth of the SMs, which stay
  // with the other parts' tree kernels -- their dependent chain expand -> select -> plan, not the tower, sets the
  // pipeline's period once they only get the leftover warp slots next to tower CTAs (tools/pipeline_trace.py)
  const int lim = net->grid_limit > 0 ? net->grid_limit : net->pipeline_limit;
  const int ctas = lim > 0 && lim < net->sm_count ? lim : net->sm_count;
  if ((pair || rt_pair_enabled()) && max_groups >= 2 && ctas >= 2) {
    // clusters of two CTAs (one TPC each); every launch in flight gets its own slot of the global head scratch -- launches of
    // different pipeline parts overlap on the GPU, and the slot is baked into a captured graph node
    const long long max_units = (max_groups + 1) / 2;
    const unsigned pairs = (unsigned)(max_units < ctas / 2 ? max_units : ctas / 2);
    float* scratch = (float*)net->d_rt_scratch +
                     (size_t)(net->rt_scratch_seq++ % kRtScratchSlots) * net->sm_count * (kRtHeadFloats + kRtFcFloats);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(RtKPair::kThreads);
    cfg.dynamicSmemBytes = RtKPair::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto pkern = gm.H == 6 && !generic_env ? net_rt_kernel<R, RtKPair, 6> : net_rt_kernel<R, RtKPair, 0>;
    const cudaError_t ce = cudaLaunchKernelEx(
        &cfg, pkern, rules, gm, (const typename R::Board*)boards, who, d_count, (long long)max_count,
        (const uint8_t*)net->d_rt_pair_weights, *reinterpret_cast<const RtConsts*>(net->h_rt_consts), (const float*)net->d_blob, net->layout,
        (const float*)net->d_pol_fc_t, (const float*)(net->d_pol_fc_t + (size_t)2 * net->H * net->W * net->A), probs, values,
        (long long*)net->d_trace, scratch);
    if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
    return caro_check_launch("net_rt_kernel (pair)");
  }
  const unsigned grid = (unsigned)(max_groups < ctas ? max_groups : ctas);
  if (mode == 2) {
    auto fkern = gm.H == 6 && !generic_env ? net_rt_kernel<R, RtKF16, 6> : net_rt_kernel<R, RtKF16, 0>;
    fkern<<<grid, RtKF16::kThreads, RtKF16::kTotal, st>>>(
        rules, gm, (const typename R::Board*)boards, who, d_count, (long long)max_count, (const uint8_t*)net->d_rt_f16_weights,
        *reinterpret_cast<const RtConsts*>(net->h_rt_consts), net->d_blob, net->layout, net->d_pol_fc_t,
        net->d_pol_fc_t + (size_t)2 * net->H * net->W * net->A, probs, values, (long long*)net->d_trace, nullptr);
    return caro_check_launch("net_rt_kernel (fp16)");
  }
  auto kern = gm.H == 6 && !generic_env ? net_rt_kernel<R, RtK, 6> : net_rt_kernel<R, RtK, 0>;
  kern<<<grid, RtK::kThreads, RtK::kTotal, st>>>(
      rules, gm, (const typename R::Board*)boards, who, d_count, (long long)max_count, (const uint8_t*)net->d_rt_weights,
      *reinterpret_cast<const RtConsts*>(net->h_rt_consts), net->d_blob, net->layout, net->d_pol_fc_t, net->d_pol_fc_t + (size_t)2 * net->H * net->W * net->A, probs,
      values, (long long*)net->d_trace, nullptr);
  return caro_check_launch("net_rt_kernel");
}

int caro_net_rt_prepare() {
  cudaError_t ce = cudaSuccess;
  auto set = [&](auto kern, int bytes) {
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  set(net_rt_kernel<C4Rules, RtK, 0>, RtK::kTotal);
  set(net_rt_kernel<C4Rules, RtK, 6>, RtK::kTotal);
  set(net_rt_kernel<MnkRules, RtK, 0>, RtK::kTotal);
  set(net_rt_kernel<MnkRules, RtK, 6>, RtK::kTotal);
  set(net_rt_kernel<C4Rules, RtKF16, 0>, RtKF16::kTotal);
  set(net_rt_kernel<C4Rules, RtKF16, 6>, RtKF16::kTotal);
  set(net_rt_kernel<MnkRules, RtKF16, 0>, RtKF16::kTotal);
  set(net_rt_kernel<MnkRules, RtKF16, 6>, RtKF16::kTotal);
  set(net_rt_kernel<C4Rules, RtKPair, 0>, RtKPair::kTotal);
  set(net_rt_kernel<C4Rules, RtKPair, 6>, RtKPair::kTotal);
  set(net_rt_kernel<MnkRules, RtKPair, 0>, RtKPair::kTotal);
  set(net_rt_kernel<MnkRules, RtKPair, 6>, RtKPair::kTotal);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

int caro_net_rt_forward(caro_net* net, int game, int n, int k, const void* d_boards, const uint8_t* d_who,
                        const int32_t* d_count, int64_t max_count, float* d_probs, float* d_values, int mode, cudaStream_t st) {
  if (game == CARO_GAME_CONNECT4) return launch_rt<C4Rules>(C4Rules(), net, d_boards, d_who, d_count, max_count, d_probs, d_values, mode, st);
  return launch_rt<MnkRules>(MnkRules{n, k}, net, d_boards, d_who, d_count, max_count, d_probs, d_values, mode, st);
}
