// Shared pieces of the row-tiled tensor-core towers (net_rt.cu: one-pass bf16; net_rx.cu: split-precision fp16 x 3):
// geometry, by-value constants, TMEM load / store macros, address-based mbarrier helpers and the FC heads.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "net.h"
#include "tc_common.cuh"

#ifndef CARO_RT_SMEM_LIMIT
#define CARO_RT_SMEM_LIMIT (232448 - 1024)  // one more 1 KB block reservation fits beside this CTA on the SM
#endif

namespace caro {

constexpr int kRtMaxH = 6;
constexpr int kRtMaxW = 7;
constexpr int kRtHalo = 2;                                   // zero rows before / after the tiles (>= 1)
constexpr int kRtActRows = kRtMaxH * 128 + 2 * kRtHalo;      // 772
constexpr int kRtChunkBytes = kRtActRows * 16;               // one 8-channel chunk of all rows
constexpr int kRtActBytes = 8 * kRtChunkBytes;               // 98,816
constexpr int kRtBlockBytes = 2 * 192 * 16;                  // 6,144: B operand of one (dx, k-step): [2 chunks][192][8]
constexpr int kRtBlockUnits = kRtBlockBytes / 16;            // in descriptor units
constexpr int kRtRegionBlocks = 6;                           // a weight region = half a layer
constexpr int kRtRegions = 3;                                // resident regions: layer L in two, the first half of L+1 in the third
constexpr int kRtRegionBytes = kRtRegionBlocks * kRtBlockBytes;          // 36,864
constexpr int kRtMaxLayers = 1 + kMaxBlocks;                 // conv_in + residual blocks (the depth is a run-time value: RtGeom::layers)
constexpr uint32_t kRtTmemCols = 512;
constexpr uint32_t kRtLoCol = 384;                           // e5m2 lo residual: tile y at 384 + 16 y
constexpr int kRtConstFcFloats = 1536;                       // FC weights that travel in the kernel parameters
constexpr int kRtHeadFloats = 2016;                          // max nb * 3 * H * W (16 Connect4 boards)
constexpr int kRtFcFloats = 928;                             // max nb * (20 + A) (32 3x3 boards)

__host__ __device__ constexpr uint32_t rt_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// operand format of the one-pass row-tiled tower: bf16 (a / b format fields = 1) or fp16 (fields = 0)
template <bool F16>
__host__ __device__ constexpr uint32_t rt_idesc_fmt(uint32_t n) {
  return F16 ? ((1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24)) : rt_idesc(n);
}
// CTA-pair form (cta_group::2, net_rt.cu with PAIR): M = 256 = the 128 lanes of both CTAs, each CTA supplies half of B
__host__ __device__ constexpr uint32_t rt_idesc_pair(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}
// B block of the pair form, per CTA: rows 0..95 = its half of the 192 stacked columns (interior tiles), 96..159 = its half of
// columns 64..191 (first board row: no out[-1]), 160..223 = its half of columns 0..127 (last board row: no out[H]).  One
// descriptor addresses BOTH CTAs' copies, so each window has to start at the same offset in both.
constexpr int kRtPairRows = 224;
constexpr int kRtPairBlockBytes = 2 * kRtPairRows * 16;       // 7,168
constexpr uint32_t kRtPairTop = 96, kRtPairBottom = 160;      // window starts in rows (= descriptor units)
constexpr int kRtScratchSlots = 16;                           // launches of one network that may be in flight at once (net_rx.cu, pair form):
                                                              // their head features / FC scratch live in global memory, one slot per launch, round robin

struct RtGeom {
  int H, W, A, pitch, pshift, nb;
  int layers;  // 1 + residual blocks of this network
};

// Small per-network constants passed BY VALUE as a __grid_constant__ kernel parameter: they are read through the
// constant cache (LDC), not through the shared-memory pipe, which the tensor cores' operand fetches saturate.
struct RtConsts {
  float bias[kRtMaxLayers * 64];   // folded conv biases
  float headw[3 * 64];          // 1x1 head convolutions: value, policy 0, policy 1
  float headb[4];               // their (folded) biases
  // transposed FC weights, policy [2 HW][A] then value FC1 [HW][20], when they fit (Connect4: 1,428 floats): the FC heads
  // then read their weights through the constant cache too; read from shared memory they were a quarter of the
  // kernel's LSU wavefronts, on the pipe that bounds the tower
  float fcw[kRtConstFcFloats];
  int fc_in_const, pad_[3];
};

#define TMEM_LD8(addr, r)                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"     \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(addr))
#define TMEM_LD4(addr, r) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr))
#define TMEM_ST8(addr, r)                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"     \
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
#define TMEM_ST4(addr, r) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory")

// four e5m2 values (one 32-bit TMEM column) -> two float pairs.  e5m2 is the upper byte of an fp16, so the unpack
// is a byte permute into two half2 registers.
__device__ __forceinline__ void e5m2x4_to_float(uint32_t w, float2& f01, float2& f23) {
  const uint32_t p0 = __byte_perm(w, 0u, 0x1404u), p1 = __byte_perm(w, 0u, 0x3424u);
  f01 = __half22float2(*reinterpret_cast<const __half2*>(&p0));
  f23 = __half22float2(*reinterpret_cast<const __half2*>(&p1));
}
__device__ __forceinline__ uint32_t float_to_e5m2x4(float2 f01, float2 f23) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(f01, __NV_SATFINITE, __NV_E5M2);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(f23, __NV_SATFINITE, __NV_E5M2);
  return lo | (hi << 16);
}
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// Fully connected heads of one group by the head warps (lib/model.py:56-72,90-93 + the softmax of lib/mcts.py:216):
// bias + LeakyReLU of the 1x1 head convolutions in place, then every (board, output) dot product in parallel with
// the transposed FC weights read from shared memory (or from global memory when the board is too large for them
// to fit), then value FC2 + tanh and the softmax over ALL actions, one warp per board.
template <int TEAM, int BAR>
__device__ __noinline__ void rt_heads(const RtGeom& gm, int nvalid, long long leaf0, int ttid, float* headf_s, float* fc_s,
                                      float hb0, float hb1, float hb2, const float* fcv, const float* polw, const float* valw,
                                      float* __restrict__ probs, float* __restrict__ values, const RtConsts& consts) {
  // fcv (shared memory): value FC1 bias [20], value FC2 weights [20], value FC2 bias [1], policy FC bias [A]
  const int HW = gm.H * gm.W, A = gm.A;
  const int per_board = 20 + A;
  float* hid = fc_s;  // [nb][20] value hidden units, logits behind them
  float* logit = fc_s + gm.nb * 20;
  if (consts.fc_in_const && gm.nb == 16) {
    // Weights from the constant cache: a warp works on two outputs (i, i + 1) for all 16 boards at once (lane = board +
    // 16 * (output & 1)), so a weight is one of two constant addresses per instruction and a feature one conflict-free
    // shared-memory wavefront (boards are 3 HW floats apart); bias + LeakyReLU of the 1x1 convolutions on the fly.  The
    // summation order of every output is that of the general path below (four partial sums by cell index mod 4).
    const int lane = ttid & 31, b = lane & 15, odd = lane >> 4;
    const int vpairs = 10, ppairs = (A + 1) >> 1;
    const float* fb = headf_s + (size_t)b * 3 * HW;
#pragma unroll 1
    for (int item = ttid >> 5; item < vpairs + ppairs; item += TEAM / 32) {
      const bool is_val = item < vpairs;
      const int i = is_val ? 2 * item + odd : 2 * (item - vpairs) + odd;   // output index within its head
      const bool live = b < nvalid && (is_val || i < A);
      const int stride = is_val ? 20 : A;
      const int woff = (is_val ? 2 * HW * A : 0) + (live ? i : 0);
      const int n = is_val ? HW : 2 * HW;
      const float* f = is_val ? fb : fb + HW;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
      auto feat = [&](int c) { return lrelu_tc(f[c] + (is_val ? hb0 : (c < HW ? hb1 : hb2))); };
      int c = 0;
#pragma unroll 2
      for (; c + 3 < n; c += 4) {
        a0 = fmaf(consts.fcw[woff + c * stride], feat(c), a0);
        a1 = fmaf(consts.fcw[woff + (c + 1) * stride], feat(c + 1), a1);
        a2 = fmaf(consts.fcw[woff + (c + 2) * stride], feat(c + 2), a2);
        a3 = fmaf(consts.fcw[woff + (c + 3) * stride], feat(c + 3), a3);
      }
      for (; c < n; ++c) a0 = fmaf(consts.fcw[woff + c * stride], feat(c), a0);
      const float acc = (a0 + a1) + (a2 + a3);
      if (live) {
        if (is_val) hid[b * 20 + i] = lrelu_tc(fcv[i] + acc);
        else logit[b * A + i] = fcv[41 + i] + acc;
      }
    }
  } else {
#pragma unroll 1
  for (int b = 0; b < nvalid; ++b) {  // bias + LeakyReLU of the 1x1 head convolutions, in place
    float* fb = headf_s + (size_t)b * 3 * HW;
    for (int i = ttid; i < 3 * HW; i += TEAM) fb[i] = lrelu_tc(fb[i] + (i < HW ? hb0 : (i < 2 * HW ? hb1 : hb2)));
  }
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(TEAM) : "memory");
#pragma unroll 1
  for (int o = ttid; o < nvalid * per_board; o += TEAM) {
    const int b = o / per_board, i = o - b * per_board;
    const float* feat = headf_s + (size_t)b * 3 * HW;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    const bool is_val = i < 20;
    const float* wt = is_val ? valw + i : polw + (i - 20);
    const int stride = is_val ? 20 : A;
    const float* f = is_val ? feat : feat + HW;
    const int n = is_val ? HW : 2 * HW;
    int c = 0;
#pragma unroll 2
    for (; c + 3 < n; c += 4) {
      a0 = fmaf(wt[(size_t)c * stride], f[c], a0);
      a1 = fmaf(wt[(size_t)(c + 1) * stride], f[c + 1], a1);
      a2 = fmaf(wt[(size_t)(c + 2) * stride], f[c + 2], a2);
      a3 = fmaf(wt[(size_t)(c + 3) * stride], f[c + 3], a3);
    }
    for (; c < n; ++c) a0 = fmaf(wt[(size_t)c * stride], f[c], a0);
    const float acc = (a0 + a1) + (a2 + a3);
    if (is_val) hid[b * 20 + i] = lrelu_tc(fcv[i] + acc);
    else logit[b * A + (i - 20)] = fcv[41 + (i - 20)] + acc;
  }
  }
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(TEAM) : "memory");
#pragma unroll 1
  for (int b = ttid >> 5; b < nvalid; b += TEAM / 32) {
    const int lane = ttid & 31;
    float part = lane < 20 ? fcv[20 + lane] * hid[b * 20 + lane] : 0.0f;  // value FC2: fixed-order butterfly sum
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) values[leaf0 + b] = tanhf(fcv[40] + part);
    const float* lrow = logit + b * A;
    float* prow = probs + (size_t)(leaf0 + b) * A;
    float mx = -INFINITY;
#pragma unroll 1
    for (int a = lane; a < A; a += 32) mx = fmaxf(mx, lrow[a]);
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.0f;
#pragma unroll 1
    for (int a = lane; a < A; a += 32) sum += expf(lrow[a] - mx);
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll 1
    for (int a = lane; a < A; a += 32) prow[a] = expf(lrow[a] - mx) / sum;
  }
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(TEAM) : "memory");  // hid / logit / features are reused by the next group
}

// mbarrier wait / tcgen05.commit on a 32-bit shared-memory ADDRESS: the MMA warp's loop keeps its barriers as addresses, so that
// no generic -> shared conversion (an S2UR of the shared window + 64-bit arithmetic per barrier) sits between two tiles, where the
// issuing thread's time is not hidden by queued MMAs (tools/cta2_probe.cu: ~95 cycles next to a commit, ~265 elsewhere)
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_a(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_a(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pair (cluster of two, cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the address of this CTA's shared-memory location `addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_a(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive in the form CUTLASS' ClusterBarrier::arrive(cta_id) uses (default .release.cta): a cluster-scope release
// compiles to a MEMBAR that waits for every outstanding memory operation of the thread -- ncu showed the epilogue warps
// stalled on it for as long as on all other reasons together (profiles/r2_net_rt_pair_ncu.txt)
__device__ __forceinline__ void mbar_arrive_cluster_a(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_a(uint32_t bar, uint32_t parity) {  // arrivals come from the peer CTA too
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_a(uint32_t bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
#define TMEM_ST16Z(addr, z)                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" \
               ::"r"(addr), "r"(z) : "memory")

}  // namespace caro
