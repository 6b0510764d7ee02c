// Fused policy/value tower on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// lib/model.py:82-94 (eval-mode BatchNorm folded on the host) as ONE persistent kernel:
//   * one CTA per SM; a CTA takes a "group" of whole boards laid out as 512 padded positions
//     (row pitch W+1, one zero row after every board) = 4 UMMA M-tiles of 128 rows;
//   * every 3x3 convolution is an implicit GEMM: for each of the 9 taps, D[128 x 64] +=
//     A_tap[128 x 64] * W_tap[64 x 64]^T, where A_tap is the SAME shared-memory activation buffer
//     addressed with a row offset of (dy*pitch + dx) -- the activations are kept in the no-swizzle
//     K-major core-matrix layout (8 rows x 16 B contiguous), in which a row shift is just a
//     16-byte-granular change of the descriptor start address, so no im2col copy is ever made;
//   * bf16 operands, fp32 accumulation in TMEM (4 tiles x 64 columns); the fp32 residual stream
//     v <- v + lrelu(conv(v)) also lives in TMEM (4 x 64 columns), only the bf16 copy that feeds
//     the next layer's MMAs goes back to shared memory -- nothing but the boards (16 B / 64 B per
//     leaf) and the priors/value (A+1 floats) touches HBM per leaf;
//   * per-layer weight images (72 KB, pre-packed on the host in UMMA B-operand layout) stream from
//     L2 with cp.async.bulk + mbarrier while the previous layer's epilogue runs;
//   * heads (1x1 convs, FCs, tanh, softmax) run on the CUDA cores out of the last epilogue.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/caro_b200.h"
#include "common_host.h"
#include "net.h"
#include "rules.cuh"

namespace caro {

constexpr int kTilesPerGroup = 4;
constexpr int kTileRows = 128;
constexpr int kGroupRows = kTilesPerGroup * kTileRows;  // 512 padded positions per CTA pass
constexpr int kHalo = 24;                                // zero positions before / after (>= pitch + 1)
constexpr int kActRows = kGroupRows + 2 * kHalo;         // 560
constexpr int kChunkBytes = kActRows * 16;               // one 8-channel chunk of all positions
constexpr int kActBytes = 8 * kChunkBytes;               // 71,680
constexpr int kTapBytes = 8 * 64 * 16;                   // 8,192: one tap of a 64->64 layer
constexpr int kTapBytesIn = 2 * 64 * 16;                 // 2,048: one tap of conv_in (K padded to 16)
constexpr int kLayerBytes = 9 * kTapBytes;               // 73,728
constexpr int kLayerBytesIn = 9 * kTapBytesIn;           // 18,432
constexpr int kNumLayers = 1 + kBlocks;                  // conv_in + 5 residual blocks
constexpr int kEpiThreads = 128;                         // warps 0-3: epilogue (TMEM lane quarter = warp index)
constexpr int kThreads = kEpiThreads + 32;               // warp 4: TMEM owner, weight producer, MMA issuer
constexpr uint32_t kTmemCols = 512;

struct TcGeom {
  int H, W, A, pitch, block, boards_per_group;
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#define TMEM_LD16(addr, r)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])   \
               : "r"(addr))
#define TMEM_ST16(addr, r)                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),     \
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])             \
               : "memory")

// Shared-memory matrix descriptor, no-swizzle K-major canonical layout (cute::UMMA::SmemDescriptor):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (stride between the two 8-element K chunks)
//   | [32,46) SBO >> 4 (stride between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N=64, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ float lrelu_tc(float x) { return fmaxf(x, kLeaky * x); }

struct TcSmem {
  // dynamic shared memory carve-up (byte offsets from a 128-aligned base)
  static constexpr int kAct = 0;
  static constexpr int kWgt = kAct + kActBytes;                    // 2 layer images (double buffer)
  static constexpr int kBias = kWgt + 2 * kLayerBytes;             // float [6][64]
  static constexpr int kHeadW = kBias + kNumLayers * 64 * 4;       // float [3][64] + [3] biases (+pad)
  static constexpr int kHeadF = kHeadW + 4 * 64 * 4;               // float [512][3] head features
  static constexpr int kFc = kHeadF + kGroupRows * 3 * 4;          // float hidden[32][20] then logits[512]... (640 floats)
  static constexpr int kBars = kFc + 640 * 4;                      // mbarriers + tmem base
  static constexpr int kTotal = kBars + 128;
};
static_assert(TcSmem::kTotal <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");

// ------------------------------------------------------------------------------------- kernel
// Roles: warps 0-3 = epilogue (warp w owns TMEM lanes [32w, 32w+32) = rows 32w.. of every tile),
//        warp 4    = TMEM allocation, weight streaming (cp.async.bulk) and single-thread MMA issue.
// Pipeline per layer L (tiles t = 0..3, in-place activation buffer):
//   MMA(L,t)  needs act_ready[t-1..t+1] of the previous stage (their bf16 rows + tile t's accumulator drained)
//   EPI(L,t)  needs acc_full[min(t+1,3)]  (tile t+1 reads the last rows of tile t as its halo)
// so the epilogue of tile t runs underneath the MMAs of tile t+2 / the next layer's tile t-1.
template <class R>
__global__ void __launch_bounds__(kThreads, 1)
net_tc_kernel(R rules, TcGeom gm, const typename R::Board* __restrict__ boards, const uint8_t* __restrict__ who,
              const int32_t* __restrict__ d_count, long long max_count, const uint8_t* __restrict__ wimg,
              const float* __restrict__ bias_g, const float* __restrict__ blob, BlobLayout L,
              const float* __restrict__ pol_fc_t, const float* __restrict__ val_fc1_t, float* __restrict__ probs,
              float* __restrict__ values) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* act = smem + TcSmem::kAct;
  uint8_t* wgt = smem + TcSmem::kWgt;
  float* bias_s = reinterpret_cast<float*>(smem + TcSmem::kBias);
  float* headw_s = reinterpret_cast<float*>(smem + TcSmem::kHeadW);
  float* headf_s = reinterpret_cast<float*>(smem + TcSmem::kHeadF);
  float* fc_s = reinterpret_cast<float*>(smem + TcSmem::kFc);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + TcSmem::kBars);  // [2] weights buffer filled
  uint64_t* bar_acc = bar_w + 2;                                        // [4] accumulator tile complete
  uint64_t* bar_act = bar_w + 6;                                        // [4] activation tile rewritten / accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 10);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const long long count = d_count ? min((long long)*d_count, max_count) : max_count;
  const int nb = gm.boards_per_group;
  const long long n_groups = (count + nb - 1) / nb;
  if ((long long)blockIdx.x >= n_groups) return;  // uniform per CTA, before any barrier / TMEM use
  const int my_groups = (int)((n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x);

  // ---- one-time setup ---------------------------------------------------------------------
  for (int i = tid; i < kActBytes / 16; i += kThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < kNumLayers * 64; i += kThreads) bias_s[i] = bias_g[i];
  for (int i = tid; i < 64; i += kThreads) {
    headw_s[i] = blob[L.val_conv_w + i];
    headw_s[64 + i] = blob[L.pol_conv_w + i];
    headw_s[128 + i] = blob[L.pol_conv_w + 64 + i];
  }
  if (tid == 0) {
    headw_s[192] = blob[L.val_conv_b];
    headw_s[193] = blob[L.pol_conv_b];
    headw_s[194] = blob[L.pol_conv_b + 1];
    mbar_init(bar_w + 0, 1);
    mbar_init(bar_w + 1, 1);
    for (int t = 0; t < 4; ++t) {
      mbar_init(bar_acc + t, 1);
      mbar_init(bar_act + t, kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== producer / MMA issuer: the whole warp runs the loop (so that the descriptor
    // arithmetic stays warp-uniform -> uniform registers), one elected lane issues TMA / MMA / commit ==========
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    const uint32_t act_addr = smem_u32(act);
    const uint32_t wgt_addr = smem_u32(wgt);
    const int total_layers = my_groups * kNumLayers;
    auto load_layer = [&](int gl) {  // global layer index -> buffer gl & 1
      const int l = gl % kNumLayers;
      uint8_t* dst = wgt + (gl & 1) * kLayerBytes;
      uint64_t* bar = bar_w + (gl & 1);
      if (l == 0) {
        mbar_expect_tx(bar, kLayerBytesIn);
        for (int tap = 0; tap < 9; ++tap) bulk_g2s(dst + tap * kTapBytesIn, wimg + tap * kTapBytesIn, kTapBytesIn, bar);
      } else {
        mbar_expect_tx(bar, kLayerBytes);
        const uint8_t* src = wimg + kLayerBytesIn + (size_t)(l - 1) * kLayerBytes;
        for (int tap = 0; tap < 9; ++tap) bulk_g2s(dst + tap * kTapBytes, src + tap * kTapBytes, kTapBytes, bar);
      }
    };
    if (elected) {
      load_layer(0);
      if (total_layers > 1) load_layer(1);
    }
    // descriptor templates: only the 14-bit start-address field (units of 16 B) changes per MMA
    const uint64_t a_desc0 = make_desc(act_addr + (uint32_t)kHalo * 16u, kChunkBytes, 128u);
    const uint64_t b_desc0 = make_desc(wgt_addr, 1024u, 128u);
    const int pitch = gm.pitch;
    for (int gl = 0; gl < total_layers; ++gl) {
      const bool first = (gl % kNumLayers) == 0;
      const uint64_t b_layer = b_desc0 + (uint64_t)((uint32_t)(gl & 1) * (kLayerBytes / 16));
      mbar_wait(bar_w + (gl & 1), (uint32_t)(gl >> 1) & 1u);
      const uint32_t act_par = (uint32_t)gl & 1u;  // stage gl of bar_act = "input / epilogue of layer gl-1"
      for (int t = 0; t < kTilesPerGroup; ++t) {
        if (t == 0) {
          mbar_wait(bar_act + 0, act_par);
          mbar_wait(bar_act + 1, act_par);
        } else if (t < kTilesPerGroup - 1) {
          mbar_wait(bar_act + t + 1, act_par);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(t * 64);
        const uint64_t a_tile = a_desc0 + (uint64_t)(uint32_t)(t * kTileRows);
        if (elected) {
          if (first) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int sh = (tap / 3 - 1) * pitch + (tap % 3 - 1);
              umma_bf16(d_tmem, a_tile + (uint64_t)(int64_t)sh, b_layer + (uint64_t)(tap * (kTapBytesIn / 16)), kIdesc, tap > 0 ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int sh = (tap / 3 - 1) * pitch + (tap % 3 - 1);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(d_tmem, a_tile + (uint64_t)(int64_t)(sh + kk * 2 * kActRows),
                          b_layer + (uint64_t)(tap * (kTapBytes / 16) + kk * (2048 / 16)), kIdesc, (tap | kk) ? 1u : 0u);
            }
          }
          umma_commit(bar_acc + t);
          // after tile 1 has been issued every MMA of layer gl-1 is known complete (tile 1 waited on the
          // epilogue of tile 2, which waited on the last commit of layer gl-1): its weight buffer is free
          if (t == 1 && gl >= 1 && gl + 1 < total_layers) load_layer(gl + 1);
        }
        __syncwarp();
      }
    }
  } else {
    // ========================================= epilogue warps =========================================
    // Code size matters here (the v1 kernel was 127 KB of SASS and lived in instruction-cache misses):
    // tiles and 32-channel chunks are real loops, only the 32-element body is unrolled.
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int HW = gm.H * gm.W;
    float* featc = headf_s;  // dense head features [board][3][HW]

    auto row_cell = [&](int p, int& b) -> int {  // r*W + c of padded position p, or -1 for padding
      b = p / gm.block;
      const int within = p - b * gm.block;
      const int r = within / gm.pitch, c = within - r * gm.pitch;
      return (b < nb && r < gm.H && c < gm.W) ? r * gm.W + c : -1;
    };

    auto write_inputs = [&](long long leaf0) {
#pragma unroll 1
      for (int t = 0; t < kTilesPerGroup; ++t) {
        const int p = t * kTileRows + tid;
        int b;
        const int cell = row_cell(p, b);
        uint32_t lo = 0u;
        const long long leaf = leaf0 + b;
        if (cell >= 0 && leaf < count) {
          const typename R::Board s = boards[leaf];
          const int wm = who[leaf];
          const int r = cell / gm.W, c = cell - r * gm.W;
          const uint32_t mine = rules.plane_value(s, wm, 0, r, c) ? 0x3F80u : 0u;  // bf16(1.0)
          const uint32_t other = rules.plane_value(s, wm, 1, r, c) ? 0x3F80u : 0u;
          lo = mine | (other << 16);
        }
        *reinterpret_cast<uint4*>(act + (size_t)(0 * kActRows + kHalo + p) * 16) = make_uint4(lo, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(act + (size_t)(1 * kActRows + kHalo + p) * 16) = make_uint4(0u, 0u, 0u, 0u);
        fence_async_smem();
        mbar_arrive(bar_act + t);
      }
    };

    write_inputs((long long)blockIdx.x * nb);
    for (int gi = 0; gi < my_groups; ++gi) {
      const long long grp = blockIdx.x + (long long)gi * gridDim.x;
      const long long leaf0 = grp * nb;
#pragma unroll 1
      for (int layer = 0; layer < kNumLayers; ++layer) {
        const int gl = gi * kNumLayers + layer;
        const uint32_t acc_par = (uint32_t)gl & 1u;
        const bool last = layer == kNumLayers - 1;
        const bool has_res = layer > 0;
        const float4* bl4 = reinterpret_cast<const float4*>(bias_s + layer * 64);
#pragma unroll 1
        for (int t = 0; t < kTilesPerGroup; ++t) {
          mbar_wait(bar_acc + (t + 1 < kTilesPerGroup ? t + 1 : kTilesPerGroup - 1), acc_par);
          __syncwarp();
          tc_fence_after();
          const int p = t * kTileRows + tid;
          int b;
          const int cell = row_cell(p, b);
          const bool real = cell >= 0;
          float av = 0.0f, ap0 = 0.0f, ap1 = 0.0f;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const uint32_t a_acc = tmem_base + lane_base + (uint32_t)(t * 64 + half * 32);
            const uint32_t a_res = a_acc + 256u;
            uint32_t ra[32], rr[32];
            TMEM_LD16(a_acc, ra);
            TMEM_LD16(a_acc + 16u, (ra + 16));
            if (has_res) {
              TMEM_LD16(a_res, rr);
              TMEM_LD16(a_res + 16u, (rr + 16));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bq = bl4[half * 8 + q];
              const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = q * 4 + e;
                float v = lrelu_tc(__uint_as_float(ra[j]) + bb[e]);
                if (has_res) v += __uint_as_float(rr[j]);
                rr[j] = __float_as_uint(v);
              }
            }
            if (!last) {
              TMEM_ST16(a_res, rr);
              TMEM_ST16(a_res + 16u, (rr + 16));
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t packed[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 h =
                      __floats2bfloat162_rn(__uint_as_float(rr[c8 * 8 + 2 * j]), __uint_as_float(rr[c8 * 8 + 2 * j + 1]));
                  packed[j] = real ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
                }
                *reinterpret_cast<uint4*>(act + (size_t)((half * 4 + c8) * kActRows + kHalo + p) * 16) =
                    make_uint4(packed[0], packed[1], packed[2], packed[3]);
              }
            } else {
              const float4* hw4 = reinterpret_cast<const float4*>(headw_s + half * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w0 = hw4[q], w1 = hw4[16 + q], w2 = hw4[32 + q];
                const float v0 = __uint_as_float(rr[q * 4]), v1 = __uint_as_float(rr[q * 4 + 1]);
                const float v2 = __uint_as_float(rr[q * 4 + 2]), v3 = __uint_as_float(rr[q * 4 + 3]);
                av = fmaf(v0, w0.x, fmaf(v1, w0.y, fmaf(v2, w0.z, fmaf(v3, w0.w, av))));
                ap0 = fmaf(v0, w1.x, fmaf(v1, w1.y, fmaf(v2, w1.z, fmaf(v3, w1.w, ap0))));
                ap1 = fmaf(v0, w2.x, fmaf(v1, w2.y, fmaf(v2, w2.z, fmaf(v3, w2.w, ap1))));
              }
            }
          }
          if (!last) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_act + t);
          } else if (real) {  // 1x1 head convolutions -> dense features (channel-major like lib/model.py:93)
            featc[(b * 3 + 0) * HW + cell] = lrelu_tc(av + headw_s[192]);
            featc[(b * 3 + 1) * HW + cell] = lrelu_tc(ap0 + headw_s[193]);
            featc[(b * 3 + 2) * HW + cell] = lrelu_tc(ap1 + headw_s[194]);
          }
        }
      }
      tc_fence_before();  // the last layer's accumulators have been read (wait::ld above)
      if (gi + 1 < my_groups) write_inputs((blockIdx.x + (long long)(gi + 1) * gridDim.x) * nb);
      epi_bar_sync();
      // ---- fully connected heads: all (board, output) pairs in parallel, coalesced transposed weights ----
      const int nvalid = (int)min((long long)nb, count - leaf0);
      const int per_board = 20 + gm.A;
      float* hid = fc_s;  // [nb][20] value hidden units
#pragma unroll 1
      for (int o = tid; o < nvalid * per_board; o += kEpiThreads) {
        const int b = o / per_board, i = o - b * per_board;
        const float* feat = featc + (size_t)b * 3 * HW;
        if (i < 20) {
          float a0 = blob[L.val_fc1_b + i], a1 = 0.0f;
          const float* wt = val_fc1_t + i;
          int cell = 0;
          for (; cell + 1 < HW; cell += 2) {
            a0 = fmaf(wt[(size_t)cell * 20], feat[cell], a0);
            a1 = fmaf(wt[(size_t)(cell + 1) * 20], feat[cell + 1], a1);
          }
          if (cell < HW) a0 = fmaf(wt[(size_t)cell * 20], feat[cell], a0);
          hid[b * 20 + i] = lrelu_tc(a0 + a1);
        } else {
          const int a = i - 20;
          float a0 = blob[L.pol_fc_b + a], a1 = 0.0f;
          const float* wt = pol_fc_t + a;
          const float* f2 = feat + HW;
          int k2 = 0;
          for (; k2 + 1 < 2 * HW; k2 += 2) {
            a0 = fmaf(wt[(size_t)k2 * gm.A], f2[k2], a0);
            a1 = fmaf(wt[(size_t)(k2 + 1) * gm.A], f2[k2 + 1], a1);
          }
          if (k2 < 2 * HW) a0 = fmaf(wt[(size_t)k2 * gm.A], f2[k2], a0);
          probs[(size_t)(leaf0 + b) * gm.A + a] = a0 + a1;  // raw logit, normalised below
        }
      }
      __threadfence_block();
      epi_bar_sync();
#pragma unroll 1
      for (int b = warp; b < nvalid; b += 4) {
        const int lane = tid & 31;
        if (lane == 0) {
          float acc = blob[L.val_fc2_b];
          for (int i = 0; i < 20; ++i) acc = fmaf(blob[L.val_fc2_w + i], hid[b * 20 + i], acc);
          values[leaf0 + b] = tanhf(acc);
        }
        float* row = probs + (size_t)(leaf0 + b) * gm.A;  // softmax over all A actions (lib/mcts.py:216)
        float mx = -INFINITY;
#pragma unroll 1
        for (int a = lane; a < gm.A; a += 32) mx = fmaxf(mx, row[a]);
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float sum = 0.0f;
#pragma unroll 1
        for (int a = lane; a < gm.A; a += 32) sum += expf(row[a] - mx);
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll 1
        for (int a = lane; a < gm.A; a += 32) row[a] = expf(row[a] - mx) / sum;
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------ host: packing
static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;  // round to nearest even
  return (uint16_t)(u >> 16);
}

}  // namespace caro

using namespace caro;

int caro_net_tc_pack(caro_net* net, const float* h) {
  const BlobLayout& L = net->layout;
  const size_t img_bytes = (size_t)kLayerBytesIn + (size_t)kBlocks * kLayerBytes;
  std::vector<uint16_t> img(img_bytes / 2, 0);
  // conv_in: [tap][chunk(2)][n=64][8] with only channels 0,1 non-zero
  for (int tap = 0; tap < 9; ++tap)
    for (int co = 0; co < 64; ++co)
      for (int ci = 0; ci < 2; ++ci)
        img[((size_t)tap * kTapBytesIn + (size_t)((ci / 8) * 64 + co) * 16) / 2 + (ci % 8)] =
            f32_to_bf16(h[L.conv_in_w + ((size_t)(co * 2 + ci) * 9 + tap)]);
  for (int l = 0; l < kBlocks; ++l) {
    const size_t base = (size_t)kLayerBytesIn + (size_t)l * kLayerBytes;
    for (int tap = 0; tap < 9; ++tap)
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 64; ++ci)
          img[(base + (size_t)tap * kTapBytes + (size_t)((ci / 8) * 64 + co) * 16) / 2 + (ci % 8)] =
              f32_to_bf16(h[L.conv_w[l] + ((size_t)(co * 64 + ci) * 9 + tap)]);
  }
  std::vector<float> bias((size_t)kNumLayers * 64);
  for (int co = 0; co < 64; ++co) bias[co] = h[L.conv_in_b + co];
  for (int l = 0; l < kBlocks; ++l)
    for (int co = 0; co < 64; ++co) bias[(size_t)(l + 1) * 64 + co] = h[L.conv_b[l] + co];
  const int HW = net->H * net->W, A = net->A;
  // transposed FC weights, policy [2*HW][A] followed by value-FC1 [HW][20]: a warp's outputs read contiguous floats
  std::vector<float> polt((size_t)2 * HW * A + (size_t)HW * 20);
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < 2 * HW; ++i) polt[(size_t)i * A + a] = h[L.pol_fc_w + (size_t)a * 2 * HW + i];
  for (int i = 0; i < 20; ++i)
    for (int c = 0; c < HW; ++c) polt[(size_t)2 * HW * A + (size_t)c * 20 + i] = h[L.val_fc1_w + (size_t)i * HW + c];
  cudaError_t ce = cudaSuccess;
  if (!net->d_tc_weights) ce = cudaMalloc(&net->d_tc_weights, img_bytes);
  if (ce == cudaSuccess && !net->d_tc_bias) ce = cudaMalloc(&net->d_tc_bias, bias.size() * sizeof(float));
  if (ce == cudaSuccess && !net->d_pol_fc_t) ce = cudaMalloc(&net->d_pol_fc_t, polt.size() * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_tc_weights, img.data(), img_bytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_tc_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_pol_fc_t, polt.data(), polt.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

void caro_net_tc_free(caro_net* net) {
  if (net->d_tc_weights) cudaFree(net->d_tc_weights);
  if (net->d_tc_bias) cudaFree(net->d_tc_bias);
  if (net->d_pol_fc_t) cudaFree(net->d_pol_fc_t);
  net->d_tc_weights = nullptr;
  net->d_tc_bias = nullptr;
  net->d_pol_fc_t = nullptr;
}

template <class R>
static int launch_tc(const R& rules, caro_net* net, const void* boards, const uint8_t* who, const int32_t* d_count,
                     int64_t max_count, float* probs, float* values, cudaStream_t st) {
  TcGeom gm;
  gm.H = net->H;
  gm.W = net->W;
  gm.A = net->A;
  gm.pitch = net->W + 1;
  gm.block = (net->H + 1) * gm.pitch;
  gm.boards_per_group = kGroupRows / gm.block;
  if (gm.boards_per_group < 1 || gm.boards_per_group > 32 || gm.pitch + 1 > kHalo)
    return caro_fail(CARO_E_ARG, "board does not fit the tensor-core tile geometry");
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  auto kern = net_tc_kernel<R>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::kTotal);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  const long long max_groups = (max_count + gm.boards_per_group - 1) / gm.boards_per_group;
  const unsigned grid = (unsigned)(max_groups < sm_count ? max_groups : sm_count);
  kern<<<grid, kThreads, TcSmem::kTotal, st>>>(rules, gm, (const typename R::Board*)boards, who, d_count, (long long)max_count,
                                               (const uint8_t*)net->d_tc_weights, net->d_tc_bias, net->d_blob, net->layout,
                                               net->d_pol_fc_t, net->d_pol_fc_t + (size_t)2 * net->H * net->W * net->A, probs, values);
  return caro_check_launch("net_tc_kernel");
}

int caro_net_tc_forward(caro_net* net, int game, int n, int k, const void* d_boards, const uint8_t* d_who,
                        const int32_t* d_count, int64_t max_count, float* d_probs, float* d_values, cudaStream_t st) {
  if (game == CARO_GAME_CONNECT4) return launch_tc<C4Rules>(C4Rules(), net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
  return launch_tc<MnkRules>(MnkRules{n, k}, net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
}
