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
constexpr int kHalo = 32;                                // zero positions before / after (>= pitch + 1)
constexpr int kActRows = kGroupRows + 2 * kHalo;         // 576
constexpr int kChunkBytes = kActRows * 16;               // one 8-channel chunk of all positions
constexpr int kActBytes = 8 * kChunkBytes;               // 73,728
constexpr int kTapBytes = 8 * 64 * 16;                   // 8,192: one tap of a 64->64 layer
constexpr int kTapBytesIn = 2 * 64 * 16;                 // 2,048: one tap of conv_in (K padded to 16)
constexpr int kLayerBytes = 9 * kTapBytes;               // 73,728
constexpr int kLayerBytesIn = 9 * kTapBytesIn;           // 18,432
constexpr int kNumLayers = 1 + kBlocks;                  // conv_in + 5 residual blocks
constexpr int kThreads = 128;
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

__device__ __forceinline__ float lrelu_tc(float x) { return x > 0.0f ? x : kLeaky * x; }

struct TcSmem {
  // dynamic shared memory carve-up (byte offsets from a 128-aligned base)
  static constexpr int kAct = 0;
  static constexpr int kWgt = kAct + kActBytes;
  static constexpr int kBias = kWgt + kLayerBytes;                 // float [6][64]
  static constexpr int kHeadW = kBias + kNumLayers * 64 * 4;       // float [3][64] + [3] biases (+pad)
  static constexpr int kHeadF = kHeadW + 4 * 64 * 4;               // float [512][3] head features
  static constexpr int kFc = kHeadF + kGroupRows * 3 * 4;          // float scratch: hidden[32 boards][20] / logits
  static constexpr int kBars = kFc + 32 * 20 * 4 + 2 * 256 * 4;    // 2 mbarriers + tmem base
  static constexpr int kTotal = kBars + 64;
};

// ------------------------------------------------------------------------------------- kernel
template <class R>
__global__ void __launch_bounds__(kThreads, 1)
net_tc_kernel(R rules, TcGeom gm, const typename R::Board* __restrict__ boards, const uint8_t* __restrict__ who,
              const int32_t* __restrict__ d_count, long long max_count, const uint8_t* __restrict__ wimg,
              const float* __restrict__ bias_g, const float* __restrict__ blob, BlobLayout L,
              const float* __restrict__ pol_fc_t, float* __restrict__ probs, float* __restrict__ values) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* act = smem + TcSmem::kAct;
  uint8_t* wgt = smem + TcSmem::kWgt;
  float* bias_s = reinterpret_cast<float*>(smem + TcSmem::kBias);
  float* headw_s = reinterpret_cast<float*>(smem + TcSmem::kHeadW);
  float* headf_s = reinterpret_cast<float*>(smem + TcSmem::kHeadF);
  float* fc_s = reinterpret_cast<float*>(smem + TcSmem::kFc);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + TcSmem::kBars);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const long long count = d_count ? min((long long)*d_count, max_count) : max_count;
  const int nb = gm.boards_per_group;
  const long long n_groups = (count + nb - 1) / nb;
  if ((long long)blockIdx.x >= n_groups) return;  // uniform per CTA, before any barrier / TMEM use

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
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t act_addr = smem_u32(act);
  const uint32_t wgt_addr = smem_u32(wgt);
  uint32_t ph_w = 0, ph_mma = 0;

  // per-thread geometry of its 4 rows (row p = 128*t + tid)
  int row_board[kTilesPerGroup], row_cell[kTilesPerGroup];  // board index in group, r*W+c (or -1 if padding)
#pragma unroll
  for (int t = 0; t < kTilesPerGroup; ++t) {
    const int p = t * kTileRows + tid;
    const int b = p / gm.block, within = p - b * gm.block;
    const int r = within / gm.pitch, c = within - r * gm.pitch;
    const bool real = b < nb && r < gm.H && c < gm.W;
    row_board[t] = b;
    row_cell[t] = real ? r * gm.W + c : -1;
  }

  // first layer's weights for the first group
  if (tid == 0) {
    mbar_expect_tx(bar_w, kLayerBytesIn);
    for (int tap = 0; tap < 9; ++tap) bulk_g2s(wgt + tap * kTapBytesIn, wimg + tap * kTapBytesIn, kTapBytesIn, bar_w);
  }

  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long leaf0 = grp * nb;
    // ---- input planes -> bf16 activations, channels 0/1 (chunk 0), chunk 1 = zeros ---------
#pragma unroll
    for (int t = 0; t < kTilesPerGroup; ++t) {
      const int p = t * kTileRows + tid;
      uint32_t lo = 0u;
      const long long leaf = leaf0 + row_board[t];
      if (row_cell[t] >= 0 && leaf < count) {
        const typename R::Board s = boards[leaf];
        const int wm = who[leaf];
        const int r = row_cell[t] / gm.W, c = row_cell[t] - r * gm.W;
        const uint32_t mine = rules.plane_value(s, wm, 0, r, c) ? 0x3F80u : 0u;   // bf16(1.0)
        const uint32_t other = rules.plane_value(s, wm, 1, r, c) ? 0x3F80u : 0u;
        lo = mine | (other << 16);
      }
      *reinterpret_cast<uint4*>(act + (size_t)(0 * kActRows + kHalo + p) * 16) = make_uint4(lo, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(act + (size_t)(1 * kActRows + kHalo + p) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    float hv[kTilesPerGroup], hp0[kTilesPerGroup], hp1[kTilesPerGroup];
    for (int layer = 0; layer < kNumLayers; ++layer) {
      const int ksteps = layer == 0 ? 1 : 4;
      const int tap_bytes = layer == 0 ? kTapBytesIn : kTapBytes;
      // ---- MMA issue (one thread) ---------------------------------------------------------
      if (tid == 0) {
        mbar_wait(bar_w, ph_w);
        tc_fence_after();
        for (int t = 0; t < kTilesPerGroup; ++t) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * 64);
          uint32_t acc = 0u;
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * gm.pitch + (tap % 3 - 1);
            const uint32_t a_row = (uint32_t)(kHalo + t * kTileRows + shift);
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t adesc = make_desc(act_addr + (uint32_t)((2 * kk) * kActRows + a_row) * 16u, kChunkBytes, 128u);
              const uint64_t bdesc = make_desc(wgt_addr + (uint32_t)(tap * tap_bytes + kk * 2048), 1024u, 128u);
              umma_bf16(d_tmem, adesc, bdesc, kIdesc, acc);
              acc = 1u;
            }
          }
        }
        umma_commit(bar_mma);
      }
      ph_w ^= 1u;
      // ---- wait for the accumulators ------------------------------------------------------
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1u;
      tc_fence_after();
      // weights buffer is free again: stream in the next layer (or next group's conv_in)
      if (tid == 0) {
        const int nl = layer + 1;
        const bool more_groups = grp + gridDim.x < n_groups;
        if (nl < kNumLayers) {
          mbar_expect_tx(bar_w, kLayerBytes);
          const uint8_t* src = wimg + kLayerBytesIn + (size_t)(nl - 1) * kLayerBytes;
          for (int tap = 0; tap < 9; ++tap) bulk_g2s(wgt + tap * kTapBytes, src + tap * kTapBytes, kTapBytes, bar_w);
        } else if (more_groups) {
          mbar_expect_tx(bar_w, kLayerBytesIn);
          for (int tap = 0; tap < 9; ++tap) bulk_g2s(wgt + tap * kTapBytesIn, wimg + tap * kTapBytesIn, kTapBytesIn, bar_w);
        }
      }
      __syncwarp();  // tcgen05.ld/st are .sync.aligned: re-converge after the single-thread branches / spin waits
      // ---- epilogue: bias + LeakyReLU (+ residual), fp32 stream -> TMEM, bf16 copy -> smem ---
      const bool last = layer == kNumLayers - 1;
      const float* bl = bias_s + layer * 64;
#pragma unroll
      for (int t = 0; t < kTilesPerGroup; ++t) {
        const int p = t * kTileRows + tid;
        const bool real = row_cell[t] >= 0;
        float av = 0.0f, ap0 = 0.0f, ap1 = 0.0f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
          const uint32_t a_acc = tmem_base + lane_base + (uint32_t)(t * 64 + ch * 16);
          const uint32_t a_res = tmem_base + lane_base + (uint32_t)(256 + t * 64 + ch * 16);
          uint32_t ra[16], rr[16];
          TMEM_LD16(a_acc, ra);
          if (layer > 0) TMEM_LD16(a_res, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = lrelu_tc(__uint_as_float(ra[j]) + bl[ch * 16 + j]);
            if (layer > 0) v += __uint_as_float(rr[j]);
            rr[j] = __float_as_uint(v);
            if (last) {
              av = fmaf(v, headw_s[ch * 16 + j], av);
              ap0 = fmaf(v, headw_s[64 + ch * 16 + j], ap0);
              ap1 = fmaf(v, headw_s[128 + ch * 16 + j], ap1);
            }
          }
          if (!last) {
            TMEM_ST16(a_res, rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(rr[2 * j]), __uint_as_float(rr[2 * j + 1]));
              packed[j] = real ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
            }
            *reinterpret_cast<uint4*>(act + (size_t)((2 * ch) * kActRows + kHalo + p) * 16) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint4*>(act + (size_t)((2 * ch + 1) * kActRows + kHalo + p) * 16) =
                make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
        hv[t] = av;
        hp0[t] = ap0;
        hp1[t] = ap1;
      }
      if (!last) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }

    // ---- heads ------------------------------------------------------------------------------
    const int HW = gm.H * gm.W;
#pragma unroll
    for (int t = 0; t < kTilesPerGroup; ++t) {
      const int p = t * kTileRows + tid;
      headf_s[p * 3 + 0] = lrelu_tc(hv[t] + headw_s[192]);
      headf_s[p * 3 + 1] = lrelu_tc(hp0[t] + headw_s[193]);
      headf_s[p * 3 + 2] = lrelu_tc(hp1[t] + headw_s[194]);
    }
    __syncthreads();
    const int nvalid = (int)min((long long)nb, count - leaf0);
    float* hid = fc_s;                 // [nb][20]
    float* logit = fc_s + 32 * 20;     // [<=512] logits, processed board by board
    // value head FC1 (HW -> 20) for all boards of the group
    for (int o = tid; o < nvalid * 20; o += kThreads) {
      const int b = o / 20, i = o - b * 20;
      float acc = blob[L.val_fc1_b + i];
      const float* wrow = blob + L.val_fc1_w + (size_t)i * HW;
      for (int cell = 0; cell < HW; ++cell) {
        const int r = cell / gm.W, c = cell - r * gm.W;
        acc = fmaf(wrow[cell], headf_s[(b * gm.block + r * gm.pitch + c) * 3], acc);
      }
      hid[o] = lrelu_tc(acc);
    }
    __syncthreads();
    for (int b = tid; b < nvalid; b += kThreads) {
      float acc = blob[L.val_fc2_b];
      for (int i = 0; i < 20; ++i) acc = fmaf(blob[L.val_fc2_w + i], hid[b * 20 + i], acc);
      values[leaf0 + b] = tanhf(acc);
    }
    // policy head FC (2*HW -> A) + softmax, one board at a time (A can be 225)
    for (int b = 0; b < nvalid; ++b) {
      for (int a = tid; a < gm.A; a += kThreads) {
        float acc = blob[L.pol_fc_b + a];
        for (int chn = 0; chn < 2; ++chn)
          for (int cell = 0; cell < HW; ++cell) {
            const int r = cell / gm.W, c = cell - r * gm.W;
            acc = fmaf(pol_fc_t[(size_t)(chn * HW + cell) * gm.A + a], headf_s[(b * gm.block + r * gm.pitch + c) * 3 + 1 + chn], acc);
          }
        logit[a] = acc;
      }
      __syncthreads();
      if (warp == 0) {
        const int lane = tid & 31;
        float mx = -INFINITY;
        for (int a = lane; a < gm.A; a += 32) mx = fmaxf(mx, logit[a]);
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float sum = 0.0f;
        for (int a = lane; a < gm.A; a += 32) sum += expf(logit[a] - mx);
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        for (int a = lane; a < gm.A; a += 32) probs[(size_t)(leaf0 + b) * gm.A + a] = expf(logit[a] - mx) / sum;
      }
      __syncthreads();
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  std::vector<float> polt((size_t)2 * HW * A);
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < 2 * HW; ++i) polt[(size_t)i * A + a] = h[L.pol_fc_w + (size_t)a * 2 * HW + i];
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
                                               net->d_pol_fc_t, probs, values);
  return caro_check_launch("net_tc_kernel");
}

int caro_net_tc_forward(caro_net* net, int game, int n, int k, const void* d_boards, const uint8_t* d_who,
                        const int32_t* d_count, int64_t max_count, float* d_probs, float* d_values, cudaStream_t st) {
  if (game == CARO_GAME_CONNECT4) return launch_tc<C4Rules>(C4Rules(), net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
  return launch_tc<MnkRules>(MnkRules{n, k}, net, d_boards, d_who, d_count, max_count, d_probs, d_values, st);
}
