// Fully connected heads of large boards (Caro 15 x 15: 450 x 225 policy FC, 225 x 20 value FC1) on the tensor cores.
//
// lib/model.py:56-72,90-93 + the softmax of lib/mcts.py:216 for the leaves of one network pass.  The tap-per-MMA tower
// (net_tc.cu) exports, per leaf, the ACTIVATED outputs of its 1x1 head convolutions as bf16 hi + lo pairs; this kernel
// evaluates both FC layers as ONE GEMM per tile of 128 leaves,
//      D[128 leaves][256] = F[128][K] x Wt[K][256],   K = 3 HW (value plane, policy plane 0, policy plane 1; padded to 64)
// with a block-structured B: columns 0 .. A-1 are the policy FC (rows HW .. 3HW-1), columns A .. A+19 the value FC1
// (rows 0 .. HW-1), everything else zero.  Features AND weights are split into bf16 hi + lo and every product is
// hi hi + lo hi + hi lo (fp32 accumulation in TMEM): fp32-class accuracy (the value head's tolerance is 1e-3 and a
// single bf16 pass over K = 225 .. 450 is too close to it), at 3 x a tensor time that is negligible anyway --
// 33 MMAs of 128 x 256 x 16 per 64-wide K chunk, 11 chunks: ~17 k cycles = 9 us per 128 leaves, one wave for 16 k leaves.
// The SIMT kernel this replaces (32 leaves per CTA, FMA from shared memory) took 139 us of the 522 us of a 4,096-leaf
// Caro pass (profiles/r1_caro_net_ncu_summary.txt).
//
// Pipeline per CTA (persistent over tiles): warp 5 streams the K chunks of A (16 + 16 KB, hi + lo) and B (32 + 32 KB)
// with cp.async.bulk into a 2-stage ring, warp 4 issues the MMAs (one elected lane) into one of two 256-column TMEM
// accumulators, warps 0-3 (thread = leaf = TMEM lane) add the biases, do the softmax over ALL A actions and the value
// FC2 + tanh, and write the priors through a shared-memory transpose so that the global stores are coalesced.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/caro_b200.h"
#include "common_host.h"
#include "net.h"
#include "tc_common.cuh"

namespace caro {

constexpr int kHtRows = 128;                       // leaves per tile
constexpr int kHtN = 256;                          // GEMM N (A + 20 <= 256)
constexpr int kHtKChunk = 64;                      // K elements per pipeline stage
constexpr int kHtStages = 2;
constexpr int kHtAChunkBytes = 8 * kHtRows * 16;   // 16,384: [8 k-chunks of 8][128 rows][8 bf16]
constexpr int kHtBChunkBytes = 8 * kHtN * 16;      // 32,768: [8][256][8]
constexpr int kHtStageBytes = 2 * kHtAChunkBytes + 2 * kHtBChunkBytes;  // 98,304
constexpr int kHtStagePitch = 33;                  // floats per row of the output transpose buffer
constexpr int kHtEpiThreads = 128;
constexpr int kHtThreads = kHtEpiThreads + 64;
constexpr int kHtSmemStage = kHtStages * kHtStageBytes;                  // 196,608
constexpr int kHtSmemOut = kHtSmemStage;                                // float [128][33]
constexpr int kHtSmemVec = kHtSmemOut + kHtRows * kHtStagePitch * 4;    // policy bias [A] | fc1 bias [20] | fc2 w [20] | fc2 b
constexpr int kHtSmemBars = kHtSmemVec + (256 + 48) * 4;
constexpr int kHtSmemTotal = kHtSmemBars + 16 * 8;
static_assert(kHtSmemTotal <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
constexpr uint32_t kHtIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// a_img: [tiles][KC8][128][8] bf16 hi image, the lo image `a_lo_off` bytes behind it; b_img: [KC64][{hi, lo}][8][256][8].
__global__ void __launch_bounds__(kHtThreads, 1)
heads_tc_kernel(const uint8_t* __restrict__ a_img, size_t a_lo_off, const uint8_t* __restrict__ b_img, int kc64, int A,
                const int32_t* __restrict__ d_count, long long max_count, const float* __restrict__ blob, BlobLayout L,
                float* __restrict__ probs, float* __restrict__ values) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* out_s = reinterpret_cast<float*>(smem + kHtSmemOut);
  float* vec_s = reinterpret_cast<float*>(smem + kHtSmemVec);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kHtSmemBars);  // [stages] chunk landed
  uint64_t* bar_empty = bar_full + kHtStages;                              // [stages] chunk consumed
  uint64_t* bar_acc_full = bar_empty + kHtStages;                          // [2] accumulator complete
  uint64_t* bar_acc_empty = bar_acc_full + 2;                              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const long long count = d_count ? min((long long)*d_count, max_count) : max_count;
  const long long tiles = (count + kHtRows - 1) / kHtRows;
  if ((long long)blockIdx.x >= tiles) return;
  const int my_tiles = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const size_t tile_bytes = (size_t)kc64 * kHtAChunkBytes;  // one tile of one A image

  for (int i = tid; i < A + 41; i += kHtThreads)
    vec_s[i < A ? i : 256 + (i - A)] = i < A ? blob[L.pol_fc_b + i]
                                       : i < A + 20 ? blob[L.val_fc1_b + i - A]
                                       : i < A + 40 ? blob[L.val_fc2_w + i - A - 20] : blob[L.val_fc2_b];
  if (tid == 0) {
    for (int s = 0; s < kHtStages; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + b, 1);
      mbar_init(bar_acc_empty + b, kHtEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ===================== loader ============================================================
    if ((tid & 31) == 0) {
      uint32_t n = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const long long tile = blockIdx.x + (long long)ti * gridDim.x;
        const uint8_t* a_hi = a_img + (size_t)tile * tile_bytes;
        for (int c = 0; c < kc64; ++c, ++n) {
          const int s = (int)(n % kHtStages);
          if (n >= (uint32_t)kHtStages) mbar_wait(bar_empty + s, ((n / kHtStages) - 1u) & 1u);
          uint8_t* dst = smem + (size_t)s * kHtStageBytes;
          mbar_expect_tx(bar_full + s, (uint32_t)kHtStageBytes);
          bulk_g2s(dst, a_hi + (size_t)c * kHtAChunkBytes, kHtAChunkBytes, bar_full + s);
          bulk_g2s(dst + kHtAChunkBytes, a_hi + a_lo_off + (size_t)c * kHtAChunkBytes, kHtAChunkBytes, bar_full + s);
          bulk_g2s(dst + 2 * kHtAChunkBytes, b_img + (size_t)c * 2 * kHtBChunkBytes, 2 * kHtBChunkBytes, bar_full + s);
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer ========================================================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    uint32_t n = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int buf = ti & 1;
      if (ti >= 2) mbar_wait(bar_acc_empty + buf, (uint32_t)((ti >> 1) - 1) & 1u);  // the epilogue has drained this accumulator
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
      for (int c = 0; c < kc64; ++c, ++n) {
        const int s = (int)(n % kHtStages);
        mbar_wait(bar_full + s, (n / kHtStages) & 1u);
        tc_fence_after();
        if (elected) {
          const uint32_t base = smem_u32(smem + (size_t)s * kHtStageBytes);
          const uint64_t a_hi = make_desc(base, 2048u, 128u), a_lo = make_desc(base + kHtAChunkBytes, 2048u, 128u);
          const uint64_t b_hi = make_desc(base + 2 * kHtAChunkBytes, 4096u, 128u);
          const uint64_t b_lo = make_desc(base + 2 * kHtAChunkBytes + kHtBChunkBytes, 4096u, 128u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ao = (uint64_t)(ks * 2 * (2048 / 16)), bo = (uint64_t)(ks * 2 * (4096 / 16));
            umma_bf16(d_tmem, a_hi + ao, b_hi + bo, kHtIdesc, (c | ks) ? 1u : 0u);
            umma_bf16(d_tmem, a_lo + ao, b_hi + bo, kHtIdesc, 1u);
            umma_bf16(d_tmem, a_hi + ao, b_lo + bo, kHtIdesc, 1u);
          }
          umma_commit(bar_empty + s);
          if (c == kc64 - 1) umma_commit(bar_acc_full + buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: thread = leaf ===========================================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float* polb = vec_s;
    const float* fc1b = vec_s + 256;
    const float* fc2w = vec_s + 276;
    const float fc2b = vec_s[296];
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int buf = ti & 1;
      const long long tile = blockIdx.x + (long long)ti * gridDim.x;
      const long long leaf = tile * kHtRows + tid;
      mbar_wait(bar_acc_full + buf, (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + (uint32_t)(buf * 256);
      uint32_t r[16];
      // pass 1: the row maximum of the logits
      float mx = -INFINITY;
      for (int c0 = 0; c0 < A; c0 += 16) {
        TMEM_LD16(acc + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < A) mx = fmaxf(mx, __uint_as_float(r[j]) + polb[c0 + j]);
      }
      // pass 2: the normaliser
      float sum = 0.0f;
      for (int c0 = 0; c0 < A; c0 += 16) {
        TMEM_LD16(acc + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < A) sum += expf(__uint_as_float(r[j]) + polb[c0 + j] - mx);
      }
      // value head: FC1 bias + LeakyReLU, FC2, tanh (columns A .. A+19; may straddle 16-column groups)
      {
        float v = fc2b;
        const int c_lo = A & ~15;
        for (int c0 = c_lo; c0 < A + 20; c0 += 16) {
          TMEM_LD16(acc + (uint32_t)c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int i = c0 + j - A;
            if (i >= 0 && i < 20) v = fmaf(fc2w[i], lrelu_tc(__uint_as_float(r[j]) + fc1b[i]), v);
          }
        }
        if (leaf < count) values[leaf] = tanhf(v);
      }
      // pass 3: priors, 32 columns at a time through the transpose buffer -> coalesced rows
      const float inv = 1.0f / sum;
      for (int c0 = 0; c0 < A; c0 += 32) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          TMEM_LD16(acc + (uint32_t)(c0 + 16 * h), r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c0 + 16 * h + j;
            out_s[tid * kHtStagePitch + 16 * h + j] = col < A ? expf(__uint_as_float(r[j]) + polb[col] - mx) * inv : 0.0f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kHtEpiThreads) : "memory");
        const int lane = tid & 31;
        if (c0 + lane < A) {
#pragma unroll 4
          for (int rr = warp * 32; rr < warp * 32 + 32; ++rr) {
            const long long lf = tile * kHtRows + rr;
            if (lf < count) probs[(size_t)lf * A + c0 + lane] = out_s[rr * kHtStagePitch + lane];
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kHtEpiThreads) : "memory");
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static uint16_t ht_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

}  // namespace caro

using namespace caro;

int caro_net_heads_kc64(const caro_net* net) { return (3 * net->H * net->W + kHtKChunk - 1) / kHtKChunk; }

bool caro_net_heads_supported(const caro_net* net) { return net->A + 20 <= kHtN; }

// B image: [KC64][{hi, lo}][8 k-chunks of 8][256 n][8] bf16 (UMMA K-major, no swizzle): B(n, k) = policy FC weight
// W[n][k - HW] for n < A, k in [HW, 3 HW); value FC1 weight W1[n - A][k] for n in [A, A + 20), k < HW; zero elsewhere.
int caro_net_heads_pack(caro_net* net, const float* h) {
  const BlobLayout& L = net->layout;
  const int HW = net->H * net->W, A = net->A, kc64 = caro_net_heads_kc64(net);
  const size_t bytes = (size_t)kc64 * 2 * kHtBChunkBytes;
  std::vector<uint16_t> img(bytes / 2, 0);
  auto put = [&](int n, int k, float w) {
    const uint16_t hi = ht_f32_to_bf16(w);
    uint32_t hb = (uint32_t)hi << 16;
    float hf;
    memcpy(&hf, &hb, 4);
    const uint16_t lo = ht_f32_to_bf16(w - hf);
    const size_t off = (size_t)(k / kHtKChunk) * 2 * kHtBChunkBytes + (size_t)(((k % kHtKChunk) / 8) * kHtN + n) * 16 + (size_t)(k % 8) * 2;
    img[off / 2] = hi;
    img[(off + kHtBChunkBytes) / 2] = lo;
  };
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < 2 * HW; ++i) put(a, HW + i, h[L.pol_fc_w + (size_t)a * 2 * HW + i]);
  for (int i = 0; i < 20; ++i)
    for (int c = 0; c < HW; ++c) put(A + i, c, h[L.val_fc1_w + (size_t)i * HW + c]);
  cudaError_t ce = cudaSuccess;
  if (!net->d_heads_b) ce = cudaMalloc(&net->d_heads_b, bytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(net->d_heads_b, img.data(), bytes, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

void caro_net_heads_free(caro_net* net) {
  if (net->d_heads_b) cudaFree(net->d_heads_b);
  net->d_heads_b = nullptr;
}

int caro_net_heads_prepare() {
  const cudaError_t ce = cudaFuncSetAttribute(heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHtSmemTotal);
  if (ce != cudaSuccess) return caro_fail(CARO_E_CUDA, cudaGetErrorString(ce));
  return CARO_OK;
}

// `a_img`: the hi image of this launch's slot ([tiles][KC8][128][8] bf16), the lo image `a_lo_off` bytes behind it.
int caro_net_heads_forward(caro_net* net, const void* a_img, size_t a_lo_off, const int32_t* d_count, int64_t max_count,
                           float* d_probs, float* d_values, cudaStream_t st) {
  const long long tiles = (max_count + kHtRows - 1) / kHtRows;
  const unsigned grid = (unsigned)(tiles < net->sm_count ? tiles : net->sm_count);
  heads_tc_kernel<<<grid, kHtThreads, kHtSmemTotal, st>>>((const uint8_t*)a_img, a_lo_off, (const uint8_t*)net->d_heads_b,
                                                          caro_net_heads_kc64(net), net->A, d_count, (long long)max_count, net->d_blob,
                                                          net->layout, d_probs, d_values);
  return caro_check_launch("heads_tc_kernel");
}
