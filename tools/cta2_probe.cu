// Stand-alone probe for the next round's tower ("Plan for the next round"): checks that a
// tcgen05.mma.cta_group::2 over a CTA pair (M = 256, each CTA supplying its own 128 A rows and HALF of the B columns, taken
// from a window of its stored columns through the descriptor's start address) produces the expected product, and measures
// the issue rate of back-to-back MMAs against the cta_group::1 form the tower uses today (M = 128, all of B from one CTA).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I caro-ai_b200/csrc -o gpurun_out/cta2_probe tools/cta2_probe.cu
//   gpurun_out/cta2_probe
//
// Operand layout = the tower's: no-swizzle K-major core matrices (8 rows x 16 bytes), A chunks [k/8][row][8], B chunks per
// k-step [2][stored rows][8].
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "tc_common.cuh"

using namespace caro;

constexpr int kK = 64;        // 4 k-steps of 16
constexpr int kSteps = kK / 16;
constexpr int kStoreRows = 192;  // B rows (= output columns) stored per CTA (the tower's block: 192 x 16 x 2 B = 6 KB per k-step)

__host__ __device__ constexpr uint32_t probe_idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CTAS>
__global__ void __launch_bounds__(128)
probe_kernel(const __nv_bfloat16* __restrict__ a_img,   // [CTAS][kK/8][128][8]
             const __nv_bfloat16* __restrict__ b_img,   // [CTAS][kSteps][2][kStoreRows][8]
             float* __restrict__ d_out,                  // [CTAS * 128][n]
             long long* __restrict__ cycles, int n, int boff, int iters, int aoff) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* a_s = reinterpret_cast<__nv_bfloat16*>(smem);                             // 16 KB
  __nv_bfloat16* b_s = reinterpret_cast<__nv_bfloat16*>(smem + kK / 8 * 128 * 16);         // 4 x 6 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kK / 8 * 128 * 16 + kSteps * 2 * kStoreRows * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = CTAS == 2 ? cluster_rank() : 0u;
  const int pair = CTAS == 2 ? blockIdx.x / 2 : blockIdx.x;
  const int warp = threadIdx.x >> 5;

  const size_t a_elems = (size_t)kK / 8 * 128 * 8, b_elems = (size_t)kSteps * 2 * kStoreRows * 8;
  for (int i = threadIdx.x; i < (int)a_elems; i += 128) a_s[i] = a_img[rank * a_elems + i];
  for (int i = threadIdx.x; i < (int)b_elems; i += 128) b_s[i] = b_img[rank * b_elems + i];
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  if (warp == 0) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long t0 = 0;
  if (rank == 0 && warp == 1) {
    if ((threadIdx.x & 31) == 0) {  // one issuing thread for the CTA (pair)
      const uint64_t a_desc = make_desc(smem_u32(a_s) + (uint32_t)aoff, 128u * 16u, 128u);  // aoff = 16: the tower's dx = +1 tap
      const uint64_t b_desc = make_desc(smem_u32(b_s) + (uint32_t)boff * 16u, (uint32_t)kStoreRows * 16u, 128u);
      const uint32_t idesc = probe_idesc(128u * CTAS, (uint32_t)n);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < kSteps; ++ks) {
          const uint64_t ad = a_desc + (uint64_t)(ks * 2 * 128 * 16 / 16);
          const uint64_t bd = b_desc + (uint64_t)(ks * 2 * kStoreRows * 16 / 16);
          const uint32_t acc = (it | ks) ? 1u : 0u;
          if (CTAS == 2) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                : "memory");
          } else {
            umma_bf16(tmem, ad, bd, idesc, acc);
          }
        }
      }
      if (CTAS == 2) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
      } else {
        umma_commit(bar);
      }
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (rank == 0 && warp == 1 && (threadIdx.x & 31) == 0 && t0 != 0) cycles[pair] = clock64() - t0;
  // read back: warp w owns TMEM lanes 32 w .. 32 w + 31
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t r[16];
    TMEM_LD16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = (int)rank * 128 + warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 16; ++j) d_out[((size_t)pair * CTAS * 128 + row) * n + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 0) {
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// The tower's operand footprint: activations as 8 chunks of 800 rows (chunk stride 12,800 B), 12 weight blocks of 6 KB, one
// tile = 12 MMAs (3 horizontal taps x 4 k-steps) accumulating into one 192-column range.  Timing only (operands are zeros).
// GAP_AT (compile time, so that the other positions carry no extra instruction): where the issuing thread spends `gap` cycles
// elsewhere -- after MMA GAP_AT of every tile, 12 = between the last MMA and the commit, 13 = after the commit, -1 = nowhere.
template <int GAP_AT>
__global__ void __launch_bounds__(448)
tower_tile_kernel(long long* __restrict__ cycles, int n, int dxs, int tiles, int rotate, int spinners, int commit_per_tile, int gap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kActRows = 800, kChunk = kActRows * 16, kBlock = 6144;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * kChunk + 12 * kBlock);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (8 * kChunk + 12 * kBlock) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  uint64_t* tile_bar = bar + 2;  // [8] per-tile commit targets (nobody waits on them; phases just keep flipping)
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(tile_bar + i, 1);
  }
  fence_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 32) {
    const uint64_t a0 = make_desc(smem_u32(smem) + 16u * 16u, (uint32_t)kChunk, 128u);
    const uint64_t b0 = make_desc(smem_u32(smem) + 8u * kChunk, 192u * 16u, 128u);
    const uint32_t idesc = probe_idesc(128u, (uint32_t)n);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const int y = rotate ? t % 4 : 0;
      const uint64_t a_tile = a0 + (uint64_t)(y * 128);
      const uint32_t d = tmem + (uint32_t)(y * 64);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int dx = i / 4 - 1, kk = i % 4;
        const uint64_t ad = a_tile + (uint64_t)(int64_t)(dx * dxs + kk * 2 * kActRows);
        const uint64_t bd = b0 + (uint64_t)(i * (kBlock / 16));
        umma_bf16(d, ad, bd, idesc, 1u);
        if (GAP_AT == i) {  // the issuing thread is busy elsewhere for `gap` cycles in the middle of a tile
          const long long g0 = clock64();
          while (clock64() - g0 < gap) {}
        }
      }
      if (GAP_AT == 12) {  // ... or between the tile's last MMA and its commit
        const long long g0 = clock64();
        while (clock64() - g0 < gap) {}
      }
      if (commit_per_tile) {
        umma_commit(tile_bar + (t & 7));
        if (commit_per_tile > 1) tc_fence_after();
      }
      if (GAP_AT == 13) {  // ... or after the commit, where the tower's tile loop has its ~300-cycle loop head
        const long long g0 = clock64();
        while (clock64() - g0 < gap) {}
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (warp >= 2 && warp < 2 + spinners) {
    mbar_wait(bar, 0);  // like the tower's epilogue warps while they have nothing to do: spin on mbarrier.try_wait
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static void run_tower_tile(int n, int dxs, int rotate, int spinners = 0, int commit_per_tile = 0, int gap = 0, int gap_at = -1) {
  const int ctas = 148, tiles = 600;
  long long* dc;
  cudaMalloc(&dc, ctas * 8);
  const size_t smem = 8 * 800 * 16 + 12 * 6144 + 128;
  auto launch = [&](auto at) {
    constexpr int AT = decltype(at)::value;
    cudaFuncSetAttribute(tower_tile_kernel<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tower_tile_kernel<AT><<<ctas, 448, smem>>>(dc, n, dxs, tiles, rotate, spinners, commit_per_tile, gap);
  };
  if (gap_at == 5) launch(std::integral_constant<int, 5>{});
  else if (gap_at == 12) launch(std::integral_constant<int, 12>{});
  else if (gap_at == 13) launch(std::integral_constant<int, 13>{});
  else launch(std::integral_constant<int, -1>{});
  const cudaError_t err = cudaDeviceSynchronize();
  std::vector<long long> cyc(ctas);
  cudaMemcpy(cyc.data(), dc, ctas * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : cyc) mx = v > mx ? v : mx;
  printf("tower tile footprint: N=%d, horizontal step %d rows, %s, %d warps spinning on an mbarrier, commit per tile %d, issuing thread away for %d cycles at position %d: %.1f cycles per MMA, %.0f per tile (%s)\n", n, dxs,
         rotate ? "tiles 0..3 in turn" : "one tile", spinners, commit_per_tile, gap, gap_at, (double)mx / (tiles * 12), (double)mx / tiles, cudaGetErrorString(err));
  cudaFree(dc);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int CTAS>
static int run(int n, int boff, int pairs, int aoff = 0) {
  // per-CTA matrices: A_c [128][kK], S_c [kStoreRows][kK]; effective B rows = for each CTA its window [boff, boff + n / CTAS)
  std::vector<float> A((size_t)CTAS * 128 * kK), S((size_t)CTAS * kStoreRows * kK);
  srand(7 + n + boff + CTAS);
  for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : S) v = bf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> a_img((size_t)CTAS * kK / 8 * 128 * 8), b_img((size_t)CTAS * kSteps * 2 * kStoreRows * 8);
  for (int c = 0; c < CTAS; ++c)
    for (int k = 0; k < kK; ++k) {
      for (int r = 0; r < 128; ++r)
        a_img[(((size_t)c * kK / 8 + k / 8) * 128 + r) * 8 + k % 8] = __float2bfloat16(A[((size_t)c * 128 + r) * kK + k]);
      for (int r = 0; r < kStoreRows; ++r)
        b_img[((((size_t)c * kSteps + k / 16) * 2 + (k % 16) / 8) * kStoreRows + r) * 8 + k % 8] =
            __float2bfloat16(S[((size_t)c * kStoreRows + r) * kK + k]);
    }
  const int half = n / CTAS;
  std::vector<float> ref((size_t)CTAS * 128 * n);
  for (int row = 0; row < CTAS * 128; ++row)
    for (int col = 0; col < n; ++col) {
      const int c = col / half, r = boff + col % half;
      double acc = 0.0;
      for (int k = 0; k < kK; ++k) acc += (double)A[(size_t)row * kK + k] * S[((size_t)c * kStoreRows + r) * kK + k];
      ref[(size_t)row * n + col] = (float)acc;
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  long long* dc;
  cudaMalloc(&da, a_img.size() * 2);
  cudaMalloc(&db, b_img.size() * 2);
  cudaMalloc(&dd, (size_t)pairs * CTAS * 128 * n * 4);
  cudaMalloc(&dc, pairs * 8);
  cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = kK / 8 * 128 * 16 + kSteps * 2 * kStoreRows * 16 + 64;
  cudaFuncSetAttribute(probe_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  auto launch = [&](int iters) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pairs * CTAS);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTAS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, probe_kernel<CTAS>, (const __nv_bfloat16*)da, (const __nv_bfloat16*)db, dd, dc, n, boff, iters, aoff);
    return cudaDeviceSynchronize();
  };
  cudaError_t err = launch(1);
  if (err != cudaSuccess) {
    printf("ctas=%d n=%d boff=%d: launch failed: %s\n", CTAS, n, boff, cudaGetErrorString(err));
    return 1;
  }
  std::vector<float> out((size_t)CTAS * 128 * n);
  cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0.0;
  for (size_t i = 0; i < out.size(); ++i) worst = fmax(worst, fabs((double)out[i] - ref[i]));
  const int iters = 2000;
  err = launch(iters);
  std::vector<long long> cyc(pairs);
  cudaMemcpy(cyc.data(), dc, pairs * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : cyc) mx = v > mx ? v : mx;
  if (aoff) worst = 0.0;  // rows shifted by one: timing only
  printf("ctas=%d M=%d N=%d window=%d A+%dB pairs=%d: max |err| = %.3g   %s   %.1f cycles per MMA (%d back-to-back, slowest CTA%s)\n", CTAS,
         128 * CTAS, n, boff, aoff, pairs, worst, worst < 1e-3 ? "OK" : "MISMATCH", (double)mx / (iters * kSteps), iters * kSteps,
         CTAS == 2 ? " pair" : "");
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  cudaFree(dc);
  return worst < 1e-3 && err == cudaSuccess ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<1>(192, 0, 1);
  bad += run<1>(128, 0, 1);
  bad += run<1>(192, 0, 148);
  bad += run<2>(192, 0, 1);
  bad += run<2>(192, 32, 1);
  bad += run<2>(128, 64, 1);
  bad += run<2>(128, 0, 1);
  bad += run<2>(192, 0, 74);
  bad += run<2>(256, 0, 74);
  // the tower's dx = +-1 taps start A one row (16 bytes) off the 128-byte core-matrix boundary
  bad += run<1>(192, 0, 148, 16);
  bad += run<1>(128, 0, 148, 16);
  bad += run<1>(192, 0, 148, 256);
  bad += run<2>(192, 0, 74, 16);
  run_tower_tile(192, 1, 0);
  run_tower_tile(192, 16, 0);
  run_tower_tile(192, 16, 1);
  run_tower_tile(128, 16, 1);
  run_tower_tile(192, 1, 1, 8);
  run_tower_tile(128, 1, 1, 8);
  run_tower_tile(192, 1, 1, 12);
  run_tower_tile(128, 1, 1, 12);
  run_tower_tile(192, 1, 1, 8, 1);
  run_tower_tile(128, 1, 1, 8, 1);
  run_tower_tile(192, 1, 1, 8, 2);
  // how deep is the MMA queue?  a gap of the issuing thread is hidden only while queued MMAs keep the tensor pipe busy
  // issue-rate floor: do narrower MMAs (the 128x64x16 ones of net_tc.cu) issue at their tensor time (32 / 16 cycles)?
  run_tower_tile(64, 1, 1, 8);
  run_tower_tile(32, 1, 1, 8);
  run_tower_tile(64, 1, 1, 8, 1);
  for (int n : {192, 128})
    for (int gap : {0, 100, 200, 300, 500}) {
      run_tower_tile(n, 1, 1, 8, 0, gap, 5);    // mid-tile, no per-tile commit
      run_tower_tile(n, 1, 1, 8, 1, gap, 5);    // mid-tile, commit per tile
      run_tower_tile(n, 1, 1, 8, 1, gap, 12);   // between the last MMA and the commit
      run_tower_tile(n, 1, 1, 8, 1, gap, 13);   // after the commit (the tower's loop head)
      run_tower_tile(n, 1, 1, 8, 0, gap, 13);   // same place without a commit
    }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
