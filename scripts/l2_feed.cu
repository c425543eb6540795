// Micro-measurement (not part of the product library): how many bytes per clock one SM can pull into shared memory with
// TMA tile loads (cp.async.bulk.tensor.2d, box {64 fp16, BR rows}, 128B swizzle, rows 1 KB apart like activation /
// weight rows) when all 148 SMs do it at once - the "L2 -> SM operand feed" that bounds the 1x1 convolutions and the
// token GEMM.  Variants: footprint resident in L2 / streaming from HBM; every CTA its own tiles / all CTAs the same
// tiles (weights); ring depth; and 2-CTA clusters where each CTA loads HALF of every tile and multicasts it to both.
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I laudnet_b200/csrc -I include \
//                              scripts/l2_feed.cu -o /tmp/l2_feed -lcuda && /tmp/l2_feed
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "umma_ptx.cuh"

using namespace laud;

constexpr int MAXS = 12;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(unsigned long long* b, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(b)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* map, unsigned long long* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void mbar_spin(unsigned long long* b, uint32_t parity) {     // non-suspending poll
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(b);
  long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) __trap();
  }
}
#define WAIT(b, p) do { if (mode & 2) mbar_spin(b, p); else mbar_wait(b, p); } while (0)

// mode bit0: all CTAs read the same tiles; bit1: spin on test_wait instead of try_wait; bit2: no consumer round trip
// (the producer never waits: every tile lands in stage i % stages, completion counted at the end); csize: cluster size (1 | 2; 2 = half tiles multicast to both CTAs)
__global__ void __launch_bounds__(192, 1)
feed_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int stages, int tiles_per_cta, int iters, int mode, int csize,
            long long* out, int chunks, int producers) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ unsigned long long full[MAXS], empty[MAXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = csize > 1 ? cluster_ctarank() : 0;
  const int tile_bytes = box_rows * 128 * chunks;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], csize); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  const int group = csize > 1 ? blockIdx.x / csize : blockIdx.x;        // CTAs of a cluster read the same tiles
  const long long t0 = clock64();
  if (warp >= 2 && warp < 2 + producers && lane == 0 && producers > 1) {
    // several producer threads (one per warp), tile i issued by producer i % producers
    const int me = warp - 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (i % producers == me) {
        const int tile = ((mode & 1) ? 0 : group * tiles_per_cta) + i % tiles_per_cta;
        WAIT(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
        tma_load_3d(base + stage * tile_bytes, &map, &full[stage], 0, tile * box_rows, 0);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 0 && producers == -2) {
    // converged warp loop: lane 0 issues the first half of the tile's rows, lane 1 the second half (two instructions per stage)
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      const int tile = ((mode & 1) ? 0 : group * tiles_per_cta) + i % tiles_per_cta;
      WAIT(&empty[stage], phase ^ 1u);
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
      __syncwarp();
      if (lane < 2) tma_load_3d(base + stage * tile_bytes + lane * (tile_bytes / 2), &map, &full[stage], 0, tile * box_rows + lane * (box_rows / 2), 0);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 0 && lane == 0 && producers == -3) {
    // one thread, two instructions per stage (the convolution's pattern: activation tile + weight tile)
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      const int tile = ((mode & 1) ? 0 : group * tiles_per_cta) + i % tiles_per_cta;
      WAIT(&empty[stage], phase ^ 1u);
      mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
      tma_load_3d(base + stage * tile_bytes, &map, &full[stage], 0, tile * box_rows, 0);
      tma_load_3d(base + stage * tile_bytes + tile_bytes / 2, &map, &full[stage], 0, tile * box_rows + box_rows / 2, 0);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 2 && warp < 4 && lane == 0 && producers == -4) {
    // two threads in two warps, BOTH issue for every stage (half a tile each); only the first does the expect_tx
    const int me = warp - 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      const int tile = ((mode & 1) ? 0 : group * tiles_per_cta) + i % tiles_per_cta;
      WAIT(&empty[stage], phase ^ 1u);
      if (me == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
      tma_load_3d(base + stage * tile_bytes + me * (tile_bytes / 2), &map, &full[stage], 0, tile * box_rows + me * (box_rows / 2), 0);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 0 && lane == 0 && producers == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      const int tile = ((mode & 1) ? 0 : group * tiles_per_cta) + i % tiles_per_cta;
      if (!(mode & 4)) WAIT(&empty[stage], phase ^ 1u);
      mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
      if (csize == 1) {
        tma_load_3d(base + stage * tile_bytes, &map, &full[stage], 0, tile * box_rows, 0);
      } else {                                                         // my half of the rows, delivered to both CTAs
        const int half_rows = box_rows / 2;
        tma_load_2d_mc(base + stage * tile_bytes + rank * half_rows * 128, &map, &full[stage], 0, tile * box_rows + rank * half_rows,
                       (uint16_t)0x3);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      WAIT(&full[stage], phase);
      if (csize == 1) mbar_arrive(&empty[stage]);
      else { mbar_arrive_cluster(&empty[stage], 0); mbar_arrive_cluster(&empty[stage], 1); }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaSetDevice(0);
  EncodeTiledFn enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn)p;
  }
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const long long PITCH = 512;                           // halves: rows 1 KB apart
  const long long ROWS = 148LL * 256 * 128;              // enough rows for 256 tiles of 128 rows per CTA (4.6 GB span)
  __half* d;
  if (cudaMalloc(&d, ROWS * PITCH * 2) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMemset(d, 0, ROWS * PITCH * 2);
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("# TMA tile loads into shared memory, all 148 SMs at once: bytes per clock per SM (slowest CTA) and aggregate TB/s at the SM clock\n");
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  struct Cfg { int csize, mode, box_rows, chunks, stages, tiles, producers; };
  const Cfg cfgs[] = {
      {1, 0, 128, 1, 4, 4, 1}, {1, 0, 256, 1, 4, 4, 1},
      // per-stage patterns, tile = 256 rows (32 KB) or 384 rows (48 KB) split in two instructions
      {1, 0, 256, 1, 4, 4, -3}, {1, 0, 256, 1, 4, 4, -2}, {1, 0, 256, 1, 4, 4, -4},
      {1, 0, 512, 1, 3, 4, -3}, {1, 0, 512, 1, 3, 4, -2}, {1, 0, 512, 1, 3, 4, -4},
      {1, 0, 256, 2, 2, 4, -3}, {1, 0, 256, 2, 2, 4, -4},
  };
  for (const Cfg& c : cfgs) {
    const int csize = c.csize, mode = c.mode, box_rows = c.box_rows, stages = c.stages, tiles = c.tiles;
    CUtensorMap map;
    cuuint64_t gdim[3] = {64, (cuuint64_t)ROWS, 8}, gstr[2] = {(cuuint64_t)PITCH * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)((csize > 1 || c.producers < 0) ? box_rows / 2 : box_rows), (cuuint32_t)c.chunks}, es[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      printf("encode failed\n");
      continue;
    }
    const int iters = 2048;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = (size_t)stages * box_rows * 128 * c.chunks + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {           // first repetition warms L2
      cudaLaunchKernelEx(&cfg, feed_kernel, map, box_rows, stages, tiles, iters, mode, csize, d_out, c.chunks, c.producers);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    double sum = 0;
    for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; sum += (double)h[i]; }
    const double bytes = (double)iters * box_rows * 128 * c.chunks;
    printf("cluster=%d box={64 ch, %3d rows, %d chunks} = %2d KB per TMA instruction, %d stages, %d producer thread(s), %s: %6.1f B/clk/SM "
           "(slowest CTA; mean %.1f), %.0f cycles per instruction, %5.2f TB/s delivered\n",
           csize, box_rows, c.chunks, box_rows * c.chunks / 8, stages, c.producers, tiles == 4 ? "L2-resident" : "from HBM   ",
           bytes / (double)mx, bytes * 148 / sum, (double)mx / iters, bytes * 148 / ((double)mx / (clk_khz * 1e3)) / 1e12);
  }
  return 0;
}
