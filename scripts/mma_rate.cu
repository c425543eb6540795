// Micro-measurements that decide kernel design (not part of the product library):
//   1. cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, SS operands) as a function of N, with the A
//      descriptor 1024-byte aligned and row-shifted by whole 128-byte rows (the halo-mode tap views), and with one or
//      two alternating accumulators;
//   2. what cp.async.bulk.tensor.2d ... tile::gather4 delivers (row order, swizzle placement) for a [rows, 64] fp16
//      matrix, box {64, 1} and {64, 4}.
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I laudnet_b200/csrc -I include \
//                              scripts/mma_rate.cu -o /tmp/mma_rate -lcuda && /tmp/mma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "umma_ptx.cuh"

using namespace laud;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, int a_shift_rows, int two_acc, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  // A: 2 x (128 rows x 128 B) + slack for shifted starts; B: 256 rows x 128 B.  Contents: zeros (timing only).
  for (int i = threadIdx.x; i < (64 * 1024 + 32 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t a_base = smem_u32(smem) + (uint32_t)a_shift_rows * 128u;
    const uint32_t b_base = smem_u32(smem) + 64 * 1024;
    const uint32_t idesc = umma_idesc_f16(n, 0);
    const uint64_t ad = umma_desc(a_base, 16, 1024), bd = umma_desc(b_base, 16, 1024);
    // warm-up
    for (int i = 0; i < 64; ++i) umma_f16_elect(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i ? 1u : 0u);
    umma_commit_elect(&bar);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      umma_f16_elect(tmem + ((two_acc && (i & 4)) ? 256u : 0u), ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
    const long long t1 = clock64();
    umma_commit_elect(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;   // issue
      out[1] = t2 - t0;   // completion
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------- gather4
__global__ void gather4_kernel(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int col, int dst_off,
                               __half* out /* 1024 halfs = 2 KB */, int* status) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 2048 / 2; i += blockDim.x) reinterpret_cast<__half*>(smem)[i] = __float2half(-1.f);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 4 * 128);
    tma_gather4(smem_u32(smem) + dst_off, &map, &bar, col, r0, r1, r2, r3);
    // bounded wait: report instead of trapping
    const long long t0 = clock64();
    int ok = 0;
    while (clock64() - t0 < 200000000ll) {
      if (mbar_try(smem_u32(&bar), 0)) { ok = 1; break; }
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaSetDevice(0);
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  const int iters = 4096;
  printf("# tcgen05.mma cta_group::1 kind::f16 M=128 K=16: cycles per MMA (issue / completion), %d MMAs back to back, 148 CTAs\n", iters);
  const int ns[] = {16, 32, 64, 96, 128, 160, 192, 224, 256};
  for (int shift : {0, 1, 17}) {
    for (int two : {0, 1}) {
      for (int n : ns) {
        if (two && n > 256) continue;
        mma_rate_kernel<<<148, 128, 100 * 1024>>>(n, iters, shift, two, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("a_shift_rows=%2d two_acc=%d N=%3d  issue %.1f  done %.1f cyc/MMA  (%.0f flop/clk/SM) %s\n", shift, two, n,
               (double)h[0] / iters, (double)h[1] / iters, 2.0 * 128 * n * 16 * iters / (double)h[1],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }

  // ---- gather4
  EncodeTiledFn enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn)p;
  }
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 0; }
  const int ROWS = 512, COLS = 256;
  std::vector<__half> h((size_t)ROWS * COLS);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < COLS; ++c) h[(size_t)r * COLS + c] = __float2half((float)(r + (c % 64) / 64.0f));   // value = row + col/64 (exact in fp16 for r < 512? no: coarse) -> use row only + col marker below
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < COLS; ++c) h[(size_t)r * COLS + c] = __float2half((float)((r % 64) * 8 + (c % 64) / 8));   // identifies row%64 and 16-byte chunk
  __half* d_m;
  cudaMalloc(&d_m, h.size() * 2);
  cudaMemcpy(d_m, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  __half* d_o;
  int* d_s;
  cudaMalloc(&d_o, 2048);
  cudaMalloc(&d_s, 4);
  for (int box1 : {1, 4}) {
    for (int swz : {1, 0}) {
      CUtensorMap map;
      cuuint64_t gdim[2] = {(cuuint64_t)COLS, (cuuint64_t)ROWS}, gstr[1] = {(cuuint64_t)COLS * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)box1}, es[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_m, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("# gather4: box {64,%d} swizzle %s: encode rc=%d\n", box1, swz ? "128B" : "none", (int)r);
      if (r != CUDA_SUCCESS) continue;
      for (int dst_off : {0, 512}) {
        cudaMemset(d_s, 0, 4);
        gather4_kernel<<<1, 128, 4096>>>(map, 5, 17, 2, 300, 64, dst_off, d_o, d_s);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        cudaMemcpy(&st, d_s, 4, cudaMemcpyDeviceToHost);
        std::vector<__half> o(1024);
        cudaMemcpy(o.data(), d_o, 2048, cudaMemcpyDeviceToHost);
        printf("rows (5,17,2,300) col 64 -> smem+%d: completed=%d %s\n", dst_off, st, e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) return 0;
        // print per 128-byte smem row: for each 16-byte chunk the (row%64, chunk) code of its first element
        for (int sr = 0; sr < 8; ++sr) {
          printf("  smem row %d:", sr);
          for (int ch = 0; ch < 8; ++ch) {
            const float v = __half2float(o[sr * 64 + ch * 8]);
            if (v < 0) printf("   --  ");
            else printf(" r%02d.c%d", (int)v / 8, (int)v % 8);
          }
          printf("\n");
        }
      }
    }
  }
  return 0;
}
