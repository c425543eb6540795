// PTX wrappers shared by the tcgen05 convolution kernels: mbarriers, cp.async,
// bulk / tensor TMA copies, tcgen05 MMA + TMEM loads, UMMA descriptors.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace laud {
namespace {

constexpr long long SPIN_CYCLES = 4000000000ll;  // watchdog (~2 s): trap instead of hanging the GPU
constexpr int UMMA_BM = 128;                     // UMMA M: accumulator rows (TMEM lanes) per tile

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok;
}
// try_wait suspends the thread in hardware for a bounded time, so the fast path is a handful of
// instructions; the watchdog clock is only read once a wait has already lasted many probes.
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    if (mbar_try(addr, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  while (!mbar_try(addr, parity)) {
    __nanosleep(32);
    if (clock64() - t0 > SPIN_CYCLES) __trap();
  }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// arrive on the mbarrier once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
// bulk (TMA engine) row copies
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// (barrier init, TMEM allocation, constant tables) while its predecessor in the stream drains; pdl_wait() returns once the
// predecessor has completed and its writes are visible (immediately when the launch carried no such attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(unsigned long long* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b))
               : "memory");
}
// Warp-uniform variants: EVERY lane of a converged warp executes these with identical operands and one elected lane
// issues.  Inside an `if (lane == 0)` region the compiler must assume divergent operands and wraps each tcgen05.mma in
// an ELECT / R2UR / BRA.U.ANY loop (~130 cycles of issue per MMA - as long as the MMA itself); warp-uniform code keeps
// the descriptors in uniform registers.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Four consecutive K=16 steps of one 64-channel chunk in ONE asm block (A advances 32 bytes per step inside the swizzle
// atom, B by `bstep` descriptor units): the tensor-core instruction takes its operands from UNIFORM registers, and ptxas
// moves every operand of every tcgen05.mma there with its own R2UR.BROADCAST (seven per MMA, ~24 issue slots - the MMA warp
// shares its scheduler with two epilogue warps, measured ~110 cycles per MMA whatever N).  Inside one block the accumulator
// address, the instruction descriptor and the descriptor high words are moved once for the four MMAs, one ELECT serves all.
__device__ __forceinline__ void umma_f16_elect_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate_first, uint32_t bstep) {
  asm volatile(
      "{\n\t.reg .pred p, pe, pt;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, st;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "cvt.u64.u32 st, %5;\n\t"
      "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
      "add.u64 b1, %2, st;\n\tadd.u64 b2, b1, st;\n\tadd.u64 b3, b2, st;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(bstep)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(unsigned long long* b) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TMA tensor-tile copies (cp.async.bulk.tensor): `map` is the generic address of a CUtensorMap
// (a __grid_constant__ kernel parameter); coordinates are innermost first.
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, unsigned long long* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, unsigned long long* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA row gather: four rows r0..r3 of a 2-d tensor (box {cols, 1}), columns [col, col + box) each, land as four
// consecutive box-rows at dst (sm_100a; measured: SWIZZLE_128B is applied by shared-memory ADDRESS, so groups of four
// rows written at consecutive 512-byte offsets form a regular K-major swizzled tile).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* map, unsigned long long* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// explicit shared-space accesses (32-bit addresses): keep the epilogue free of generic-address arithmetic
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u1(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_i1(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// one 32-byte global store (STG.256, sm_100): `p` 32-byte aligned
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Shared-memory matrix descriptors (SWIZZLE_128B).
//  K-major : rows (m or n) of 128 B = 64 k; 8-row groups `sbo` bytes apart.
//  MN-major: rows (k) of 128 B = 64 n; 8-k groups `sbo` bytes apart, 64-n blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_f16(int n, int b_mn_major) {
  return (1u << 4)                      // D = fp32
         | (0u << 7) | (0u << 10)       // A, B = fp16
         | (0u << 15)                   // A K-major
         | ((uint32_t)b_mn_major << 16) // B K-major (0) / MN-major (1)
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UMMA_BM >> 4) << 24);
}
// byte offset of (row r, 16-byte chunk c) inside a swizzled K-major tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}


}  // namespace
}  // namespace laud
