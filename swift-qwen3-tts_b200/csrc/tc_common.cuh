// Shared device helpers of the tcgen05 kernels: PTX wrappers (mbarrier, TMA, tcgen05), the K-major
// 128B-swizzle smem descriptor, and 16-element row load/store helpers for the fused epilogues.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace q3 {
namespace tc {

// Host: the dynamic-shared-memory opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) is a property of a kernel ON ONE DEVICE.
// One handle per GPU in one process launches the same kernels on every device, so the opt-in is made once per
// (kernel set, device) -- a process-wide once would leave devices 1..N-1 at the 48 KB default and their launches would fail.
// `set` applies the attributes on the CURRENT device and returns the first error.
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done[4] = {0, 0, 0, 0};   // 256 device ordinals
  template <typename F>
  cudaError_t ensure(F&& set) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long& word = done[(dev >> 6) & 3];
    const unsigned long long bit = 1ull << (dev & 63);
    if (word & bit) return cudaSuccess;
    e = set();
    if (e == cudaSuccess) word |= bit;
    return e;
  }
};

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// Same wait, for roles that expect to wait long: the hardware may suspend the thread up to `hint_ns` per try (it is woken by the
// barrier's phase flip), so a waiting warp polls a few times per microsecond instead of burning issue slots the working warps need.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 4000) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity), "r"(hint_ns) : "memory");
  } while (!done);
}
// Wait for roles that expect to wait LONG (epilogue / element-wise warps of a pipeline whose tensor work runs for microseconds).
// ncu on the fused residual unit (profiles/r2_resunit.md): with `try_wait ..., suspendTimeHint` ptxas emits
// TRYWAIT + NANOSLEEP.SYNCS + PHASECHK + BRA and the sleep returns on ANY barrier event of the CTA -- 79 iterations per wait, 32 % of
// all issued warp-instructions were these four.  Here a waiting warp really sleeps (nanosleep.u32 is not woken by barrier traffic)
// and polls with the non-blocking test_wait: a few instructions per microsecond per waiting warp.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  if (mbar_test(bar, parity)) return;
  if (ns == 0) { mbar_wait(bar, parity); return; }
  do {
    asm volatile("nanosleep.u32 %0;" ::"r"(ns) : "memory");
  } while (!mbar_test(bar, parity));
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Column-sliced TMEM access (the mma.m16n8 accumulator layout): 16 lanes x 256 bits, twice along the columns.  With taddr = (lane l0,
// column c), thread t of the warp holds, for j = 0, 1:
//   r[4j+0], r[4j+1] = (lane l0 + t/4,     columns c + 8j + 2(t%4), + 1)
//   r[4j+2], r[4j+3] = (lane l0 + t/4 + 8, same columns)
// A thread sees 4 of every 16 columns, so per-column constants of a 48-column slice are 12 values per thread: they live in registers.
// (With .32x32b a thread holds a whole row, needs every column's constants, and re-reads them from shared memory for every tile.)
__device__ __forceinline__ void tc_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
// The matching store of 16-bit pairs: 16 lanes x 128 bits, twice.  Thread t holds r[2j] = (lane l0 + t/4, 32-bit column c + 4j + t%4),
// r[2j+1] = (lane l0 + t/4 + 8, same column): two adjacent fp32 columns of the load above, packed, land in one 32-bit column.
__device__ __forceinline__ void tc_st_16x128b_x2(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// K-major, 128-byte-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart (SBO), descriptor v1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <typename T16> struct Cvt;
template <> struct Cvt<__half> {
  static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { __half2 h = __hadd2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&h; }
  static __device__ __forceinline__ uint32_t pack(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *(uint32_t*)&h; }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*(__half2*)&u); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { __nv_bfloat162 h = __hadd2(*(__nv_bfloat162*)&a, *(__nv_bfloat162*)&b); return *(uint32_t*)&h; }
  static __device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *(uint32_t*)&h; }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*(__nv_bfloat162*)&u); }
};

// 16 consecutive values of one row -> global, as stream type TS (float or T16)
template <typename T16>
__device__ __forceinline__ void store16(void* base, bool as_f32, long long off, const float (&v)[16]) {
  if (as_f32) {
    float4* p = (float4*)((float*)base + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4* p = (uint4*)((T16*)base + off);
#pragma unroll
    for (int i = 0; i < 2; ++i)
      p[i] = make_uint4(Cvt<T16>::pack(v[8 * i], v[8 * i + 1]), Cvt<T16>::pack(v[8 * i + 2], v[8 * i + 3]),
                        Cvt<T16>::pack(v[8 * i + 4], v[8 * i + 5]), Cvt<T16>::pack(v[8 * i + 6], v[8 * i + 7]));
  }
}
template <typename T16>
__device__ __forceinline__ void load16(const void* base, bool as_f32, long long off, float (&v)[16]) {
  if (as_f32) {
    const float4* p = (const float4*)((const float*)base + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float4 t = p[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
  } else {
    const uint4* p = (const uint4*)((const T16*)base + off);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 t = p[i];
      float2 a = Cvt<T16>::unpack(t.x), b = Cvt<T16>::unpack(t.y), c = Cvt<T16>::unpack(t.z), d = Cvt<T16>::unpack(t.w);
      v[8 * i] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = b.x; v[8 * i + 3] = b.y;
      v[8 * i + 4] = c.x; v[8 * i + 5] = c.y; v[8 * i + 6] = d.x; v[8 * i + 7] = d.y;
    }
  }
}


// ---- cluster / multicast variants ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- TMA store (smem -> global) and sub-CTA named barriers -------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all but the 2 most recent bulk groups of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- CTA-pair (cta_group::2) variants: one MMA spans two SMs, M = 256, each CTA holds half of B -------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's leader CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from TENSOR MEMORY (lane = row, one 32-bit column = two consecutive K elements), B from smem.
__device__ __forceinline__ void tc_mma_f16_2sm_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// Arrive on `bar` of CTA `cta_rank` of the cluster.  Default (.release.cta) semantics on purpose: the only thing ordered
// before this arrive is a completed tcgen05.ld (tcgen05.wait::ld + tcgen05.fence::before_thread_sync); the
// .release.cluster form costs a MEMBAR.ALL.GPU + ERRBAR per arrive (7 % of the epilogue's cycles in the first capture).
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta_rank) : "memory");
}
// One lane of the (converged) warp: returns 1 in the elected lane, 0 elsewhere.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// ---- explicit shared-memory accesses (32-bit shared addresses: guaranteed LDS / STS, no generic-address checks) ---
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds4f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- shared pieces of the fused epilogues ---------------------------------------------------------------------
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// Four per-column constants: from smem (32-bit shared address, LDS.128) or from global memory (read-only path).
template <bool kSmem>
__device__ __forceinline__ float4 ldc4(const float* gptr, uint32_t saddr, int off) {
  if (kSmem) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr + 4u * (uint32_t)off));
    return v;
  }
  return __ldg((const float4*)(gptr + off));
}

// Block epilogue, one 32-row x 32-column chunk of one warp (a lane owns one row):
//   v = acc + bias [+ res16] [-> exact GELU];   y = v (16-bit, kY);   a = v + ib * sin^2(v * ea) (16-bit, kA)
// The four 16-byte pieces of a row go to row_y / row_a + ((chunk_base + c) ^ swz) * 16: a warp-private staging row
// (64-byte rows, SWIZZLE_64B: swz = (lane >> 1) & 3) that leaves through the warp's own TMA store, a row of a 128B-swizzled
// operand tile in smem (swz = row & 7, chunk_base = 0 or 4), or the row in global memory itself (swz = 0).
// bias / ea / ib come from smem (staged constants) or global memory.
template <typename T16, bool kRes, bool kY, bool kA, bool kSmem, bool kGelu = false>
__device__ __forceinline__ void epi_block_chunk(const uint32_t (&r)[32], const float* bias, const float* ea, const float* ib,
                                                uint32_t s_bias, uint32_t s_ea, uint32_t s_ib, const uint4 (&rres)[4],
                                                uint8_t* row_y, uint8_t* row_a, uint32_t chunk_base, uint32_t swz, bool store_ok) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // 8 columns per step
    float v[8];
    const float4 b0 = ldc4<kSmem>(bias, s_bias, 8 * c), b1 = ldc4<kSmem>(bias, s_bias, 8 * c + 4);
    v[0] = __uint_as_float(r[8 * c + 0]) + b0.x; v[1] = __uint_as_float(r[8 * c + 1]) + b0.y;
    v[2] = __uint_as_float(r[8 * c + 2]) + b0.z; v[3] = __uint_as_float(r[8 * c + 3]) + b0.w;
    v[4] = __uint_as_float(r[8 * c + 4]) + b1.x; v[5] = __uint_as_float(r[8 * c + 5]) + b1.y;
    v[6] = __uint_as_float(r[8 * c + 6]) + b1.z; v[7] = __uint_as_float(r[8 * c + 7]) + b1.w;
    if (kRes) {
      const uint4 u = rres[c];
      const float2 r0 = Cvt<T16>::unpack(u.x), r1 = Cvt<T16>::unpack(u.y), r2 = Cvt<T16>::unpack(u.z), r3 = Cvt<T16>::unpack(u.w);
      v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y; v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
    }
    if (kGelu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = gelu_exact(v[e]);
    }
    const uint32_t at = ((chunk_base + (uint32_t)c) ^ swz) << 4;
    if (kY && store_ok)
      *(uint4*)(row_y + at) = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]),
                                         Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
    if (kA) {
      const float4 e0 = ldc4<kSmem>(ea, s_ea, 8 * c), e1 = ldc4<kSmem>(ea, s_ea, 8 * c + 4);
      const float4 i0 = ldc4<kSmem>(ib, s_ib, 8 * c), i1 = ldc4<kSmem>(ib, s_ib, 8 * c + 4);
      const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float sn = __sinf(v[e] * ee[e]);
        v[e] = fmaf(ii[e], sn * sn, v[e]);
      }
      if (store_ok)
        *(uint4*)(row_a + at) = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]),
                                           Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
    }
  }
}

// Two-pass form of the residual epilogue for ONE staging buffer per warp (fused residual unit): pass 1 turns the accumulator
// chunk into v = acc + bias + residual IN PLACE (fp32 bits in r) and writes the packed stream row; pass 2 writes snake(v).
// `row` is this lane's 64-byte row of the warp's SWIZZLE_64B staging buffer (swz = (lane >> 1) & 3).
template <typename T16>
__device__ __forceinline__ void epi_res_pass1(uint32_t (&r)[32], uint32_t s_bias, const uint4 (&rres)[4], uint32_t row, uint32_t swz) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b0 = lds4f(s_bias + 32u * c), b1 = lds4f(s_bias + 32u * c + 16u);
    const uint4 u = rres[c];
    const float2 r0 = Cvt<T16>::unpack(u.x), r1 = Cvt<T16>::unpack(u.y), r2 = Cvt<T16>::unpack(u.z), r3 = Cvt<T16>::unpack(u.w);
    float v[8];
    v[0] = __uint_as_float(r[8 * c + 0]) + b0.x + r0.x; v[1] = __uint_as_float(r[8 * c + 1]) + b0.y + r0.y;
    v[2] = __uint_as_float(r[8 * c + 2]) + b0.z + r1.x; v[3] = __uint_as_float(r[8 * c + 3]) + b0.w + r1.y;
    v[4] = __uint_as_float(r[8 * c + 4]) + b1.x + r2.x; v[5] = __uint_as_float(r[8 * c + 5]) + b1.y + r2.y;
    v[6] = __uint_as_float(r[8 * c + 6]) + b1.z + r3.x; v[7] = __uint_as_float(r[8 * c + 7]) + b1.w + r3.y;
#pragma unroll
    for (int e = 0; e < 8; ++e) r[8 * c + e] = __float_as_uint(v[e]);
    sts128(row + ((((uint32_t)c) ^ swz) << 4), make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]),
                                                          Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7])));
  }
}
// acc + bias -> [snake] -> 16 packed words (the 1x1 conv's operand, kept in tensor memory)
template <typename T16>
__device__ __forceinline__ void epi_snake_pack(const uint32_t (&r)[32], uint32_t s_bias, uint32_t s_ea, uint32_t s_ib, bool snake, uint32_t (&out)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b0 = lds4f(s_bias + 32u * c), b1 = lds4f(s_bias + 32u * c + 16u);
    float v[8];
    v[0] = __uint_as_float(r[8 * c + 0]) + b0.x; v[1] = __uint_as_float(r[8 * c + 1]) + b0.y;
    v[2] = __uint_as_float(r[8 * c + 2]) + b0.z; v[3] = __uint_as_float(r[8 * c + 3]) + b0.w;
    v[4] = __uint_as_float(r[8 * c + 4]) + b1.x; v[5] = __uint_as_float(r[8 * c + 5]) + b1.y;
    v[6] = __uint_as_float(r[8 * c + 6]) + b1.z; v[7] = __uint_as_float(r[8 * c + 7]) + b1.w;
    if (snake) {
      const float4 e0 = lds4f(s_ea + 32u * c), e1 = lds4f(s_ea + 32u * c + 16u);
      const float4 i0 = lds4f(s_ib + 32u * c), i1 = lds4f(s_ib + 32u * c + 16u);
      const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float sn = __sinf(v[e] * ee[e]);
        v[e] = fmaf(ii[e], sn * sn, v[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) out[4 * c + e] = Cvt<T16>::pack(v[2 * e], v[2 * e + 1]);
  }
}
template <typename T16>
__device__ __forceinline__ void epi_res_pass2(uint32_t (&r)[32], uint32_t s_ea, uint32_t s_ib) {   // r <- packed snake(v), 16 words
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 e0 = lds4f(s_ea + 32u * c), e1 = lds4f(s_ea + 32u * c + 16u);
    const float4 i0 = lds4f(s_ib + 32u * c), i1 = lds4f(s_ib + 32u * c + 16u);
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = __uint_as_float(r[8 * c + e]);
      const float sn = __sinf(v[e] * ee[e]);
      v[e] = fmaf(ii[e], sn * sn, v[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) r[4 * c + e] = Cvt<T16>::pack(v[2 * e], v[2 * e + 1]);
  }
}

}  // namespace tc
}  // namespace q3
