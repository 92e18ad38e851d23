// Fused residual unit of the 96-channel decoder block (ST.swift:408-437, block 3 of the 12 Hz decoder):
//
//     X' = X + conv1x1( snake2( conv7_dil(snake1(X)) + b7 ) ) + b1          [+ out = snake3(X') for the last unit]
//
// as ONE persistent sm_100a kernel.  The unfused chain (conv7 GEMM, then conv1 GEMM) moves 6 x 192 B per output row
// through HBM (A in, C out, C in, X in, X out, A out) and sits at the HBM roofline; here X is read once and X' written
// once (2.2 x 192 B per row with the halo), and both weight tensors stay resident in shared memory:
//
//   * the two CTAs of a cluster form a cta_group::2 pair (M = 256): each CTA walks its OWN contiguous strip of 128-row
//     tiles in time order and holds half (48 output channels) of W7 [7,96,96] and W1 [96,96] in smem (72 KB);
//   * activations live in a 3-slot smem ring of 128-row tiles per 32-channel block (64-byte rows, SWIZZLE_64B, K-major).
//     A tile's causal halo (6*dil rows) IS the tail of the previous slot, so it is neither re-read nor re-activated; slot 0's
//     halo is a mirror of slot 2's tail.  Only the first tile of a strip loads its halo from HBM (out-of-range rows of an
//     utterance's first tile are TMA zero fill = the causal padding, and snake(0) = 0);
//   * pass 1 (12 warps): snake1 in place on the freshly landed tile;  conv7 = 42 tcgen05.mma (7 taps x 3 blocks x 2 k-steps,
//     tap j = the ring viewed from row j*dil) into TMEM;  epilogue 1: + b7, snake2, 16-bit, written straight into the
//     swizzled operand tile of conv1 (6 MMAs);  epilogue 2: + b1 + X (residual rows re-read from L2), TMA store;
//   * software pipeline across tiles: while the tensor pipe runs conv7(i), the CUDA cores run pass1(i+1), epilogue1(i-1)
//     and epilogue2(i-2); MMA order is c7(0) c7(1) c1(0) c7(2) c1(1) ...
//   * cross-CTA hand-offs (operand tile ready, accumulator drained) are counted per CTA in smem; the last arriving warp
//     forwards ONE arrive to the leader's mbarrier (a cluster-scope release per warp would cost a MEMBAR.ALL.GPU each).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace q3 {

namespace {
using namespace tc;

constexpr int R_C = 96, R_SUB = 3, R_BM = 128, R_SLOTS = 3, R_THREADS = 512, R_EW = 12;
constexpr uint32_t R_WBLK = 48 * 64;                 // one (tap, 32-channel block) of this CTA's weight half: 48 rows x 64 B
constexpr uint32_t R_W7_BYTES = 7 * R_SUB * R_WBLK;  // 64512
constexpr uint32_t R_W1_BYTES = R_SUB * R_WBLK;      // 9216
constexpr uint32_t R_CSUB = R_BM * 64;               // one 32-channel block of the conv1 operand tile
constexpr uint32_t R_STAGE = R_EW * 2048;            // per-warp output staging (32 rows x 64 B)
constexpr uint32_t R_CST = 6 * R_C * 4;              // b7, ea2, ib2, b1, ea3, ib3
constexpr uint32_t R_TMEM_COLS = 512;

struct Res96Params {
  int B, Tmax, rows_per_frame;
  const int* len_frames;
  int dil, halo, hb;              // halo = 6*dil rows; hb = halo rounded up to 8 rows (TMA box and smem alignment)
  int tiles_per_cta;
  uint32_t tsub_bytes;            // one 32-channel block of the ring: (hb + 3*128) rows x 64 B, rounded to 1024
  uint32_t idesc7;                // M = 256, N = 96
  const float *b7, *ea1, *ib1, *ea2, *ib2, *b1, *ea3, *ib3;
  int out_snake;                  // write snake3(X') instead of X' (last unit of the block)
  const void* x_in; long long x_bstride;   // residual rows (elements)
};

// Position of a CTA in its strip of valid tiles (warp-uniform).
struct Walker {
  int b, t0, valid;
  bool first;
  __device__ __forceinline__ void load_valid(const Res96Params& p) {
    while (b < p.B) {
      valid = __ldg(p.len_frames + b) * p.rows_per_frame;
      if (valid > 0) return;
      ++b;
    }
    valid = 0;
  }
  __device__ __forceinline__ void init(const Res96Params& p, long long g0) {   // g0 = index of the first tile in the list of valid tiles
    b = 0; t0 = 0; first = true;
    long long rem = g0;
    for (; b < p.B; ++b) {
      const int v = __ldg(p.len_frames + b) * p.rows_per_frame;
      const long long nt = (v + R_BM - 1) / R_BM;
      if (rem < nt) break;
      rem -= nt;
    }
    t0 = (int)rem * R_BM;
    load_valid(p);
  }
  __device__ __forceinline__ bool live(const Res96Params& p) const { return b < p.B; }
  __device__ __forceinline__ void next(const Res96Params& p) {
    if (b >= p.B) return;
    t0 += R_BM;
    first = false;
    if (t0 >= valid) { ++b; t0 = 0; first = true; load_valid(p); }
  }
};

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}

// Called by lane 0 of each of the 12 elementwise warps (after __syncwarp): the last warp of this CTA forwards one
// arrive to the LEADER's barrier (count 2: one per CTA of the pair).
__device__ __forceinline__ void arrive_pair(uint32_t* cnt, uint64_t* leader_bar, uint32_t rank) {
  __threadfence_block();
  const uint32_t old = atomicAdd(cnt, 1u);
  if (old == R_EW - 1) {
    atomicExch(cnt, 0u);
    __threadfence_block();
    if (rank == 0) {
      mbar_arrive(leader_bar);
    } else {
      asm volatile(
          "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\t"
          "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(leader_bar)) : "memory");
    }
  }
}

template <typename T16>
__global__ void __launch_bounds__(R_THREADS, 1)
resunit96_kernel(const __grid_constant__ CUtensorMap map_main, const __grid_constant__ CUtensorMap map_halo,
                 const __grid_constant__ CUtensorMap map_w7, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_out, Res96Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* t_ring = smem;                                    // [3 blocks][hb + 384 rows][64 B]
  uint8_t* c_tile = t_ring + (size_t)R_SUB * p.tsub_bytes;   // [3 blocks][128 rows][64 B]
  uint8_t* w7s = c_tile + R_SUB * R_CSUB;                    // [7 taps][3 blocks][48 rows][64 B]
  uint8_t* w1s = w7s + R_W7_BYTES;                           // [3 blocks][48 rows][64 B]
  uint8_t* staging = w1s + R_W1_BYTES;
  float* cst = (float*)(staging + R_STAGE);
  uint64_t* bars = (uint64_t*)((uint8_t*)cst + R_CST);
  uint64_t* w_full = bars;              // 1
  uint64_t* t_full = bars + 1;          // [3] local: X tile landed
  uint64_t* a_ready = bars + 4;         // [3] leader: pass 1 done in both CTAs
  uint64_t* c7_done = bars + 7;         // [3] both: conv7 of the tile in this slot has completed
  uint64_t* acc1_full = bars + 10;      // [2] both
  uint64_t* c_ready = bars + 12;        // [2] leader: epilogue 1 done in both CTAs (operand tile written, acc1 drained)
  uint64_t* acc2_full = bars + 14;      // [2] both: conv1 has completed (operand tile free again)
  uint64_t* acc2_free = bars + 16;      // [2] leader: epilogue 2 has drained acc2 in both CTAs
  uint32_t* cnt = (uint32_t*)(bars + 18);   // cnt_a[3], cnt_c[2], cnt_f[2]
  uint32_t* tmem_ptr = cnt + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint16_t mc_mask = 3;
  const int q = p.tiles_per_cta;
  const int hb = p.hb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_main) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_halo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w7) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&t_full[i], 1); mbar_init(&a_ready[i], 2); mbar_init(&c7_done[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&c_ready[i], 2); mbar_init(&acc2_full[i], 1); mbar_init(&acc2_free[i], 2); }
    for (int i = 0; i < 8; ++i) cnt[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(R_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= 4) {
    for (int i = (int)threadIdx.x - 128; i < R_C; i += R_EW * 32) {
      cst[i] = __ldg(p.b7 + i); cst[R_C + i] = __ldg(p.ea2 + i); cst[2 * R_C + i] = __ldg(p.ib2 + i);
      cst[3 * R_C + i] = __ldg(p.b1 + i);
      cst[4 * R_C + i] = p.out_snake ? __ldg(p.ea3 + i) : 0.f; cst[5 * R_C + i] = p.out_snake ? __ldg(p.ib3 + i) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long g0 = (long long)blockIdx.x * q;

  if (warp == 0) {
    // ================= producer: weights once, then one X tile per step =================
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(w_full, 2u * (R_W7_BYTES + R_W1_BYTES));
      for (int blk = 0; blk < 7 * R_SUB; ++blk)
        tma_load_2d_2sm(w7s + (size_t)blk * R_WBLK, &map_w7, w_full, (blk % R_SUB) * 32, (blk / R_SUB) * R_C + (int)rank * 48);
      for (int c = 0; c < R_SUB; ++c) tma_load_2d_2sm(w1s + (size_t)c * R_WBLK, &map_w1, w_full, c * 32, (int)rank * 48);
    }
    __syncwarp();
    Walker w;
    w.init(p, g0);
    for (int j = 0; j < q; ++j) {
      const int slot = j % R_SLOTS;
      if (j >= 2) mbar_wait(&c7_done[(j - 2) % R_SLOTS], (uint32_t)(((j - 2) / R_SLOTS) & 1));   // slot's old tile and its use as a halo are over
      if (w.first && j >= 1) mbar_wait(&c7_done[(j - 1) % R_SLOTS], (uint32_t)(((j - 1) / R_SLOTS) & 1));   // the halo lands in the previous slot's tail
      if (elect_one()) {
        if (w.live(p)) {
          mbar_expect_tx(&t_full[slot], (uint32_t)R_SUB * (uint32_t)(R_BM + (w.first ? hb : 0)) * 64u);
          for (int c = 0; c < R_SUB; ++c)
            tma_load_3d(t_ring + (size_t)c * p.tsub_bytes + (size_t)(hb + slot * R_BM) * 64, &map_main, &t_full[slot], c * 32, w.t0, w.b);
          if (w.first)
            for (int c = 0; c < R_SUB; ++c)
              tma_load_3d(t_ring + (size_t)c * p.tsub_bytes + (size_t)(slot * R_BM) * 64, &map_halo, &t_full[slot], c * 32, w.t0 - hb, w.b);
        } else {
          mbar_arrive(&t_full[slot]);
        }
      }
      __syncwarp();
      w.next(p);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint64_t desc_fixed = ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29)) << 32) | (1u << 16);   // SBO 512 B, v1, SWIZZLE_64B
      const uint32_t t_u32 = smem_u32(t_ring), c_u32 = smem_u32(c_tile), w7_u32 = smem_u32(w7s), w1_u32 = smem_u32(w1s);
      const uint32_t idesc = p.idesc7;
      const uint64_t dTap = (uint64_t)(p.dil * 4);             // dil rows x 64 B, in 16-byte units
      const uint32_t tsub16 = p.tsub_bytes >> 4;
      mbar_wait(w_full, 0);
      for (int i = 0; i <= q; ++i) {
        if (i < q) {
          const int slot = i % R_SLOTS;
          mbar_wait_cluster(&a_ready[slot], (uint32_t)((i / R_SLOTS) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((i & 1) * 128);
            uint64_t ad = desc_fixed | (uint64_t)((t_u32 + (uint32_t)(hb + slot * R_BM - p.halo) * 64u) >> 4);
            uint64_t wd = desc_fixed | (uint64_t)(w7_u32 >> 4);
#pragma unroll 1
            for (int tap = 0; tap < 7; ++tap) {
#pragma unroll
              for (int c = 0; c < R_SUB; ++c) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const uint32_t accumulate = (c | k) == 0 ? (uint32_t)tap : 1u;
                  tc_mma_f16_2sm(d_tmem, ad + (uint64_t)(c * tsub16 + 2 * k), wd + (uint64_t)(c * (R_WBLK >> 4) + 2 * k), idesc, accumulate);
                }
              }
              ad += dTap;
              wd += (uint64_t)(R_SUB * (R_WBLK >> 4));
            }
            tc_commit_2sm(&c7_done[slot], mc_mask);
            tc_commit_2sm(&acc1_full[i & 1], mc_mask);
          }
          __syncwarp();
        }
        if (i >= 1) {
          const int t = i - 1;
          mbar_wait_cluster(&c_ready[t & 1], (uint32_t)((t >> 1) & 1));
          if (t >= 2) mbar_wait_cluster(&acc2_free[t & 1], (uint32_t)(((t - 2) >> 1) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + 256u + (uint32_t)((t & 1) * 128);
            const uint64_t ad = desc_fixed | (uint64_t)(c_u32 >> 4), wd = desc_fixed | (uint64_t)(w1_u32 >> 4);
#pragma unroll
            for (int c = 0; c < R_SUB; ++c) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                tc_mma_f16_2sm(d_tmem, ad + (uint64_t)(c * (R_CSUB >> 4) + 2 * k), wd + (uint64_t)(c * (R_WBLK >> 4) + 2 * k), idesc, (uint32_t)(c | k));
            }
            tc_commit_2sm(&acc2_full[t & 1], mc_mask);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================= elementwise warps: pass 1, epilogue 1, epilogue 2 =================
    const int ew = warp - 4, quarter = warp & 3, grp = ew >> 2;     // epilogues: TMEM lane quarter, 32-channel block
    const int pc = ew % R_SUB, pg = ew / R_SUB;                      // pass 1: 32-channel block, row-group phase (0..3)
    const int kch = lane & 3;                                        // pass 1: this thread's 16-byte chunk = channels pc*32 + 8*kch ..
    float ea1[8], ib1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ea1[e] = __ldg(p.ea1 + pc * 32 + kch * 8 + e); ib1[e] = __ldg(p.ib1 + pc * 32 + kch * 8 + e); }
    uint8_t* const my_sub = t_ring + (size_t)pc * p.tsub_bytes;
    const uint32_t cst_u32 = smem_u32(cst);
    uint8_t* const my_stage = staging + (size_t)ew * 2048;
    uint8_t* const my_ctile = c_tile + (size_t)grp * R_CSUB + (size_t)quarter * 32 * 64;
    const int slot_rows = p.Tmax * p.rows_per_frame;
    const int ngroups = hb / 8 + R_BM / 8;
    Walker w;
    w.init(p, g0);
    int hb1 = 0, ht1 = 0, hb2 = 0, ht2 = 0;          // coordinates of tiles j-1 and j-2
    bool hl1 = false, hl2 = false;
    for (int j = 0; j < q + 2; ++j) {
      // ---- pass 1 of tile j: snake1 in place ----
      if (j < q) {
        const int slot = j % R_SLOTS;
        mbar_wait(&t_full[slot], (uint32_t)((j / R_SLOTS) & 1));
        if (w.live(p)) {
          for (int G = (w.first ? 0 : hb / 8) + pg; G < ngroups; G += 4) {
            const int row = slot * R_BM + G * 8 + (lane >> 2);        // row inside the ring block (row 0 = mirror start)
            uint8_t* ptr = my_sub + (size_t)row * 64 + (size_t)((kch ^ ((row >> 1) & 3)) << 4);
            const uint4 u = *(const uint4*)ptr;
            const uint32_t in[4] = {u.x, u.y, u.z, u.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = Cvt<T16>::unpack(in[e]);
              const float s0 = __sinf(f.x * ea1[2 * e]), s1 = __sinf(f.y * ea1[2 * e + 1]);
              o[e] = Cvt<T16>::pack(fmaf(ib1[2 * e], s0 * s0, f.x), fmaf(ib1[2 * e + 1], s1 * s1, f.y));
            }
            const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
            *(uint4*)ptr = v;
            if (slot == R_SLOTS - 1 && row >= R_SLOTS * R_BM) *(uint4*)(ptr - (size_t)R_SLOTS * R_BM * 64) = v;   // mirror = halo of slot 0
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) arrive_pair(&cnt[slot], &a_ready[slot], rank);
      }
      // ---- epilogue 1 of tile j-1: acc1 + b7 -> snake2 -> conv1 operand tile ----
      if (j >= 1 && j <= q) {
        const int t = j - 1;
        mbar_wait(&acc1_full[t & 1], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((t & 1) * 128 + grp * 32), r);
        tc_wait_ld();
        if (t >= 1) mbar_wait(&acc2_full[(t - 1) & 1], (uint32_t)(((t - 1) >> 1) & 1));   // conv1(t-1) has finished reading the operand tile
        if (hl1) {
          const uint32_t sb = cst_u32 + 4u * (uint32_t)(grp * 32), se = sb + 4u * R_C, si = se + 4u * R_C;
          const uint4 none[4] = {};
          epi_block_chunk<T16, false, false, true, true>(r, nullptr, nullptr, nullptr, sb, se, si, none, nullptr, my_ctile, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_pair(&cnt[3 + (t & 1)], &c_ready[t & 1], rank);
      }
      // ---- epilogue 2 of tile j-2: acc2 + b1 + X -> X' (or snake3(X')) -> TMA store ----
      if (j >= 2) {
        const int t = j - 2;
        const int row = min(ht2 + quarter * 32 + lane, slot_rows - 1);
        uint4 rres[4] = {};
        if (hl2) {
          const T16* res_row = (const T16*)p.x_in + (long long)hb2 * p.x_bstride + (long long)row * R_C + grp * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) rres[c] = __ldg((const uint4*)(res_row + 8 * c));
        }
        mbar_wait(&acc2_full[t & 1], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + (t & 1) * 128 + grp * 32), r);
        tc_wait_ld();
        if (hl2) {
          if (lane == 0) tma_store_wait_read0();
          __syncwarp();
          const uint32_t sb = cst_u32 + 4u * (uint32_t)(3 * R_C + grp * 32), se = sb + 4u * R_C, si = se + 4u * R_C;
          if (p.out_snake) epi_block_chunk<T16, true, false, true, true>(r, nullptr, nullptr, nullptr, sb, se, si, rres, nullptr, my_stage, lane);
          else epi_block_chunk<T16, true, true, false, true>(r, nullptr, nullptr, nullptr, sb, se, si, rres, my_stage, nullptr, lane);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_out, my_stage, grp * 32, ht2 + quarter * 32, hb2);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_pair(&cnt[5 + (t & 1)], &acc2_free[t & 1], rank);
      }
      hb2 = hb1; ht2 = ht1; hl2 = hl1;
      hb1 = w.b; ht1 = w.t0; hl1 = (j < q) && w.live(p);
      if (j < q) w.next(p);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(R_TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_r() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}
}  // namespace

bool resunit96_supported(const ResUnitParams& p, int op_dtype) {
  static const int on = []() { const char* e = getenv("Q3TTS_FUSED_RES"); return e ? atoi(e) : 1; }();
  if (!on) return false;
  if (op_dtype != DT_F16 && op_dtype != DT_BF16) return false;
  if (p.C != R_C || p.dil < 1 || 6 * p.dil + R_BM > 256) return false;
  return encode_fn_r() != nullptr;
}

cudaError_t launch_resunit96(const ResUnitParams& p, const BatchGeom& g, int op_dtype, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn_r();
  if (!enc) return cudaErrorNotSupported;
  const CUtensorMapDataType dt = op_dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int slot_rows = g.Tmax * p.rows_per_frame;
  Res96Params q{};
  q.B = g.B; q.Tmax = g.Tmax; q.rows_per_frame = p.rows_per_frame; q.len_frames = g.len_frames;
  q.dil = p.dil; q.halo = 6 * p.dil; q.hb = (q.halo + 7) & ~7;
  q.tsub_bytes = ((uint32_t)(q.hb + R_SLOTS * R_BM) * 64u + 1023u) & ~1023u;
  const uint32_t fmt = op_dtype == DT_F16 ? 0u : 1u;
  q.idesc7 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(R_C >> 3) << 17) | ((uint32_t)((2 * R_BM) >> 4) << 24);
  q.b7 = p.b7; q.ea1 = p.ea1; q.ib1 = p.ib1; q.ea2 = p.ea2; q.ib2 = p.ib2; q.b1 = p.b1; q.ea3 = p.ea3; q.ib3 = p.ib3;
  q.out_snake = p.ea3 != nullptr;
  q.x_in = p.x_in; q.x_bstride = (long long)slot_rows * R_C;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles_bound = (g.valid_frames * p.rows_per_frame + R_BM - 1) / R_BM + g.B;   // sum of per-utterance ceilings <= this
  int grid = (int)std::min<long long>(sms / 2 * 2, (tiles_bound + 1) / 2 * 2);
  grid = std::max(grid, 2);
  q.tiles_per_cta = (int)((tiles_bound + grid - 1) / grid);
  CUtensorMap map_main, map_halo, map_w7, map_w1, map_out;
  auto act_map = [&](CUtensorMap* m, const void* base, cuuint32_t rows) -> bool {
    cuuint64_t dims[3] = {(cuuint64_t)R_C, (cuuint64_t)slot_rows, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)R_C * 2, (cuuint64_t)slot_rows * R_C * 2};
    cuuint32_t box[3] = {32, rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto w_map = [&](CUtensorMap* m, const void* base, int rows) -> bool {
    cuuint64_t dims[2] = {(cuuint64_t)R_C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)R_C * 2};
    cuuint32_t box[2] = {32, 48};
    cuuint32_t es[2] = {1, 1};
    return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!act_map(&map_main, p.x_in, R_BM) || !act_map(&map_halo, p.x_in, (cuuint32_t)q.hb) || !act_map(&map_out, p.out, 32) ||
      !w_map(&map_w7, p.w7, 7 * R_C) || !w_map(&map_w1, p.w1, R_C))
    return cudaErrorInvalidValue;
  const size_t smem = (size_t)R_SUB * q.tsub_bytes + R_SUB * R_CSUB + R_W7_BYTES + R_W1_BYTES + R_STAGE + R_CST + 512 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(R_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static std::once_flag once;
  std::call_once(once, []() {
    cudaFuncSetAttribute(resunit96_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(resunit96_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (op_dtype == DT_F16) return cudaLaunchKernelEx(&cfg, resunit96_kernel<__half>, map_main, map_halo, map_w7, map_w1, map_out, q);
  return cudaLaunchKernelEx(&cfg, resunit96_kernel<__nv_bfloat16>, map_main, map_halo, map_w7, map_w1, map_out, q);
}

}  // namespace q3
