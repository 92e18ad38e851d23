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
//   * activations live in a 3- or 4-slot smem ring per 32-channel block (64-byte rows, SWIZZLE_64B, K-major); a slot is the
//     tile's causal halo (6*dil rows) followed by its 128 rows.  The halo is a copy of the previous tile's last rows made by
//     pass 1, so it is neither re-read nor re-activated.  Only the first tile of a strip loads its halo from HBM (out-of-range
//     rows of an utterance's first tile are TMA zero fill = the causal padding, and snake(0) = 0).  A slot is free again as
//     soon as conv7 of its tile has completed;
//   * pass 1 (6 warps): snake1 in place on the freshly landed tile;  conv7 = 42 tcgen05.mma (7 taps x 3 blocks x 2 k-steps,
//     tap j = the ring viewed from row j*dil) into TMEM;  epilogue 1 (8 warps): + b7, snake2, packed to 16 bits and written
//     with tcgen05.st over the accumulator columns it has just read -- conv1 (6 MMAs) takes its A operand from TENSOR MEMORY,
//     so there is no operand tile in smem at all;  epilogue 2 (8 warps): + b1 + X -> X'.  The residual rows come in by TMA
//     (L2 hits) into a per-warp staging buffer, X' is written over them and leaves by TMA: direct 16-byte global accesses
//     with a 192-byte row pitch cost 32 LSU wavefronts per instruction and slowed every role by 10-20 %;
//   * the roles are decoupled (producer, MMA issuer, pass-1, epilogue-1 and epilogue-2 warps) and meet only at mbarriers, so
//     pass 1 of tile i+1 and the epilogues of tiles i-1 / i-2 fill the issue slots while conv7(i) runs; MMA order is
//     c7(0) c7(1) c1(0) c7(2) c1(1) ...
//   * cross-CTA hand-offs (operand tile ready, accumulator drained): every role warp arrives once on the LEADER's mbarrier
//     with default (CTA-scope) semantics; a cluster-scope release per warp would cost a MEMBAR.ALL.GPU each.
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

// clock64 stamps of CTA 0 (pipeline debugging): compiled in only in the kDbg instantiation -- the predicates alone cost every role a
// few instructions per tile, and the kernel is bound by instruction issue.
#define R_STAMP(ev, tile) do { if (kDbg && p.dbg && blockIdx.x == 0 && (tile) < 32 && lane == 0) p.dbg[(ev) * 32 + (tile)] = clock64(); } while (0)

constexpr int R_C = 96, R_SUB = 3, R_BM = 128, R_SLOTS = 4, R_CTRL = 2, R_PW = 9, R_E1W = 8, R_E2W = 8;   // control (producer, MMA), pass-1, epilogue-1, epilogue-2 warps
constexpr int R_THREADS = (R_CTRL + R_PW + R_E1W + R_E2W) * 32;
constexpr uint32_t R_WBLK = 48 * 64;                 // one (tap, 32-channel block) of this CTA's weight half: 48 rows x 64 B
constexpr uint32_t R_W7_BYTES = 7 * R_SUB * R_WBLK;  // 64512
constexpr uint32_t R_W1_BYTES = R_SUB * R_WBLK;      // 9216
constexpr uint32_t R_CST = 3 * R_C * 4;              // b1, ea3, ib3 (epilogue 2; epilogue 1 keeps b7, ea2, ib2 in registers)
constexpr uint32_t R_TMEM_COLS = 512;
constexpr uint32_t R_STG_TILE = 32 * 48 * 2;         // epilogue 2: one warp's 32 rows x 48 channels (residual in, X' out)
constexpr uint32_t R_STG_WARP = 2 * R_STG_TILE;      // two buffers per warp, alternating by tile
constexpr uint32_t R_STG_BYTES = R_E2W * R_STG_WARP;
constexpr uint32_t R_WOUT_BYTES = R_SUB * 8 * 64;     // tail mode: this CTA's 8 rows of the [16][96] hi / lo outConv tile

struct Res96Params {
  int B, Tmax, rows_per_frame;
  const int* len_frames;
  int dil, halo, hb;              // halo = 6*dil rows; hb = halo rounded up to 8 rows (TMA box and smem alignment)
  int tiles_per_cta;
  int nslots;                     // ring depth (3 or 4 tiles)
  uint32_t tsub_bytes;            // one 32-channel block of the ring: 4 slots x (hb + 128) rows x 64 B, rounded to 1024
  void* out;                      // [B, slot_rows, 96] 16-bit
  uint32_t idesc7;                // M = 256, N = 96
  const float *b7, *ea1, *ib1, *ea2, *ib2, *b1, *ea3, *ib3;
  int out_snake;                  // write snake3(X') instead of X' (last unit of the block)
  const void* x_in; long long x_bstride;   // residual rows (elements)
  long long* dbg;                 // optional [16 events][32 tiles] clock64 stamps of CTA 0 (pipeline debugging)
  uint32_t sleep_ns;              // back-off of the waiting role warps (Q3TTS_RES_SLEEP; 0 = poll)
  int bridge;                     // 1: the epilogue warps block on named barriers released by warp 0 (see the producer role)
  // Fused tail (last unit of the last block): the unit's output a = snake3(X') is NOT written.  Epilogue 2 packs it back into tensor
  // memory and a third MMA multiplies it with outConv's 7 taps (ST.swift:674-678), split into hi + lo 16-bit halves:
  // P[row][0..6] = sum_c w_hi[j][c] a[row][c], P[row][8..14] = the same with w_lo.  Only P (64 B per row instead of 192 B of
  // activations) goes to HBM; the tail kernel adds the 7 shifted rows.  0: off.
  int tail;
  int stg_bufs;                   // staging buffers per epilogue-2 warp (tail mode needs one: nothing is stored from it)
  uint32_t idesc3;                // M = 256, N = 16
  float* p_out; long long p_bstride; int slot_rows;   // P [B, slot_rows, 16] fp32
};

// Named barriers (id 0 is __syncthreads): the 8 warps of an epilogue role block in bar.sync -- a blocked warp issues nothing -- and
// warp 0, which polls the accumulator mbarriers anyway, releases them with bar.arrive.  Two ids per role: accumulators are
// double-buffered, so arrivals for tile t + 2 cannot start before every warp has left the barrier of tile t.
constexpr int R_BAR_E1 = 1, R_BAR_E2 = 3, R_BAR_E1_COUNT = (R_E1W + 1) * 32, R_BAR_E2_COUNT = (R_E2W + 1) * 32;
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

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

// Called by lane 0 of each warp of a role (after __syncwarp): one arrive on the LEADER's barrier (count = the role's warps in
// both CTAs).  Default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id): the operand tiles were published
// to the async proxy by their writers (fence.proxy.async).  Counting arrivals in smem first (atomics + fences) cost each warp
// several hundred cycles per tile.
__device__ __forceinline__ void arrive_leader(uint64_t* leader_bar, uint32_t rank) {
  if (rank == 0) mbar_arrive(leader_bar); else mbar_arrive_remote(leader_bar, 0);
}

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

template <typename T16, bool kDbg, bool kTail>
__global__ void __launch_bounds__(R_THREADS, 1)
resunit96_kernel(const __grid_constant__ CUtensorMap map_main, const __grid_constant__ CUtensorMap map_halo,
                 const __grid_constant__ CUtensorMap map_w7, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                 const __grid_constant__ CUtensorMap map_wout, Res96Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* t_ring = smem;                                    // [3 blocks][4 slots][hb halo rows + 128 rows][64 B]
  uint8_t* w7s = t_ring + (size_t)R_SUB * p.tsub_bytes;      // [7 taps][3 blocks][48 rows][64 B]
  uint8_t* w1s = w7s + R_W7_BYTES;                           // [3 blocks][48 rows][64 B]
  uint8_t* wout_s = w1s + R_W1_BYTES;                        // tail mode: [3 blocks][8 rows (hi taps in CTA 0, lo taps in CTA 1)][64 B]
  uint8_t* stg = wout_s + (kTail ? R_WOUT_BYTES : 0u);      // [8 epilogue-2 warps][stg_bufs buffers][32 rows][96 B]
  float* cst = (float*)(stg + (size_t)R_E2W * R_STG_TILE * (kTail ? 1u : 2u));   // b1, ea3, ib3
  uint64_t* bars = (uint64_t*)((uint8_t*)cst + R_CST);
  uint64_t* w_full = bars;              // 1
  uint64_t* t_full = bars + 1;          // [4] local: X tile landed
  uint64_t* a_ready = bars + 5;         // [4] leader: pass 1 done in both CTAs
  uint64_t* c7_done = bars + 9;         // [4] both: conv7 of the tile in this slot has completed (the slot may be refilled)
  uint64_t* acc1_full = bars + 17;      // [2] both
  uint64_t* c_ready = bars + 19;        // [2] leader: epilogue 1 done in both CTAs (conv1 operand written, acc1 drained)
  uint64_t* acc2_full = bars + 21;      // [2] both: conv1 has completed
  uint64_t* acc2_free = bars + 23;      // [2] leader: epilogue 2 has drained acc2 in both CTAs
  uint64_t* xbar = bars + 25;           // [8][2] local: an epilogue-2 warp's residual rows have landed (per staging buffer)
  uint64_t* a3_ready = xbar + 2 * R_E2W;   // [2] leader: tail mode, epilogue 2 has packed a = snake3(X') into tensor memory (both CTAs)
  uint64_t* acc3_full = a3_ready + 2;      // [2] both: the outConv partial products of the tile are in tensor memory
  uint32_t* tmem_ptr = (uint32_t*)(acc3_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint16_t mc_mask = 3;
  const int q = p.tiles_per_cta;
  const int hb = p.hb, rs = p.hb + R_BM;          // rows per ring slot: its own halo rows, then the tile
  const int ns = p.nslots;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_main) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_halo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w7) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    if (kTail) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wout) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < R_SLOTS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&a_ready[i], 2 * R_PW); mbar_init(&c7_done[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&c_ready[i], 2 * R_E1W); mbar_init(&acc2_full[i], 1); mbar_init(&acc2_free[i], 2 * R_E2W); }
    for (int i = 0; i < 2 * R_E2W; ++i) mbar_init(&xbar[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&a3_ready[i], 2 * R_E2W); mbar_init(&acc3_full[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(R_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= R_CTRL) {
    for (int i = (int)threadIdx.x - R_CTRL * 32; i < R_C; i += (R_PW + R_E1W + R_E2W) * 32) {
      cst[i] = __ldg(p.b1 + i);
      cst[R_C + i] = p.out_snake ? __ldg(p.ea3 + i) : 0.f; cst[2 * R_C + i] = p.out_snake ? __ldg(p.ib3 + i) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();   // programmatic dependent launch (kernels.cuh): the prologue above touched no activations
  if (warp != 0) pdl_wait();   // (the producer requests the weights first)
  const long long g0 = (long long)blockIdx.x * q;

  if (warp == 0) {
    // ================= producer: weights once, then one X tile per step =================
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(w_full, 2u * (R_W7_BYTES + R_W1_BYTES + (kTail ? R_WOUT_BYTES : 0u)));
      for (int blk = 0; blk < 7 * R_SUB; ++blk)
        tma_load_2d_2sm(w7s + (size_t)blk * R_WBLK, &map_w7, w_full, (blk % R_SUB) * 32, (blk / R_SUB) * R_C + (int)rank * 48);
      for (int c = 0; c < R_SUB; ++c) tma_load_2d_2sm(w1s + (size_t)c * R_WBLK, &map_w1, w_full, c * 32, (int)rank * 48);
      if (kTail)
        for (int c = 0; c < R_SUB; ++c) tma_load_2d_2sm(wout_s + (size_t)c * 512, &map_wout, w_full, c * 32, (int)rank * 8);
    }
    __syncwarp();
    pdl_wait();
    Walker w;
    w.init(p, g0);
    auto issue_tile = [&](int j) {          // TMA of tile j into slot j % ns (the walker stands at tile j)
      const int slot = j % ns;
      R_STAMP(0, j);
      if (elect_one()) {
        if (w.live(p)) {
          mbar_expect_tx(&t_full[slot], (uint32_t)R_SUB * (uint32_t)(R_BM + (w.first ? hb : 0)) * 64u);
          for (int c = 0; c < R_SUB; ++c)
            tma_load_3d(t_ring + (size_t)c * p.tsub_bytes + (size_t)(slot * rs + hb) * 64, &map_main, &t_full[slot], c * 32, w.t0, w.b);
          if (w.first)
            for (int c = 0; c < R_SUB; ++c)
              tma_load_3d(t_ring + (size_t)c * p.tsub_bytes + (size_t)(slot * rs) * 64, &map_halo, &t_full[slot], c * 32, w.t0 - hb, w.b);
        } else {
          mbar_arrive(&t_full[slot]);
        }
      }
      __syncwarp();
      w.next(p);
    };
    if (p.bridge) {
      // This warp is also the BRIDGE between the tensor pipe's completion mbarriers and the epilogue warps.  Sixteen epilogue warps
      // polling acc1_full / acc2_full took a third of all issued instructions (ncu, profiles/r2_resunit.md) -- away from the pass-1
      // warps, which are the critical role -- and neither try_wait's suspend hint nor nanosleep really parks a warp at these time
      // scales.  So ONE warp polls, in the order the tensor pipe completes things (c7(0) c7(1) c1(0) c7(2) c1(1) ...), and releases
      // the role's warps from a named barrier.  conv7(i) complete also means ring slot i % ns may be refilled.
      const int pre = q < ns ? q : ns;
      for (int j = 0; j < pre; ++j) issue_tile(j);
      for (int i = 0; i <= q; ++i) {
        if (i < q) {
          mbar_wait(&acc1_full[i & 1], (uint32_t)((i >> 1) & 1));
          named_bar_arrive(R_BAR_E1 + (i & 1), R_BAR_E1_COUNT);
          if (i + ns < q) issue_tile(i + ns);
        }
        if (i >= 1) {
          mbar_wait(&acc2_full[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));
          named_bar_arrive(R_BAR_E2 + ((i - 1) & 1), R_BAR_E2_COUNT);
        }
      }
    } else {
      for (int j = 0; j < q; ++j) {
        if (j >= ns) mbar_wait_backoff(&c7_done[j % ns], (uint32_t)(((j - ns) / ns) & 1), p.sleep_ns);   // conv7 of the slot's previous tile is over
        issue_tile(j);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint64_t desc_fixed = ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29)) << 32) | (1u << 16);   // SBO 512 B, v1, SWIZZLE_64B
      const uint32_t t_u32 = smem_u32(t_ring), w7_u32 = smem_u32(w7s), w1_u32 = smem_u32(w1s);
      const uint32_t idesc = p.idesc7;
      const uint64_t dTap = (uint64_t)(p.dil * 4);             // dil rows x 64 B, in 16-byte units
      const uint32_t tsub16 = p.tsub_bytes >> 4;
      mbar_wait(w_full, 0);
      const uint32_t wo_u32 = smem_u32(wout_s);
      for (int i = 0; i <= q + (kTail ? 2 : 0); ++i) {
        if (i < q) {
          const int slot = i % ns;
          mbar_wait(&a_ready[slot], (uint32_t)((i / ns) & 1));
          tc_fence_after();
          R_STAMP(4, i);
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((i & 1) * 128);
            uint64_t ad = desc_fixed | (uint64_t)((t_u32 + (uint32_t)(slot * rs + hb - p.halo) * 64u) >> 4);
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
          R_STAMP(5, i);
        }
        if (kTail && i >= 3) {
          // outConv partial products of tile t = i - 3: P[256 x 16] = a[256 x 96] (tensor memory, packed by epilogue 2 over the drained
          // conv1 accumulator) x [hi taps | lo taps]^T, into the 16 spare columns behind that accumulator.  Issued three tiles late --
          // epilogue 2 of tile t finished long ago, so this wait never stalls the issuer -- and right BEFORE conv1 of tile t + 2, which
          // overwrites the slot: the tensor pipe executes in order, so conv1 needs no "slot drained" hand-shake in tail mode.
          const int t = i - 3;
          mbar_wait(&a3_ready[t & 1], (uint32_t)((t >> 1) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t slot_t = tmem_base + 256u + (uint32_t)((t & 1) * 128);
            const uint64_t wd = desc_fixed | (uint64_t)(wo_u32 >> 4);
#pragma unroll
            for (int c = 0; c < R_SUB; ++c) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int j16 = 2 * c + k;
                const uint32_t a_col = (uint32_t)(j16 < 3 ? 8 * j16 : 48 + 8 * (j16 - 3));
                tc_mma_f16_2sm_ts(slot_t + 96u, slot_t + a_col, wd + (uint64_t)(c * (512 >> 4) + 2 * k), p.idesc3, (uint32_t)(c | k));
              }
            }
            tc_commit_2sm(&acc3_full[t & 1], mc_mask);
          }
          __syncwarp();
        }
        if (i >= 1 && i <= q) {
          const int t = i - 1;
          mbar_wait(&c_ready[t & 1], (uint32_t)((t >> 1) & 1));
          if (t >= 2 && !kTail) mbar_wait(&acc2_free[t & 1], (uint32_t)(((t - 2) >> 1) & 1));
          tc_fence_after();
          R_STAMP(6, t);
          if (elect_one()) {
            // epilogue 1 packed conv1's operand over the drained conv7 accumulator: channels 0..47 in columns 0..23,
            // channels 48..95 in columns 48..71 of the slot (each epilogue warp writes inside the columns it has read)
            const uint32_t d_tmem = tmem_base + 256u + (uint32_t)((t & 1) * 128), a_tmem = tmem_base + (uint32_t)((t & 1) * 128);
            const uint64_t wd = desc_fixed | (uint64_t)(w1_u32 >> 4);
#pragma unroll
            for (int c = 0; c < R_SUB; ++c) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int j16 = 2 * c + k;                       // 16-channel step
                const uint32_t a_col = (uint32_t)(j16 < 3 ? 8 * j16 : 48 + 8 * (j16 - 3));
                tc_mma_f16_2sm_ts(d_tmem, a_tmem + a_col, wd + (uint64_t)(c * (R_WBLK >> 4) + 2 * k), idesc, (uint32_t)(c | k));
              }
            }
            tc_commit_2sm(&acc2_full[t & 1], mc_mask);
          }
          __syncwarp();
          R_STAMP(7, t);
        }
      }
    }
  } else if (warp >= R_CTRL && warp < R_CTRL + R_PW) {
    // ================= pass-1 warps: snake1 in place on every freshly landed tile =================
    // NP warps per 32-channel block, each a contiguous range of the tile's sixteen 8-row groups, two groups per iteration
    // (independent chunks: ILP).  A lane's 8 channels are fixed, so its 16 SnakeBeta constants live in registers.  A quarter warp
    // touches 2 rows x 64 B = 128 contiguous bytes: conflict-free.  The last hb rows of a tile are also written into the halo rows
    // of the NEXT slot (unless the next tile opens a strip and brings its halo by TMA).
    const int pw = warp - R_CTRL, pc = pw % R_SUB, par = pw / R_SUB, kch = lane & 3;
    constexpr int NP = R_PW / R_SUB;                              // warps per 32-channel block
    const int tg_lo = (par * (R_BM / 8)) / NP, tg_hi = ((par + 1) * (R_BM / 8)) / NP;
    float ea1[8], ib1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ea1[e] = __ldg(p.ea1 + pc * 32 + kch * 8 + e); ib1[e] = __ldg(p.ib1 + pc * 32 + kch * 8 + e); }
    // slot rows and 128 are multiples of 8, so the 64-byte swizzle term ((row >> 1) & 3) depends on the lane only
    const uint32_t lane_off = (uint32_t)(lane >> 2) * 64u + (uint32_t)((kch ^ ((lane >> 3) & 3)) << 4);
    const uint32_t sub_u32 = smem_u32(t_ring) + (uint32_t)pc * p.tsub_bytes + lane_off;
    int slot = 0, cnt = 0;                                          // j % ns, j / ns
    Walker w;
    w.init(p, g0);
    for (int j = 0; j < q; ++j) {
      int nslot = slot + 1, ncnt = cnt;
      if (nslot == ns) { nslot = 0; ++ncnt; }
      const bool live = w.live(p), first = w.first;
      w.next(p);                                                  // now describes tile j+1
      const bool copy_tail = live && j + 1 < q && w.live(p) && !w.first;
      if (pw == 0) R_STAMP(1, j);
      mbar_wait_backoff(&t_full[slot], (uint32_t)(cnt & 1), p.sleep_ns);
      if (pw == 0) R_STAMP(2, j);
      bool tail_ok = !(copy_tail && j >= ns - 1);          // else: conv7 of the next slot's previous tile may still read its halo rows
      if (live) {
        const uint32_t slot_u32 = sub_u32 + (uint32_t)(slot * rs) * 64u;
        const uint32_t nslot_u32 = sub_u32 + (uint32_t)(nslot * rs - R_BM) * 64u;
        const int G_end = hb / 8 + tg_hi;
        for (int G = (par == 0 && first) ? 0 : hb / 8 + tg_lo; G < G_end; G += 2) {
          const bool two = G + 1 < G_end;
          if (!tail_ok && (G + 2) * 8 > R_BM) {                     // first iteration that touches the tail (groups ascend)
            mbar_wait_backoff(&c7_done[nslot], (uint32_t)((ncnt ^ 1) & 1), p.sleep_ns);
            tail_ok = true;
          }
          const uint32_t addr0 = slot_u32 + (uint32_t)G * 512u, naddr0 = nslot_u32 + (uint32_t)G * 512u;
          uint4 u[2];
          u[0] = lds128(addr0);
          if (two) u[1] = lds128(addr0 + 512u); else u[1] = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t in[4] = {u[g].x, u[g].y, u[g].z, u[g].w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = Cvt<T16>::unpack(in[e]);
              const float s0 = __sinf(f.x * ea1[2 * e]), s1 = __sinf(f.y * ea1[2 * e + 1]);
              o[e] = Cvt<T16>::pack(fmaf(ib1[2 * e], s0 * s0, f.x), fmaf(ib1[2 * e + 1], s1 * s1, f.y));
            }
            const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
            if (g == 0 || two) {
              sts128(addr0 + 512u * g, v);
              if (copy_tail && (G + g) * 8 >= R_BM) sts128(naddr0 + 512u * g, v);   // the last hb rows of the tile (an 8-row group is all tail or none)
            }
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (pw == 0) R_STAMP(3, j);
      if (pw >= 2 && pw < 6) R_STAMP(10 + pw, j);       // debug: when do the other pass-1 warps finish?
      if (lane == 0) arrive_leader(&a_ready[slot], rank);
      slot = nslot; cnt = ncnt;
    }
  } else if (warp >= R_CTRL + R_PW && warp < R_CTRL + R_PW + R_E1W) {
    // ================= epilogue-1 warps (two per TMEM lane quarter, 48 channels each): acc1 + b7 -> snake2 -> conv1's operand =================
    // Column-sliced TMEM access (16x256b): a lane sees columns {2(lane%4), +1, +8, +9} of every 16, so its 12 x (b7, ea2, ib2) stay in
    // registers for the whole kernel.  The packed operand is written (16x128b) over accumulator columns that have already been read.
    const int quarter = warp & 3, col_base = ((warp - R_CTRL - R_PW) >> 2) * 48, c0 = 2 * (lane & 3);
    float kb[3][4], ke[3][4], ki[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int col = col_base + 16 * m + c0 + (jj & 1) + 8 * (jj >> 1);
        kb[m][jj] = __ldg(p.b7 + col); ke[m][jj] = __ldg(p.ea2 + col); ki[m][jj] = __ldg(p.ib2 + col);
      }
    Walker w;
    w.init(p, g0);
    for (int t = 0; t < q; ++t) {
      const bool live = w.live(p);
      if (p.bridge) named_bar_sync(R_BAR_E1 + (t & 1), R_BAR_E1_COUNT);
      else mbar_wait_backoff(&acc1_full[t & 1], (uint32_t)((t >> 1) & 1), p.sleep_ns);
      tc_fence_after();
      if (warp == R_CTRL + R_PW) R_STAMP(8, t);
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((t & 1) * 128 + col_base);
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        uint32_t a[2][8];
        tc_ld_16x256b_x2(tbase + 16u * (uint32_t)m, a[0]);                    // lanes +0 .. +15 of the quarter
        tc_ld_16x256b_x2(tbase + (16u << 16) + 16u * (uint32_t)m, a[1]);      // lanes +16 .. +31
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int lh = 0; lh < 2; ++lh) {
            uint32_t pk[4];
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {          // pr = 2 * (column group j) + (row + 8): registers a[4j + 2 (row+8) + {0, 1}]
              const int j = pr >> 1;
              float v[2];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const float x = __uint_as_float(a[lh][4 * j + 2 * (pr & 1) + k]) + kb[m][2 * j + k];
                const float sn = __sinf(x * ke[m][2 * j + k]);
                v[k] = fmaf(ki[m][2 * j + k], sn * sn, x);
              }
              pk[pr] = Cvt<T16>::pack(v[0], v[1]);    // store order: {row, col grp 0}, {row+8, grp 0}, {row, grp 1}, {row+8, grp 1}
            }
            tc_st_16x128b_x2(tbase + ((uint32_t)(16 * lh) << 16) + 8u * (uint32_t)m, pk);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (warp == R_CTRL + R_PW) R_STAMP(9, t);
      if (lane == 0) arrive_leader(&c_ready[t & 1], rank);
      w.next(p);
    }
  } else if (warp >= R_CTRL + R_PW + R_E1W) {
    // ================= epilogue-2 warps (two per TMEM lane quarter, 48 channels each): acc2 + b1 + X -> X' (or snake3(X')) =================
    // The residual rows (L2 hits: the tile went through L2 a few microseconds ago) come in by TMA into the warp's staging buffer, X' is
    // written over them and leaves by TMA: ONE load and ONE store of a 32 x 48 box per warp and tile.  (Direct 16-byte global accesses
    // with a 192-byte row pitch cost 32 LSU wavefronts per instruction; splitting the box into 16- or 24-channel boxes made the
    // staging conflict-free but the 3-6x TMA instructions per tile cost the warp more than the conflicts: profiles/r2_resunit.md.)
    // TWO buffers per warp, alternating by tile.  With one, the chain [X' store drains the buffer (~2000 cycles in the TMA queue) ->
    // residual load of the next tile -> compute] was serial: the warps worked 2300 cycles and stalled 2100 per tile, and that chain --
    // not the tensor pipe, not pass 1 -- set the kernel's pace (clock64 timeline).  Now the next tile's residual is requested right
    // after this tile's math, into the buffer whose store -- a whole tile ago -- has drained meanwhile.
    const int e2w = warp - R_CTRL - R_PW - R_E1W;
    const int quarter = warp & 3, col_base = (e2w >> 2) * 48;
    const uint32_t cst_u32 = smem_u32(cst);
    uint8_t* my_stg = stg + (size_t)e2w * R_STG_TILE * (kTail ? 1u : 2u);
    uint64_t* my_xbar = xbar + 2 * e2w;
    // Tail mode (the decoder's last unit): the output a = snake3(X') stays on chip.  It is packed back into tensor memory over the
    // conv1 accumulator columns just read (as epilogue 1 does for conv1's operand); the issuer multiplies it with outConv's taps three
    // tiles later, and the 16 partial products per row -- 64 B instead of the 192 B of activations -- are stored from here, two
    // tiles late, straight from registers (a warp's 32 rows are 2 KB contiguous in P).
    long long h_off[2] = {-1, -1};                                  // where this lane's row of tiles u - 1, u - 2 goes in P (< 0: nowhere)
    auto store_partials = [&](int t, long long off) {               // P of tile t (its MMA was issued at issuer step t + 3)
      if (col_base != 0) return;                                    // the two warps of a lane quarter share the rows: one of them stores
      mbar_wait_backoff(&acc3_full[t & 1], (uint32_t)((t >> 1) & 1), p.sleep_ns);
      tc_fence_after();
      uint32_t r[1][16];
      tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + (t & 1) * 128 + 96), r[0]);
      tc_wait_ld();
      tc_fence_before();
      if (off >= 0) {
        float4* dst = (float4*)(p.p_out + off);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_float4(__uint_as_float(r[0][4 * c]), __uint_as_float(r[0][4 * c + 1]), __uint_as_float(r[0][4 * c + 2]), __uint_as_float(r[0][4 * c + 3]));
      }
    };
    Walker w;
    w.init(p, g0);
    if (q > 0 && w.live(p) && lane == 0) {
      mbar_expect_tx(&my_xbar[0], R_STG_TILE);
      tma_load_3d(my_stg, &map_res, &my_xbar[0], col_base, w.t0 + quarter * 32, w.b);
    }
    for (int u = 0; u < q; ++u) {
      const bool row_ok = w.live(p);                                // warp-uniform
      const int row0 = w.t0 + quarter * 32, wb = w.b, bf = u & 1;
      w.next(p);                                                    // now describes tile u + 1
      const bool next_ok = u + 1 < q && w.live(p);
      const int sb_ = !kTail ? bf : 0;                      // staging buffer of this tile
      const uint32_t my_row = smem_u32(my_stg) + (uint32_t)(R_STG_TILE * sb_) + (uint32_t)lane * 96u;
      if (p.bridge) named_bar_sync(R_BAR_E2 + (u & 1), R_BAR_E2_COUNT);
      else mbar_wait_backoff(&acc2_full[u & 1], (uint32_t)((u >> 1) & 1), p.sleep_ns);
      tc_fence_after();
      if (warp == R_CTRL + R_PW + R_E1W) R_STAMP(10, u);
      if (kTail && u >= 2) store_partials(u - 2, h_off[1]);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(256 + (u & 1) * 128 + col_base);
      if (row_ok) mbar_wait_backoff(&my_xbar[sb_], !kTail ? (uint32_t)((u >> 1) & 1) : (uint32_t)(u & 1), p.sleep_ns);
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        uint32_t r[1][16];
        uint4 rr[6];
        tc_ld16(taddr + 16u * (uint32_t)m, r[0]);
        if (row_ok) { rr[2 * m] = lds128(my_row + 32u * (uint32_t)m); rr[2 * m + 1] = lds128(my_row + 32u * (uint32_t)m + 16u); }
        tc_wait_ld();
        if (m == 2 && !kTail) {   // the accumulator has been read completely (tail mode: the issuer orders the slot's reuse itself)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&acc2_free[u & 1], rank);
        }
        if (row_ok) {
          const int col = col_base + 16 * m;
          const uint32_t sb = cst_u32 + 4u * (uint32_t)col, se = sb + 4u * R_C, si = se + 4u * R_C;
          uint32_t o[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 b0 = lds4f(sb + 32u * h), b1 = lds4f(sb + 32u * h + 16u);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const uint4 rq = rr[2 * m + h];
            const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
            if (p.out_snake) {
              const float4 e0 = lds4f(se + 32u * h), e1 = lds4f(se + 32u * h + 16u);
              const float4 i0 = lds4f(si + 32u * h), i1 = lds4f(si + 32u * h + 16u);
              const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
              const float ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
              float v[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = Cvt<T16>::unpack(rw[e]);
                v[2 * e] = __uint_as_float(r[0][8 * h + 2 * e]) + bb[2 * e] + f.x;
                v[2 * e + 1] = __uint_as_float(r[0][8 * h + 2 * e + 1]) + bb[2 * e + 1] + f.y;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float sn = __sinf(v[e] * ee[e]);
                v[e] = fmaf(ii[e], sn * sn, v[e]);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) o[4 * h + e] = Cvt<T16>::pack(v[2 * e], v[2 * e + 1]);
            } else {
              // the branch output is rounded to 16 bits, then added to the 16-bit stream with a packed add
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[4 * h + e] = Cvt<T16>::add2(Cvt<T16>::pack(__uint_as_float(r[0][8 * h + 2 * e]) + bb[2 * e],
                                                             __uint_as_float(r[0][8 * h + 2 * e + 1]) + bb[2 * e + 1]), rw[e]);
            }
          }
          if (kTail) {
            // packed a, over accumulator columns that are already in registers: channels col_base + 16 m .. + 15 -> columns col_base + 8 m .. + 7
            tc_st8(taddr + 8u * (uint32_t)m, o);
          } else {
            sts128(my_row + 32u * (uint32_t)m, make_uint4(o[0], o[1], o[2], o[3]));
            sts128(my_row + 32u * (uint32_t)m + 16u, make_uint4(o[4], o[5], o[6], o[7]));
          }
        }
      }
      if (kTail) {
        tc_wait_st();
        tc_fence_before();
        __syncwarp();                                              // every lane has read its residual row and written its operand row
        if (lane == 0) {
          arrive_leader(&a3_ready[u & 1], rank);
          if (next_ok) {                                           // nothing is stored from the staging buffer: one buffer, refilled at once
            mbar_expect_tx(&my_xbar[0], R_STG_TILE);
            tma_load_3d(my_stg, &map_res, &my_xbar[0], col_base, w.t0 + quarter * 32, w.b);
          }
        }
        h_off[1] = h_off[0];
        h_off[0] = (row_ok && row0 + lane < p.slot_rows) ? (long long)wb * p.p_bstride + (long long)(row0 + lane) * 16 : -1;
      } else if (row_ok) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (next_ok) {
            tma_store_wait_read0();                                  // the store of tile u - 1 (issued a whole tile ago) has drained the other buffer
            mbar_expect_tx(&my_xbar[bf ^ 1], R_STG_TILE);
            tma_load_3d(my_stg + R_STG_TILE * (bf ^ 1), &map_res, &my_xbar[bf ^ 1], col_base, w.t0 + quarter * 32, w.b);
          }
          tma_store_3d(&map_out, my_stg + R_STG_TILE * bf, col_base, row0, wb);
          tma_store_commit();
        }
      }
      if (warp == R_CTRL + R_PW + R_E1W) R_STAMP(11, u);
    }
    if (kTail) {
      if (q >= 2) store_partials(q - 2, h_off[1]);
      if (q >= 1) store_partials(q - 1, h_off[0]);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
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
  const bool tail = p.p_out != nullptr;
  if (tail && (!p.wout || !p.ea3)) return cudaErrorInvalidValue;
  q.tail = tail ? 1 : 0;
  q.stg_bufs = tail ? 1 : 2;
  q.p_out = p.p_out; q.slot_rows = slot_rows; q.p_bstride = (long long)slot_rows * 16;
  const size_t fixed = R_W7_BYTES + R_W1_BYTES + (tail ? R_WOUT_BYTES : 0) + (size_t)R_E2W * R_STG_TILE * q.stg_bufs + R_CST + 512 + 1024;
  auto tsub = [&](int n) { return ((uint32_t)(n * (q.hb + R_BM)) * 64u + 511u) & ~511u; };   // SWIZZLE_64B repeats every 512 B
  // Three ring slots: a slot is busy from the TMA issue to the end of conv7 of its tile (load ~1.5 k + pass 1 ~2.7 k + conv7 ~2 k cycles),
  // about two tile periods.  The fourth slot's 26-35 KB now hold the second epilogue-2 staging buffer.
  static const int slots_env = []() { const char* e = getenv("Q3TTS_RES_SLOTS"); return e ? atoi(e) : 3; }();
  q.nslots = (slots_env == 4 && (size_t)R_SUB * tsub(4) + fixed <= 227 * 1024) ? 4 : 3;
  q.tsub_bytes = tsub(q.nslots);
  q.out = p.out;
  const uint32_t fmt = op_dtype == DT_F16 ? 0u : 1u;
  q.idesc7 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(R_C >> 3) << 17) | ((uint32_t)((2 * R_BM) >> 4) << 24);
  q.idesc3 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)((2 * R_BM) >> 4) << 24);
  q.b7 = p.b7; q.ea1 = p.ea1; q.ib1 = p.ib1; q.ea2 = p.ea2; q.ib2 = p.ib2; q.b1 = p.b1; q.ea3 = p.ea3; q.ib3 = p.ib3;
  q.out_snake = p.ea3 != nullptr;
  q.x_in = p.x_in; q.x_bstride = (long long)slot_rows * R_C;
  q.dbg = (long long*)p.dbg;
  static const int sleep_env = []() { const char* e = getenv("Q3TTS_RES_SLEEP"); return e ? atoi(e) : 100; }();
  q.sleep_ns = (uint32_t)sleep_env;
  static const int bridge_env = []() { const char* e = getenv("Q3TTS_RES_BRIDGE"); return e ? atoi(e) : 1; }();
  q.bridge = bridge_env;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles_bound = (g.valid_frames * p.rows_per_frame + R_BM - 1) / R_BM + g.B;   // sum of per-utterance ceilings <= this
  int grid = (int)std::min<long long>(sms / 2 * 2, (tiles_bound + 1) / 2 * 2);
  grid = std::max(grid, 2);
  q.tiles_per_cta = (int)((tiles_bound + grid - 1) / grid);
  CUtensorMap map_main, map_halo, map_w7, map_w1, map_res, map_out, map_wout;
  auto act_map = [&](CUtensorMap* m, const void* base, cuuint32_t cols, cuuint32_t rows) -> bool {
    cuuint64_t dims[3] = {(cuuint64_t)R_C, (cuuint64_t)slot_rows, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)R_C * 2, (cuuint64_t)slot_rows * R_C * 2};
    cuuint32_t box[3] = {cols, rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (cols == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto w_map = [&](CUtensorMap* m, const void* base, int rows, cuuint32_t box_rows) -> bool {
    cuuint64_t dims[2] = {(cuuint64_t)R_C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)R_C * 2};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!act_map(&map_main, p.x_in, 32, R_BM) || !act_map(&map_halo, p.x_in, 32, (cuuint32_t)q.hb) ||
      !w_map(&map_w7, p.w7, 7 * R_C, 48) || !w_map(&map_w1, p.w1, R_C, 48) ||
      !act_map(&map_res, p.x_in, 48, 32) || !act_map(&map_out, tail ? p.x_in : p.out, 48, 32) ||   // tail mode stores no activations: any valid map
      !w_map(&map_wout, tail ? p.wout : p.w1, 16, 8))
    return cudaErrorInvalidValue;
  const size_t smem = (size_t)R_SUB * q.tsub_bytes + fixed;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(R_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const CUtensorMap, Res96Params);
  static const KernelFn kFns[2][2][2] = {   // [bf16][debug stamps][tail mode]
      {{resunit96_kernel<__half, false, false>, resunit96_kernel<__half, false, true>},
       {resunit96_kernel<__half, true, false>, resunit96_kernel<__half, true, true>}},
      {{resunit96_kernel<__nv_bfloat16, false, false>, resunit96_kernel<__nv_bfloat16, false, true>},
       {resunit96_kernel<__nv_bfloat16, true, false>, resunit96_kernel<__nv_bfloat16, true, true>}}};
  static tc::PerDeviceOnce optin;
  const cudaError_t oe = optin.ensure([]() {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 8 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute((const void*)kFns[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e;
  });
  if (oe != cudaSuccess) return oe;
  return cudaLaunchKernelEx(&cfg, kFns[op_dtype == DT_F16 ? 0 : 1][q.dbg ? 1 : 0][tail ? 1 : 0], map_main, map_halo, map_w7, map_w1, map_res, map_out,
                            map_wout, q);
}

}  // namespace q3
