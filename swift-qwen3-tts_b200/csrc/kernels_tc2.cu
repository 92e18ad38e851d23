// tcgen05 multi-tap GEMM: HALO tiles + CTA-pair MMA + lean issue loops -- the tensor-core engine behind every conv, linear and
// transposed conv of the decoder in the 16-bit modes (see kernels.cuh for the multi-tap GEMM definition).
//
// The first version of this kernel re-read the activation tile once per tap and the weight tile once per 128-row tile; ncu
// showed block3's conv7 moving 18 GB L2->SM for 1.1 GB of input (profiles/r1_v1_*).  Here:
//  * the A operand of ALL taps of a 64-channel block comes from ONE TMA box of 128 + (taps-1)*dil rows (the halo
//    tile).  Tap j is the same smem tile viewed from row j*dil: its UMMA descriptor simply starts j*dil*128 bytes
//    later (the 128B swizzle is a function of the absolute smem address, which TMA and the MMA unit share);
//  * the two CTAs of a cluster form a cta_group::2 pair on consecutive M tiles of the same N tile: ONE tcgen05.mma
//    spans both SMs (M = 256), each CTA stages its own rows of A and HALF of the weight tile;
//  * weights that fit stay RESIDENT in smem for the whole kernel (block 3's convs, block 2's 1x1): every stage of the
//    weight ring is then loaded exactly once instead of once per tile, which takes ~2/3 of the L2->SM traffic away;
//  * the producer and MMA warps run WARP-UNIFORM loops (tile validity goes through a vote, so ptxas keeps descriptors,
//    coordinates and barrier addresses in uniform registers) and elect one lane only around the TMA / MMA / commit
//    instructions.  The first version wrapped the loops in `if (lane == 0)`: every tcgen05.mma was then preceded by a
//    VOTEU/ELECT/5xR2UR.BROADCAST/BRA.U.ANY uniformisation loop and the tensor pipe idled 84 % of the time on N = 96;
//  * partial K blocks (Cin % 64 != 0, e.g. 96) issue only the k-steps that hold data instead of multiplying zeros;
//  * small-N GEMMs (BN <= 128) use a 4-deep TMEM accumulator ring instead of 2, so the epilogue may lag three tiles;
//  * two epilogues, compiled as separate kernels: the BLOCK epilogue (bias [+ residual] [+ stream out] + SnakeBeta
//    operand out, all 16-bit; 12 warps, per-warp TMA stores, per-column constants staged in smem) used by the decoder
//    blocks, and the GENERIC one (GELU / SwiGLU / layer scale / fp32 streams / taps; 8 warps).
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

constexpr int T2_BM = 128, T2_BK = 64, T2_MAX_NA = 8, T2_MAX_NW = 16, T2_MAX_ACC = 4;
constexpr uint32_t T2_CHUNK_BYTES = 128 * 32;            // generic epilogue: one staged chunk = 128 rows x 16 columns x 2 B
constexpr uint32_t T2_STAGING_BYTES = 32 * 1024;         // generic epilogue: 2 groups x 2 outputs x 2 buffers x 4 KB
constexpr int T2_THREADS_GENERIC = 384, T2_THREADS_BLOCK = 512, T2_EPI_WARPS = 8, T2_EPI_WARPS_BLOCK = 12;
constexpr uint32_t T2_TMEM_COLS = 512;
constexpr int T2_CST_MAX_N = 768;                        // per-column constants (bias, snake) staged in smem up to this N

struct Tc2Params {
  int B, Tmax, rows_per_frame;
  const int* len_frames;
  int N, BN, Cin, taps, dil, ncb, halo;          // halo = (taps-1)*dil rows in front of every tile
  int tiles_per_utt, n_tiles, m_tiles_total, cs; // cs = cluster size (1 or 2)
  uint32_t idesc, a_stage_bytes, a_tx_bytes, w_stage_bytes;
  int na, nw;                                    // ring depths (A halo tiles, W stages)
  int wg, ngroups;                               // taps per W stage; W stages per 64-channel block
  uint32_t w_tap_bytes;                          // bytes of one tap's weight tile in this CTA (BN/cs rows x 128 B)
  int w_resident;                                // every W stage of the kernel has its own slot: loaded once
  int nacc, acc_stride;                          // TMEM accumulator ring: depth (2 or 4) and column stride
  int nmma;                                      // MMA issuer warps (1 or 2)
  int cst_staged;                                // bias / snake constants of all N columns live in smem
  int tma_y, tma_a;                              // 16-bit outputs leave through smem staging + TMA store
  uint32_t staging_bytes;                        // smem reserved for the output staging buffers
  int stg_bufs;                                  // block epilogue: staging buffers per warp (2: a chunk is staged while the previous one drains)
  uint32_t div_nt_m, div_nt_s, div_tpu_m, div_tpu_s;   // magic numbers: x / n_tiles, x / tiles_per_utt (x < 2^31)
  const float* bias; int act;
  const void* res; int ldres; long long res_bstride;
  const float* scale;
  void* out_y; int ldy; long long y_bstride;
  void* out_a; int lda_out; long long ao_bstride;
  const float* snake_ea; const float* snake_ib;
  float* out_tap; int ldt; long long tap_bstride;
  // fused residual unit (block epilogue kernel only): this GEMM is conv7 (+ bias + snake -> a 128B-swizzled operand tile in
  // smem), followed by the unit's 1x1 conv from that tile (W1 resident in smem) with the usual residual epilogue
  // The 1x1 conv runs as two N halves through ONE extra accumulator of BN/2 columns (TMEM: 2 x BN + BN/2 <= 512), issued
  // in the middle of the NEXT tile's conv7, so that neither epilogue ever holds up the tensor pipe.
  int a_tmem;                                    // conv1's operand lives in TENSOR MEMORY, packed over the drained conv7 accumulator
  int fuse, ncb2, acc2_col, c1_after1;           // c1_after1: 64-channel block of the next conv7 after which half 1 of conv1 is issued (half 0: after block 0)
  uint32_t idesc2;                               // N = BN/2
  uint32_t w1_blk_bytes;                         // one (N half, 64-channel block) of this CTA's part of W1: (BN/4) rows x 128 B
  const float* bias2; const float* ea2; const float* ib2;
  const void* res2; void* out_y2; void* out_a2; long long o2_bstride;
};

__device__ __forceinline__ void ld4(const float* p, float (&v)[16], int i) {
  const float4 t = __ldg((const float4*)p + i);
  v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
}

// `ntap` taps x NK k-steps of one W stage.  Descriptors advance by 2 (32 bytes) per k-step and by dA / dW per tap;
// only the very first MMA of a tile runs with accumulate = 0.
template <bool kPair, int NK>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint64_t ad, uint64_t wd, uint64_t dA, uint64_t dW, uint32_t idesc,
                                           uint32_t first, int ntap) {
#pragma unroll 1
  for (int j = 0; j < ntap; ++j) {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const uint32_t accumulate = k == 0 ? (first | (uint32_t)j) : 1u;
      if (kPair) tc_mma_f16_2sm(d_tmem, ad + 2 * k, wd + 2 * k, idesc, accumulate);
      else tc_mma_f16(d_tmem, ad + 2 * k, wd + 2 * k, idesc, accumulate);
    }
    ad += dA; wd += dW;
  }
}

// kPair: the two CTAs of the cluster form a cta_group::2 pair -- ONE tcgen05.mma spans both SMs (M = 256), each CTA
// stages its own 128(+halo) rows of A and HALF of the weight tile, so weight ingest and B-operand smem reads per SM
// halve.  Only the leader (rank 0) issues MMAs; both CTAs run producers and epilogues (each on its own TMEM half).
template <typename T16, bool kPair, bool kBlockEpi>
__global__ void __launch_bounds__(kBlockEpi ? T2_THREADS_BLOCK : T2_THREADS_GENERIC, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_o,
                     const __grid_constant__ CUtensorMap map_w2, Tc2Params p, int y_is_f32) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  const int T2_NA = p.na, T2_NW = p.nw;
  uint8_t* w_ring = smem + (size_t)T2_NA * p.a_stage_bytes;
  uint8_t* staging = w_ring + (size_t)T2_NW * p.w_stage_bytes;          // 1024-aligned (all stage sizes are)
  uint8_t* c_tile = staging + p.staging_bytes;                          // fused unit: conv1's operand, [ncb2][128 rows][128 B]
  uint8_t* w1s = c_tile + ((p.fuse && !p.a_tmem) ? (size_t)p.ncb2 * 16384 : 0);        // fused unit: [N half][ncb2][BN/4 rows][128 B]
  float* cst = (float*)(w1s + (p.fuse ? (size_t)2 * p.ncb2 * p.w1_blk_bytes : 0));   // [bias | ea | ib] x N when staged (x2 when fused)
  uint64_t* bars = (uint64_t*)((uint8_t*)cst + (p.cst_staged ? (size_t)(p.fuse ? 6 : 3) * p.N * 4 : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + T2_MAX_NA;
  uint64_t* w_full = a_empty + T2_MAX_NA;
  uint64_t* w_empty = w_full + T2_MAX_NW;
  uint64_t* tmem_full = w_empty + T2_MAX_NW;
  uint64_t* tmem_empty = tmem_full + T2_MAX_ACC;
  uint64_t* tmem_full2 = tmem_empty + T2_MAX_ACC;   // [2] fused unit: N half h of conv1 has completed
  uint64_t* c_ready = tmem_full2 + 2;               // fused unit: the operand tile of conv1 is written (both CTAs)
  uint64_t* w1_full = c_ready + 1;
  uint64_t* acc2_empty = w1_full + 1;               // fused unit: the conv1 accumulator has been drained
  uint64_t* xbar = acc2_empty + 1;                  // [12] fused unit: a warp's residual chunk has landed in its staging buffer
  uint32_t* tmem_ptr = (uint32_t*)(xbar + T2_EPI_WARPS_BLOCK);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cs = p.cs;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1);
  const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
  const int groups = (p.m_tiles_total + cs - 1) / cs;
  const int items = groups * p.n_tiles;
  constexpr int kEpiWarps = kBlockEpi ? T2_EPI_WARPS_BLOCK : T2_EPI_WARPS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (p.tma_y) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    if (p.tma_a) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < T2_NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < T2_NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], kPair ? 1u : (uint32_t)cs); }
    for (int i = 0; i < p.nacc; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], (uint32_t)((kPair ? 2 : 1) * kEpiWarps)); }
    for (int i = 0; i < 2; ++i) mbar_init(&tmem_full2[i], 1);
    mbar_init(c_ready, (uint32_t)((kPair ? 2 : 1) * kEpiWarps));
    mbar_init(w1_full, 1);
    mbar_init(acc2_empty, (uint32_t)((kPair ? 2 : 1) * kEpiWarps));
    for (int i = 0; i < T2_EPI_WARPS_BLOCK; ++i) mbar_init(&xbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(T2_TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(T2_TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (kBlockEpi && p.cst_staged && warp >= 4) {
    for (int i = (int)threadIdx.x - 128; i < p.N; i += kEpiWarps * 32) {
      cst[i] = __ldg(p.bias + i);
      cst[p.N + i] = p.snake_ea ? __ldg(p.snake_ea + i) : 0.f;
      cst[2 * p.N + i] = p.snake_ib ? __ldg(p.snake_ib + i) : 0.f;
      if (p.fuse) {
        cst[3 * p.N + i] = __ldg(p.bias2 + i);
        cst[4 * p.N + i] = p.ea2 ? __ldg(p.ea2 + i) : 0.f;
        cst[5 * p.N + i] = p.ib2 ? __ldg(p.ib2 + i) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive / multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Everything above touched only shared / tensor memory and weights.  The next kernel of the stream may start its own prologue now;
  // this one must not read activations before the previous kernel has completed (kernels.cuh, programmatic dependent launch).
  pdl_trigger();
  pdl_wait();

  // item -> (n tile, this CTA's M tile).  Returns "some CTA of the cluster has rows to produce"; the result and `mine`
  // go through a vote so that the callers' control flow is warp-uniform by construction.  Divisions are multiplications
  // by host-computed magic numbers: every warp of the CTA walks the item list, so this runs 16 times per tile.
  // The rows of utterances vb0 and vb0 + 1 stay in registers: a cluster's consecutive items sit in the same one or two utterances for
  // hundreds of items, and a dependent global load per item and warp was a third of the producer warp's time -- whose instruction
  // stream, not the memory system, set the pace of the transposed convs (profiles/r2_gemm_experiments.md section 3).
  int vb0 = -2, vrows0 = 0, vrows1 = 0, last_nt = 0;
  auto coords = [&](int item, int& b, int& t0, int& n0, bool& mine) -> bool {
    int g = item, nt = 0;
    if (p.n_tiles > 1) {
      g = (int)(__umulhi((uint32_t)item, p.div_nt_m) >> p.div_nt_s);   // n_tiles >= 2
      nt = item - g * p.n_tiles;
    }
    n0 = nt * p.BN;
    last_nt = nt;
    const int mg0 = g * cs;
    const int b0 = p.tiles_per_utt == 1 ? mg0 : (int)(__umulhi((uint32_t)mg0, p.div_tpu_m) >> p.div_tpu_s);
    const int i0 = mg0 - b0 * p.tiles_per_utt;                    // M tile index inside its utterance
    if (b0 != vb0) {
      vb0 = b0;
      vrows0 = b0 < p.B ? __ldg(p.len_frames + b0) * p.rows_per_frame : 0;
      vrows1 = b0 + 1 < p.B ? __ldg(p.len_frames + b0 + 1) * p.rows_per_frame : 0;
    }
    bool any = false;
    mine = false;
    b = p.B; t0 = 0;                 // out-of-range utterance: TMA zero-fills, epilogue stores nothing
    for (int r = 0; r < cs; ++r) {
      if (mg0 + r >= p.m_tiles_total) continue;
      const bool wrap = i0 + r >= p.tiles_per_utt;                // cs <= 2: the pair's second tile may open the next utterance
      const int bb = wrap ? b0 + 1 : b0, tt = (wrap ? 0 : i0 + r) * T2_BM;
      const bool ok = tt < (wrap ? vrows1 : vrows0);
      any |= ok;
      if (r == (int)rank) { b = bb; t0 = tt; mine = ok; }
    }
    mine = __any_sync(0xffffffffu, mine);
    return __any_sync(0xffffffffu, any);
  };

  if (warp == 0) {
    // ================= TMA producer (one elected lane) + residual L2 prefetch (all lanes) =================
    // The producer runs one to two tiles ahead of the epilogue, so pulling this item's residual rows into L2 here
    // turns the epilogue's residual reads from HBM-latency loads into L2 hits.
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0, w_loaded = 0;      // w_loaded: bit s = resident W stage s has been requested
    const uint32_t w_all = p.w_resident ? (uint32_t)((1ull << T2_NW) - 1) : 0xffffffffu;   // non-resident: w_loaded stays 0, never equal
    const int wrows = p.BN / cs;
    const int res_es = y_is_f32 ? 4 : 2;
    const int stages_per_tile = p.ncb * p.ngroups;
    if (kBlockEpi && kPair && p.fuse && elect_one()) {   // the 1x1 conv's weights: resident for the whole kernel
      if (rank == 0) mbar_expect_tx(w1_full, 4u * (uint32_t)p.ncb2 * p.w1_blk_bytes);
      for (int h = 0; h < 2; ++h)
        for (int cb = 0; cb < p.ncb2; ++cb)
          tma_load_2d_2sm(w1s + (size_t)(h * p.ncb2 + cb) * p.w1_blk_bytes, &map_w2, w1_full, cb * T2_BK,
                          h * (p.BN / 2) + (int)rank * (p.BN / 4));
    }
    __syncwarp();
    for (int item = cid; item < items; item += ncl) {
      int b, t0, n0; bool mine;
      if (!coords(item, b, t0, n0, mine)) continue;
      const void* res_any = (kBlockEpi && p.fuse) ? p.res2 : p.res;       // fused unit: the residual belongs to the 1x1 stage
      if (res_any != nullptr && mine) {
        const int valid = __ldg(p.len_frames + b) * p.rows_per_frame;
        const int line_cnt = (p.BN * res_es + 127) / 128;
        const long long rb = (kBlockEpi && p.fuse) ? p.o2_bstride : p.res_bstride;
        const int rl = (kBlockEpi && p.fuse) ? p.N : p.ldres;
        for (int r = lane; r < T2_BM; r += 32) {
          if (t0 + r >= valid) break;
          const char* ptr = (const char*)res_any + ((long long)b * rb + (long long)(t0 + r) * rl + n0) * res_es;
          for (int l = 0; l < line_cnt; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + l * 128));
        }
        __syncwarp();
      }
      if (p.w_resident) sw = last_nt * stages_per_tile;
      for (int cb = 0; cb < p.ncb; ++cb) {
        mbar_wait(&a_empty[sa], pa ^ 1);
        if (elect_one()) {
          if (kPair) {   // both CTAs' boxes complete on the LEADER's barrier
            if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * p.a_tx_bytes);
            tma_load_3d_2sm(a_ring + (size_t)sa * p.a_stage_bytes, &map_a, &a_full[sa], cb * T2_BK, t0 - p.halo, b);
          } else {
            mbar_expect_tx(&a_full[sa], p.a_tx_bytes);
            tma_load_3d(a_ring + (size_t)sa * p.a_stage_bytes, &map_a, &a_full[sa], cb * T2_BK, t0 - p.halo, b);
          }
        }
        if (++sa == T2_NA) { sa = 0; pa ^= 1; }
        if (w_loaded == w_all) continue;            // resident weights: every stage has been requested
        for (int tap0 = 0; tap0 < p.taps; tap0 += p.wg) {
          const int ntap = min(p.wg, p.taps - tap0);
          bool load = true;
          if (p.w_resident) {
            load = !((w_loaded >> sw) & 1u);
            w_loaded |= 1u << sw;
          } else {
            mbar_wait(&w_empty[sw], pw ^ 1);            // every CTA of the cluster has drained this slot
          }
          if (load && elect_one()) {
            uint8_t* wst = w_ring + (size_t)sw * p.w_stage_bytes;
            if (kPair) {   // this CTA's half of each weight tile (rows rank*BN/2 ..), at the SAME smem offsets in both CTAs
              if (rank == 0) mbar_expect_tx(&w_full[sw], (uint32_t)ntap * (uint32_t)p.BN * 128u);
              for (int j = 0; j < ntap; ++j)
                tma_load_2d_2sm(wst + (size_t)j * p.w_tap_bytes, &map_w, &w_full[sw], cb * T2_BK, (tap0 + j) * p.N + n0 + (int)rank * wrows);
            } else {
              mbar_expect_tx(&w_full[sw], (uint32_t)ntap * (uint32_t)p.BN * 128u);
              for (int j = 0; j < ntap; ++j) {
                uint8_t* dst = wst + (size_t)j * p.w_tap_bytes + (size_t)rank * wrows * 128;
                if (cs == 1) tma_load_2d(dst, &map_w, &w_full[sw], cb * T2_BK, (tap0 + j) * p.N + n0);
                else tma_load_2d_mc(dst, &map_w, &w_full[sw], cb * T2_BK, (tap0 + j) * p.N + n0 + (int)rank * wrows, mc_mask);
              }
            }
          }
          if (++sw == T2_NW) { sw = 0; pw ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 || (warp == 3 && p.nmma == 2)) {
    // ================= MMA issuers: warp-uniform loops, one elected lane issues =================
    // Two issuer warps take alternate tiles (each tile has its own TMEM accumulator and its own ring slots, and a
    // tcgen05.commit only tracks the MMAs of the thread that executes it), which doubles the issue rate of small tiles.
    // The single issuing thread is latency-bound on its own instruction stream (uniform-datapath ops cost 5-10 cycles
    // each when dependent), so the per-MMA work is two 64-bit descriptor adds: descriptors are linear in (tap, k-step)
    // and the address field cannot carry out of its 14 bits (smem addresses are < 256 KB).
    if (!kPair || rank == 0) {
      int sa = 0, sw = 0, acc = 0;
      uint32_t pa = 0, pw = 0, pacc = 0;
      const uint64_t desc_fixed = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);   // SBO 1024 B, v1, SWIZZLE_128B, LBO 1
      const int stages_per_tile = p.ncb * p.ngroups;
      const int taps = p.taps, wg = p.wg, ncb = p.ncb, nacc = p.nacc, w_resident = p.w_resident;
      const uint32_t idesc = p.idesc;
      const uint64_t dA = (uint64_t)(p.dil * 8), dW = (uint64_t)(p.w_tap_bytes >> 4);   // per-tap descriptor steps (16-byte units)
      const uint32_t a_ring_u32 = smem_u32(a_ring), w_ring_u32 = smem_u32(w_ring);
      uint32_t w_seen = 0;                       // resident weights: bit s = stage s has been waited for once (it never changes again)
      const int which = warp == 1 ? 0 : 1;
      int turn = 0;
      // fused unit: the two N halves of conv1 of tile k-1 are issued after the first and second 64-channel block of conv7 of
      // tile k: by then epilogue 1 of tile k-1 has written the operand tile, and each half's drain overlaps conv7 MMAs.
      int n_done = 0;
      auto issue_conv1 = [&](int h) {
        if (h == 0) {
          mbar_wait(c_ready, (uint32_t)((n_done - 1) & 1));
          if (n_done == 1) mbar_wait(w1_full, 0u);
        }
        mbar_wait(acc2_empty, (uint32_t)(h ^ 1));      // fill f = 2 (n_done-1) + h needs the drain of fill f-1
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d2 = tmem_base + (uint32_t)p.acc2_col;
          const uint64_t ad = desc_fixed | (uint64_t)(smem_u32(c_tile) >> 4);
          const uint64_t wd = desc_fixed | (uint64_t)((smem_u32(w1s) + (uint32_t)(h * p.ncb2) * p.w1_blk_bytes) >> 4);
          const uint32_t a_t = tmem_base + (uint32_t)(((n_done - 1) & 1) * p.acc_stride);   // the previous tile's accumulator slot
          for (int cb = 0; cb < p.ncb2; ++cb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a2 = ad + (uint64_t)(cb * (16384 >> 4) + 2 * k), w2 = wd + (uint64_t)(cb * (int)(p.w1_blk_bytes >> 4) + 2 * k);
              if (kPair && p.a_tmem) tc_mma_f16_2sm_ts(d2, a_t + (uint32_t)(cb * 64 + k * 8), w2, p.idesc2, (uint32_t)(cb | k));
              else if (kPair) tc_mma_f16_2sm(d2, a2, w2, p.idesc2, (uint32_t)(cb | k));
              else tc_mma_f16(d2, a2, w2, p.idesc2, (uint32_t)(cb | k));
            }
          }
          if (kPair) tc_commit_2sm(&tmem_full2[h], mc_mask); else tc_commit(&tmem_full2[h]);
        }
        __syncwarp();
      };
      for (int item = cid; item < items; item += ncl) {
        int b, t0, n0; bool mine;
        if (!coords(item, b, t0, n0, mine)) continue;
        if (p.nmma == 2) {
          const bool skip = turn != which;
          turn ^= 1;
          if (skip) {   // the other issuer's tile: step over its ring slots and its accumulator
            sa += ncb;
            while (sa >= T2_NA) { sa -= T2_NA; pa ^= 1; }
            if (!w_resident) {
              sw += stages_per_tile;
              while (sw >= T2_NW) { sw -= T2_NW; pw ^= 1; }
            }
            if (++acc == nacc) { acc = 0; pacc ^= 1; }
            continue;
          }
        }
        mbar_wait(&tmem_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        if (w_resident) sw = last_nt * stages_per_tile;
        for (int cb = 0; cb < ncb; ++cb) {
          mbar_wait(&a_full[sa], pa);
          const uint64_t ad_cb = desc_fixed | (uint64_t)((a_ring_u32 + (uint32_t)sa * p.a_stage_bytes) >> 4);
          const int nk = min(T2_BK, p.Cin - cb * T2_BK) / 16;
          for (int tap0 = 0; tap0 < taps; tap0 += wg) {
            const int ntap = min(wg, taps - tap0);
            if (!w_resident) {
              mbar_wait(&w_full[sw], pw);
            } else if (!((w_seen >> sw) & 1u)) {
              mbar_wait(&w_full[sw], 0u);
              w_seen |= 1u << sw;
            }
            tc_fence_after();
            if (elect_one()) {
              uint64_t ad = ad_cb + (uint64_t)tap0 * dA;
              uint64_t wd = desc_fixed | (uint64_t)((w_ring_u32 + (uint32_t)sw * p.w_stage_bytes) >> 4);
              const uint32_t first = (cb | tap0) != 0 ? 1u : 0u;
              if (nk == 4) issue_taps<kPair, 4>(d_tmem, ad, wd, dA, dW, idesc, first, ntap);
              else if (nk == 2) issue_taps<kPair, 2>(d_tmem, ad, wd, dA, dW, idesc, first, ntap);
              else if (nk == 3) issue_taps<kPair, 3>(d_tmem, ad, wd, dA, dW, idesc, first, ntap);
              else issue_taps<kPair, 1>(d_tmem, ad, wd, dA, dW, idesc, first, ntap);
              if (!w_resident) {
                if (kPair) tc_commit_2sm(&w_empty[sw], mc_mask);
                else if (cs == 1) tc_commit(&w_empty[sw]); else tc_commit_mc(&w_empty[sw], mc_mask);
              }
              if (tap0 + ntap >= taps) {   // last W stage of this 64-channel block: the A halo tile is free
                if (kPair) tc_commit_2sm(&a_empty[sa], mc_mask); else tc_commit(&a_empty[sa]);
                if (cb == ncb - 1) {
                  if (kPair) tc_commit_2sm(&tmem_full[acc], mc_mask); else tc_commit(&tmem_full[acc]);
                }
              }
            }
            __syncwarp();
            if (++sw == T2_NW) { sw = 0; pw ^= 1; }
          }
          if (++sa == T2_NA) { sa = 0; pa ^= 1; }
          if (kBlockEpi && p.fuse && n_done > 0) {
            if (cb == 0) issue_conv1(0);
            if (cb == p.c1_after1) issue_conv1(1);
          }
        }
        if (kBlockEpi && p.fuse) ++n_done;
        if (++acc == nacc) { acc = 0; pacc ^= 1; }
      }
      if (kBlockEpi && p.fuse && n_done > 0) { issue_conv1(0); issue_conv1(1); }
    }
  } else if (kBlockEpi && warp >= 4 && p.fuse) {
    // ================= fused residual unit: epilogue 1 (conv7 -> operand tile) and epilogue 2 (conv1 -> stream, operand) =================
    const int ew = warp - 4, quarter = warp & 3, grp = ew >> 2;
    const int nch = p.BN / 32;
    const int slot_rows = p.Tmax * p.rows_per_frame;
    const uint32_t cst_u32 = smem_u32(cst);
    const uint32_t sw128 = (uint32_t)(lane & 7);                         // 128B swizzle: chunk ^= row & 7
    uint8_t* const c_row = c_tile + (size_t)(quarter * 32 + lane) * 128;
    const bool has_a2 = p.out_a2 != nullptr, has_y2 = p.out_y2 != nullptr;
    const int hch = p.BN / 64;                                           // 32-column chunks per N half
    int acc = 0, pv_b = 0, pv_t0 = 0;
    uint32_t pacc = 0, ptile = 0;                                        // ptile: parity of the previous tile's conv1 barriers
    bool have_prev = false, pv_mine = false;
    // Epilogue 2 of the previous tile, N half h: acc2 = conv1(...); + b1 + X -> X' [, snake_next(X')].  The residual chunk
    // (32 rows x 32 columns of X) comes in by TMA into this warp's staging buffer, X' is written over it and leaves by TMA,
    // then the operand reuses the buffer: no scattered 16-byte global accesses (they cost 32 LSU wavefronts per instruction).
    // (A second 2 KB buffer for the operand -- so that it need not wait for the X' store to drain this one -- measured SLOWER: the
    // 24 KB come out of the A / W rings, 957 vs 1041-1107 TFLOP/s on block 2.)
    const uint32_t buf = smem_u32(staging) + (uint32_t)ew * 2048u, my_row = buf + (uint32_t)lane * 64u;
    const uint32_t sw64 = (uint32_t)((lane >> 1) & 3);
    uint32_t xpar = 0;
    auto epilogue2 = [&](int h) {
      const int n = h * (p.BN / 2) + grp * 32, trow0 = pv_t0 + quarter * 32;
      const bool ok = pv_mine && grp < hch;                              // warp-uniform
      if (ok && lane == 0) {
        tma_store_wait_read0();                                          // my previous store has drained the buffer
        mbar_expect_tx(&xbar[ew], 2048u);
        tma_load_3d((void*)(staging + (size_t)ew * 2048), &map_y, &xbar[ew], n, trow0, pv_b);
      }
      mbar_wait(&tmem_full2[h], ptile);
      tc_fence_after();
      if (ok) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p.acc2_col + grp * 32), r);
        mbar_wait(&xbar[ew], xpar);
        xpar ^= 1;
        uint4 rres[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rres[c] = lds128(my_row + ((((uint32_t)c) ^ sw64) << 4));
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                                 // the accumulator is in registers: conv1's next half may run
          if (kPair && rank != 0) mbar_arrive_remote(acc2_empty, 0); else mbar_arrive(acc2_empty);
        }
        const uint32_t sb = cst_u32 + 4u * (uint32_t)(3 * p.N + n), se = sb + 4u * (uint32_t)p.N, si = se + 4u * (uint32_t)p.N;
        epi_res_pass1<T16>(r, sb, rres, my_row, sw64);
        if (has_y2) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_3d(&map_y, staging + (size_t)ew * 2048, n, trow0, pv_b); tma_store_commit(); }
        }
        if (has_a2) {
          epi_res_pass2<T16>(r, se, si);
          if (lane == 0) tma_store_wait_read0();
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c) sts128(my_row + ((((uint32_t)c) ^ sw64) << 4), make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_3d(&map_o, staging + (size_t)ew * 2048, n, trow0, pv_b); tma_store_commit(); }
        }
      } else {
        __syncwarp();
        if (lane == 0) {
          if (kPair && rank != 0) mbar_arrive_remote(acc2_empty, 0); else mbar_arrive(acc2_empty);
        }
      }
    };
    for (int item = cid; item < items; item += ncl) {
      int b, t0, n0; bool mine;
      if (!coords(item, b, t0, n0, mine)) continue;
      if (have_prev) { epilogue2(0); epilogue2(1); ptile ^= 1; }          // also: conv1(prev) no longer reads the operand tile
      // ---- epilogue 1: acc + b7 -> snake -> conv1's operand tile (128B-swizzled, K-major) ----
      mbar_wait(&tmem_full[acc], pacc);
      tc_fence_after();
      if (p.a_tmem) {
        // warp (quarter, grp) owns accumulator columns [64 grp, 64 grp + 64): it reads them as two 32-column chunks and
        // writes the packed 16-bit operand back over the first 32 of them (a chunk is in registers before its columns are reused)
        const uint32_t t0a = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_stride + grp * 64);
        for (int c2 = 0; c2 < 2 && 2 * grp + c2 < nch; ++c2) {
          uint32_t r[32], pk[16];
          tc_ld32(t0a + (uint32_t)(c2 * 32), r);
          tc_wait_ld();
          const int n = (2 * grp + c2) * 32;
          const uint32_t sb = cst_u32 + 4u * (uint32_t)n, se = sb + 4u * (uint32_t)p.N, si = se + 4u * (uint32_t)p.N;
          epi_snake_pack<T16>(r, sb, se, si, true, pk);
          tc_st16(t0a + (uint32_t)(c2 * 16), pk);
        }
        tc_wait_st();
      } else
      for (int ch = grp; ch < nch; ch += 3) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_stride + ch * 32), r);
        tc_wait_ld();
        const int n = ch * 32;
        const uint32_t sb = cst_u32 + 4u * (uint32_t)n, se = sb + 4u * (uint32_t)p.N, si = se + 4u * (uint32_t)p.N;
        const uint4 none[4] = {};
        uint8_t* row_a = c_row + (size_t)(ch >> 1) * 16384;
        epi_block_chunk<T16, false, false, true, true>(r, nullptr, nullptr, nullptr, sb, se, si, none, nullptr, row_a, (uint32_t)((ch & 1) * 4), sw128, true);
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the conv7 accumulator is drained and this warp's part of the operand tile is written
        if (kPair && rank != 0) { mbar_arrive_remote(&tmem_empty[acc], 0); mbar_arrive_remote(c_ready, 0); }
        else { mbar_arrive(&tmem_empty[acc]); mbar_arrive(c_ready); }
      }
      have_prev = true; pv_b = b; pv_t0 = t0; pv_mine = mine;
      if (++acc == p.nacc) { acc = 0; pacc ^= 1; }
    }
    if (have_prev) { epilogue2(0); epilogue2(1); }
    if (lane == 0) tma_store_wait_all();
  } else if (kBlockEpi && warp >= 4) {
    // ================= block epilogue (12 warps) =================
    const int ew = warp - 4, quarter = warp & 3, grp = ew >> 2;          // grp 0..2 takes chunks grp, grp+3, ...
    const int nch = p.BN / 32;
    const bool has_res = p.res != nullptr, has_y = p.out_y != nullptr, has_a = p.out_a != nullptr;
    // This warp's staging: 32 rows x 64 B for a, then for y -- times stg_bufs.  With ONE buffer every chunk began by waiting until the
    // TMA engine had read the previous chunk's stores out of it (queued behind the producer's loads: ~1-2 k cycles), which is why the
    // HBM-bound GEMMs (1x1 convs, transposed convs) sat at 55-60 % of the HBM roofline; two buffers take that wait off the chain.
    const uint32_t stg_one = (has_y && has_a) ? 4096u : 2048u;
    uint8_t* const stg_base = staging + (size_t)ew * stg_one * (uint32_t)p.stg_bufs;
    uint32_t chunk_ctr = 0;
    const int slot_rows = p.Tmax * p.rows_per_frame;
    const uint32_t cst_u32 = smem_u32(cst);
    const uint32_t sw64 = (uint32_t)((lane >> 1) & 3);                    // SWIZZLE_64B staging rows
    int acc = 0;
    uint32_t pacc = 0;
    for (int item = cid; item < items; item += ncl) {
      int b, t0, n0; bool mine;
      if (!coords(item, b, t0, n0, mine)) continue;
      const int t = min(t0 + quarter * 32 + lane, slot_rows - 1);         // clamp: rows past the slot are clipped by the TMA store
      const T16* res_row = (has_res && mine) ? (const T16*)p.res + (long long)b * p.res_bstride + (long long)t * p.ldres + n0 : nullptr;
      uint4 rres[4] = {};
      if (res_row && grp < nch) {   // residual of my first chunk: in flight while the MMAs of this tile finish
#pragma unroll
        for (int c = 0; c < 4; ++c) rres[c] = __ldg((const uint4*)(res_row + grp * 32 + 8 * c));
      }
      mbar_wait(&tmem_full[acc], pacc);
      tc_fence_after();
      if (mine) {
        for (int ch = grp; ch < nch; ch += 3) {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_stride + ch * 32), r);
          if (has_res && ch != grp) {
#pragma unroll
            for (int c = 0; c < 4; ++c) rres[c] = __ldg((const uint4*)(res_row + ch * 32 + 8 * c));
          }
          tc_wait_ld();
          uint8_t* buf_a = stg_base + (p.stg_bufs == 2 ? (chunk_ctr & 1u) * stg_one : 0u);
          uint8_t* buf_y = buf_a + (has_a ? 2048 : 0);
          ++chunk_ctr;
          if (lane == 0) { if (p.stg_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read0(); }   // the stores that last used this buffer have drained it
          __syncwarp();
          const int n = n0 + ch * 32;
          const uint32_t sb = cst_u32 + 4u * (uint32_t)n, se = sb + 4u * (uint32_t)p.N, si = se + 4u * (uint32_t)p.N;
#define Q3_EPI(RES, Y, SM) epi_block_chunk<T16, RES, Y, true, SM>(r, p.bias + n, p.snake_ea + n, p.snake_ib + n, sb, se, si, rres, buf_y + lane * 64, buf_a + lane * 64, 0u, sw64, true)
#define Q3_EPI_Y(SM) epi_block_chunk<T16, false, true, false, SM>(r, p.bias + n, nullptr, nullptr, sb, se, si, rres, buf_y + lane * 64, buf_a + lane * 64, 0u, sw64, true)
#define Q3_EPI_YG(SM) epi_block_chunk<T16, false, true, false, SM, true>(r, p.bias + n, nullptr, nullptr, sb, se, si, rres, buf_y + lane * 64, buf_a + lane * 64, 0u, sw64, true)
          if (!has_a) {   // one output: the stream (the consumer applies its own activation), or a plain / GELU operand
            if (p.act == ACT_GELU) { if (p.cst_staged) Q3_EPI_YG(true); else Q3_EPI_YG(false); }
            else if (p.cst_staged) Q3_EPI_Y(true); else Q3_EPI_Y(false);
          } else if (p.cst_staged) {
            if (has_res) { if (has_y) Q3_EPI(true, true, true); else Q3_EPI(true, false, true); }
            else { if (has_y) Q3_EPI(false, true, true); else Q3_EPI(false, false, true); }
          } else {
            if (has_res) { if (has_y) Q3_EPI(true, true, false); else Q3_EPI(true, false, false); }
            else { if (has_y) Q3_EPI(false, true, false); else Q3_EPI(false, false, false); }
          }
#undef Q3_EPI
#undef Q3_EPI_Y
#undef Q3_EPI_YG
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (has_y) tma_store_3d(&map_y, buf_y, n, t0 + quarter * 32, b);
            if (has_a) tma_store_3d(&map_o, buf_a, n, t0 + quarter * 32, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && rank != 0) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == p.nacc) { acc = 0; pacc ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (!kBlockEpi && warp >= 4) {
    // ================= generic epilogue (8 warps) =================
    // Two groups of four warps (one warp per TMEM lane quarter) split the tile's 16-column chunks.  16-bit outputs
    // are written to a 128B-per-4-rows swizzled staging chunk in smem and leave through ONE TMA store per chunk and
    // tensor (coalesced, asynchronous) instead of 32-way scattered 16-byte global stores per warp instruction.
    const int ew = warp - 4, quarter = warp & 3, chalf = ew >> 2;
    const bool yf32 = y_is_f32 != 0;
    const int nchunks = p.BN / 16;
    const int c_begin = chalf == 0 ? 0 : (nchunks + 1) / 2, c_end = chalf == 0 ? (nchunks + 1) / 2 : nchunks;
    const bool prefetch_res = p.res != nullptr && !yf32;
    const bool use_tma = (p.tma_y | p.tma_a) != 0;
    const bool leader = quarter == 0 && lane == 0;
    const int rin = quarter * 32 + lane;                              // row inside the tile == TMEM lane
    const uint32_t sw = (uint32_t)((rin >> 2) & 1);                   // SWIZZLE_32B: 16-byte chunk index ^= address bit 7
    uint8_t* my_stage = staging + (size_t)chalf * (4 * T2_CHUNK_BYTES);   // [buf][y|a] chunks of this group
    uint32_t ci = 0;                                                  // running chunk counter -> staging buffer parity
    int acc = 0;
    uint32_t pacc = 0;
    for (int item = cid; item < items; item += ncl) {
      int b, t0, n0; bool mine;
      if (!coords(item, b, t0, n0, mine)) continue;
      const int t = t0 + rin;
      const bool row_ok = mine && t < __ldg(p.len_frames + min(b, p.B - 1)) * p.rows_per_frame;
      uint4 pre[16];   // 16-bit residual rows of this thread's chunks, in flight while the MMAs run
      if (prefetch_res && row_ok) {
        const uint4* rp = (const uint4*)((const T16*)p.res + (long long)b * p.res_bstride + (long long)t * p.ldres + n0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c_begin + i < c_end) { pre[2 * i] = rp[2 * (c_begin + i)]; pre[2 * i + 1] = rp[2 * (c_begin + i) + 1]; }
      }
      mbar_wait(&tmem_full[acc], pacc);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = c_begin + i;
        if (ch >= c_end) break;
        uint32_t r[16];
        __syncwarp();
        tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_stride + ch * 16), r);
        tc_wait_ld();
        const int n = n0 + ch * 16;
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
        if (p.bias) {
          float bb[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) ld4(p.bias + n, bb, e);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += bb[e];
        }
        if (p.act == ACT_SWIGLU) {   // (never staged: its output has N/2 columns)
          if (row_ok) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float g0 = v[4 * e], u0 = v[4 * e + 1], g1 = v[4 * e + 2], u1 = v[4 * e + 3];
              o[e] = Cvt<T16>::pack(g0 / (1.0f + __expf(-g0)) * u0, g1 / (1.0f + __expf(-g1)) * u1);
            }
            *(uint4*)((T16*)p.out_a + (long long)b * p.ao_bstride + (long long)t * p.lda_out + (n >> 1)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
          continue;
        }
        if (p.act == ACT_GELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = gelu_exact(v[e]);
        }
        if (p.res) {
          float rr[16];
          if (prefetch_res) {
            if (row_ok) {
              const uint4 u0 = pre[2 * i], u1 = pre[2 * i + 1];
              const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float2 f = Cvt<T16>::unpack(w[e]); rr[2 * e] = f.x; rr[2 * e + 1] = f.y; }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) rr[e] = 0.f;
            }
          } else if (row_ok) {
            load16<T16>(p.res, yf32, (long long)b * p.res_bstride + (long long)t * p.ldres + n, rr);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) rr[e] = 0.f;
          }
          if (p.scale) {
            float sc[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) ld4(p.scale + n, sc, e);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaf(sc[e], v[e], rr[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += rr[e];
          }
        }
        uint8_t* buf = my_stage + (size_t)(ci & 1) * (2 * T2_CHUNK_BYTES);
        if (p.out_y) {
          if (p.tma_y) {
            uint4* d = (uint4*)(buf + rin * 32);
            d[0 ^ sw] = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]), Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
            d[1 ^ sw] = make_uint4(Cvt<T16>::pack(v[8], v[9]), Cvt<T16>::pack(v[10], v[11]), Cvt<T16>::pack(v[12], v[13]), Cvt<T16>::pack(v[14], v[15]));
          } else if (row_ok) {
            store16<T16>(p.out_y, yf32, (long long)b * p.y_bstride + (long long)t * p.ldy + n, v);
          }
        }
        if (p.out_tap && row_ok) store16<T16>(p.out_tap, true, (long long)b * p.tap_bstride + (long long)t * p.ldt + n, v);
        if (p.out_a) {
          if (p.snake_ea) {
            float ea[16], ib[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) { ld4(p.snake_ea + n, ea, e); ld4(p.snake_ib + n, ib, e); }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float sn = __sinf(v[e] * ea[e]);
              v[e] = fmaf(ib[e], sn * sn, v[e]);
            }
          }
          if (p.tma_a) {
            uint4* d = (uint4*)(buf + T2_CHUNK_BYTES + rin * 32);
            d[0 ^ sw] = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]), Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
            d[1 ^ sw] = make_uint4(Cvt<T16>::pack(v[8], v[9]), Cvt<T16>::pack(v[10], v[11]), Cvt<T16>::pack(v[12], v[13]), Cvt<T16>::pack(v[14], v[15]));
          } else if (row_ok) {
            store16<T16>(p.out_a, false, (long long)b * p.ao_bstride + (long long)t * p.lda_out + n, v);
          }
        }
        if (use_tma) {
          fence_async_smem();                              // my st.shared -> visible to the async (TMA) proxy
          if (leader) tma_store_wait_read0();              // the previous chunk's store has drained the OTHER buffer
          named_bar_sync(1 + chalf, 128);
          if (leader && mine) {
            if (p.tma_y) tma_store_3d(&map_y, buf, n, t0, b);
            if (p.tma_a) tma_store_3d(&map_o, buf + T2_CHUNK_BYTES, n, t0, b);
            tma_store_commit();
          }
          ++ci;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // pair mode: the accumulator ring is owned by the leader's MMA warp
        if (kPair && rank != 0) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == p.nacc) { acc = 0; pacc ^= 1; }
    }
    if (use_tma && leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into this CTA or arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn2() {
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
int pick_bn2(int N, bool prefer32 = false) {
  if (prefer32)
    for (int bn = 256; bn >= 64; bn -= 32)
      if (N % bn == 0) return bn;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}
// The block epilogue: bias + [16-bit residual] + [16-bit stream out] + SnakeBeta operand out, nothing else.
bool block_epilogue_ok(const ConvGemmParams& p, int y_dtype) {
  const bool with_a = p.out_a && p.snake_ea && p.act == ACT_NONE, y_only = !p.out_a && !p.snake_ea && p.out_y && !p.res && p.act == ACT_NONE;
  const bool plain_a = p.out_a && !p.snake_ea && !p.out_y && !p.res && (p.act == ACT_NONE || p.act == ACT_GELU);   // one 16-bit operand out
  return (with_a || y_only || plain_a) && p.bias && !p.out_tap && !p.scale &&
         (!p.res || !p.out_y || p.res == p.out_y) && (!p.out_y || y_dtype != DT_F32) && pick_bn2(p.N, true) % 32 == 0 &&
         pick_bn2(p.N, true) >= 64;
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T16>
cudaError_t launch_variant(const cudaLaunchConfig_t& cfg, bool pair, bool block, const CUtensorMap& ma, const CUtensorMap& mw,
                           const CUtensorMap& my, const CUtensorMap& mo, const CUtensorMap& mw2, const Tc2Params& q, int yf) {
  static tc::PerDeviceOnce optin;   // one per T16 instantiation
  const cudaError_t oe = optin.ensure([]() {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<T16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<T16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<T16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_tc2_kernel<T16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e;
  });
  if (oe != cudaSuccess) return oe;
  if (pair) {
    if (block) return cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<T16, true, true>, ma, mw, my, mo, mw2, q, yf);
    return cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<T16, true, false>, ma, mw, my, mo, mw2, q, yf);
  }
  if (block) return cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<T16, false, true>, ma, mw, my, mo, mw2, q, yf);
  return cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<T16, false, false>, ma, mw, my, mo, mw2, q, yf);
}
}  // namespace

// Q3TTS_TC_HALO=0 disables the tensor-core path (CUDA-core 16-bit GEMM everywhere: a debugging aid)
int tc2_mode() {
  static const int m = env_int("Q3TTS_TC_HALO", 1);
  return m;
}

bool tc2_supported(const ConvGemmParams& p, int op_dtype) {
  if (tc2_mode() == 0) return false;
  if (op_dtype != DT_F16 && op_dtype != DT_BF16) return false;
  if (p.Cin % 16 || p.Cin < 64 || p.lda != p.Cin) return false;
  if (p.N % 16 || pick_bn2(p.N) < 32) return false;
  if ((p.taps - 1) * p.dil + T2_BM > 256) return false;          // TMA box rows
  if (p.act == ACT_SWIGLU && (!p.out_a || p.out_y || p.res)) return false;
  return encode_fn2() != nullptr;
}

bool tc2_fuse_supported(const ConvGemmParams& p, int op_dtype) {
  // conv7 + conv1 of a residual unit in one kernel: 0.48 ms vs 0.36 + 0.34 ms for N = 192 (960 k rows).  Q3TTS_TC_FUSE=0 disables.
  static const int on = env_int("Q3TTS_TC_FUSE", 1);
  if (!on || !tc2_supported(p, op_dtype)) return false;
  // one N tile (the 1x1 conv needs every channel of the row), whole 64-channel blocks, operand tile + W1 half + rings in smem
  return p.N == p.Cin && p.N % 64 == 0 && p.N <= 192 && pick_bn2(p.N, true) == p.N && p.bias && p.snake_ea && p.act == ACT_NONE &&
         !p.res && !p.out_y && !p.out_tap && !p.scale;
}

static cudaError_t launch_tc2_impl(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s,
                                   const FusedConv1* fuse, bool allow_pair);

cudaError_t launch_conv_gemm_tc2(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s,
                                 const FusedConv1* fuse) {
  return launch_tc2_impl(p, g, op_dtype, y_dtype, s, fuse, true);
}

static cudaError_t launch_tc2_impl(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s,
                                   const FusedConv1* fuse, bool allow_pair) {
  EncodeTiledFn enc = encode_fn2();
  if (!enc) return cudaErrorNotSupported;
  static const int cs_env = env_int("Q3TTS_TC_CLUSTER", 2);
  static const int pair_env = env_int("Q3TTS_TC_PAIR", 1);
  static const int block_env = env_int("Q3TTS_TC_BLOCK_EPI", 1);
  static const int resident_env = env_int("Q3TTS_TC_W_RESIDENT", 1);
  static const int nacc_env = env_int("Q3TTS_TC_NACC", 4);
  static const int cst_env = env_int("Q3TTS_TC_CST", 1);
  const bool epi_block = fuse != nullptr || (block_env && block_epilogue_ok(p, y_dtype));
  const int BN = pick_bn2(p.N, epi_block);
  const int slot_rows = g.Tmax * p.rows_per_frame;
  const int halo = (p.taps - 1) * p.dil;
  Tc2Params q{};
  q.B = g.B; q.Tmax = g.Tmax; q.rows_per_frame = p.rows_per_frame; q.len_frames = g.len_frames;
  q.N = p.N; q.BN = BN; q.Cin = p.Cin; q.taps = p.taps; q.dil = p.dil; q.ncb = (p.Cin + T2_BK - 1) / T2_BK; q.halo = halo;
  q.tiles_per_utt = (slot_rows + T2_BM - 1) / T2_BM;
  q.n_tiles = p.N / BN;
  q.m_tiles_total = g.B * q.tiles_per_utt;
  int cs = (cs_env == 2 && BN % 16 == 0 && (q.m_tiles_total >= 2 || fuse)) ? 2 : 1;   // the fused unit always runs as a pair
  const bool pair = pair_env && cs == 2 && allow_pair;
  if (fuse && !pair) return cudaErrorNotSupported;
  q.cs = cs;
  q.nacc = (!fuse && BN <= 128 && nacc_env >= 4) ? 4 : 2;   // the fused unit tracks two conv1 accumulators
  q.acc_stride = fuse ? BN : (int)T2_TMEM_COLS / q.nacc;            // fused unit: 2 x BN + BN/2 columns
  const CUtensorMapDataType dt = op_dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap map_a, map_w;
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)slot_rows, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.a_bstride * 2};
    cuuint32_t box[3] = {T2_BK, (cuuint32_t)(T2_BM + halo), 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&map_a, dt, 3, const_cast<void*>(p.A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)p.taps * p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {T2_BK, (cuuint32_t)(BN / cs)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&map_w, dt, 2, const_cast<void*>(p.W), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const uint32_t fmt = op_dtype == DT_F16 ? 0u : 1u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((pair ? 2 * T2_BM : T2_BM) >> 4) << 24);
  q.a_tx_bytes = (uint32_t)(T2_BM + halo) * 128u;
  q.a_stage_bytes = (q.a_tx_bytes + 1023u) & ~1023u;
  q.w_tap_bytes = (uint32_t)(pair ? BN / 2 : BN) * 128u;     // pair mode: each CTA stages half of the weight tile
  {  // several taps share one W stage when the tiles are small: one barrier wait / commit per wg*nk MMAs
    static const int wg_env = env_int("Q3TTS_TC_WG_KB", 40);
    static const int fuse_ts_env = env_int("Q3TTS_FUSE_TS", 1);
    // one wait + commit per W stage costs the issuing thread ~the time of two MMAs: 1-tap stages run N = 192 at 1.09 PFLOP/s,
    // 3-tap stages at 1.36.  The fused unit can afford big stages only when conv1's operand lives in tensor memory.
    const int wg_max = (fuse && !fuse_ts_env) ? 1 : std::max(1, (wg_env * 1024) / (int)q.w_tap_bytes);
    q.ngroups = (p.taps + wg_max - 1) / wg_max;
    q.wg = (p.taps + q.ngroups - 1) / q.ngroups;
    q.ngroups = (p.taps + q.wg - 1) / q.wg;
  }
  q.w_stage_bytes = q.w_tap_bytes * (uint32_t)q.wg;
  q.bias = p.bias; q.act = p.act;
  q.res = p.res; q.ldres = p.ldres; q.res_bstride = p.res_bstride; q.scale = p.scale;
  q.out_y = p.out_y; q.ldy = p.ldy; q.y_bstride = p.y_bstride;
  q.out_a = p.out_a; q.lda_out = p.lda_out; q.ao_bstride = p.ao_bstride;
  const bool plain_a = epi_block && !fuse && p.out_a && !p.snake_ea;   // block epilogue, single plain / GELU operand: it leaves as "y"
  if (plain_a) { q.out_y = p.out_a; q.ldy = p.lda_out; q.y_bstride = p.ao_bstride; q.out_a = nullptr; }
  q.snake_ea = p.snake_ea; q.snake_ib = p.snake_ib;
  q.out_tap = (float*)p.out_tap; q.ldt = p.ldt; q.tap_bstride = p.tap_bstride;
  // 16-bit outputs leave through smem staging + TMA stores
  static const int tma_store_env = env_int("Q3TTS_TC_TMA_STORE", 1);
  const int yf = y_dtype == DT_F32;
  q.tma_y = !fuse && (epi_block || tma_store_env) && q.out_y && (!yf || plain_a) && p.act != ACT_SWIGLU;
  q.tma_a = !fuse && (epi_block || tma_store_env) && q.out_a && p.act != ACT_SWIGLU;
  if (fuse) { q.tma_y = 1; q.tma_a = fuse->out_a != nullptr; }   // X (TMA load of the residual + store of X', in place) and the operand
  q.cst_staged = epi_block && (cst_env || fuse) && p.N <= T2_CST_MAX_N;
  CUtensorMap map_w2 = map_w;
  if (fuse) {
    q.c1_after1 = std::min(1, q.ncb - 1);
    static const int ts_env = env_int("Q3TTS_FUSE_TS", 1);
    q.a_tmem = ts_env != 0;
    q.fuse = 1; q.ncb2 = p.N / T2_BK; q.w1_blk_bytes = (uint32_t)(BN / 4) * 128u; q.acc2_col = 2 * BN;
    q.idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 4) << 17) | ((uint32_t)((2 * T2_BM) >> 4) << 24);
    q.bias2 = fuse->bias; q.ea2 = fuse->ea; q.ib2 = fuse->ib;
    q.res2 = fuse->res; q.out_y2 = fuse->out_y; q.out_a2 = fuse->out_a; q.o2_bstride = (long long)slot_rows * p.N;
    cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.N * 2};
    cuuint32_t box[2] = {T2_BK, (cuuint32_t)(BN / 4)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&map_w2, dt, 2, const_cast<void*>(fuse->W1), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  CUtensorMap map_y = map_a, map_o = map_a;   // placeholders when unused
  auto out_map = [&](CUtensorMap* m, void* base, int ld, long long bstride) -> bool {
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)slot_rows, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
    cuuint32_t box[3] = {16, T2_BM, 1};                       // generic: one 16-column chunk of the whole tile
    if (epi_block) { box[0] = 32; box[1] = 32; }               // block mode: one warp's 32 rows x 32 columns
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               epi_block ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (fuse) {
    if (fuse->out_y && fuse->res != fuse->out_y) return cudaErrorNotSupported;   // the stream is updated in place (or not written)
    if (!out_map(&map_y, const_cast<void*>(fuse->res), p.N, (long long)slot_rows * p.N)) return cudaErrorInvalidValue;
    if (q.tma_a && !out_map(&map_o, fuse->out_a, p.N, (long long)slot_rows * p.N)) return cudaErrorInvalidValue;
  } else {
    if (q.tma_y && !out_map(&map_y, q.out_y, q.ldy, q.y_bstride)) return cudaErrorInvalidValue;
    if (q.tma_a && !out_map(&map_o, q.out_a, q.lda_out, q.ao_bstride)) return cudaErrorInvalidValue;
  }
  static const int stg2_env = env_int("Q3TTS_TC_STG2", 0);   // measured: the second buffer costs ring depth and the transposed convs lose more than the 1x1 convs gain
  q.stg_bufs = (epi_block && !fuse && stg2_env) ? 2 : 1;
  q.staging_bytes = !(q.tma_y || q.tma_a) ? 0u : (epi_block ? (uint32_t)T2_EPI_WARPS_BLOCK * ((q.tma_y && q.tma_a && !fuse) ? 4096u : 2048u) * (uint32_t)q.stg_bufs : T2_STAGING_BYTES);
  const size_t fixed = q.staging_bytes + (fuse ? (size_t)q.ncb2 * ((q.a_tmem ? 0 : 16384) + 2 * q.w1_blk_bytes) : 0) + (q.cst_staged ? (size_t)(fuse ? 6 : 3) * p.N * 4 : 0) + 640 + 1024;
  auto magic = [](uint32_t d, uint32_t* m, uint32_t* sh) {   // x / d == umulhi(x, m) >> sh for 0 <= x < 2^31
    uint32_t lg = 0;
    while ((1u << lg) < d) ++lg;
    const uint64_t mm = ((1ull << (31 + lg)) / d) + 1;
    if (mm >> 32) return false;
    *m = (uint32_t)mm; *sh = lg == 0 ? 0 : lg - 1;   // d == 1 is special-cased in the kernel
    return true;
  };
  if (!magic((uint32_t)q.n_tiles, &q.div_nt_m, &q.div_nt_s) || !magic((uint32_t)q.tiles_per_utt, &q.div_tpu_m, &q.div_tpu_s))
    return cudaErrorInvalidConfiguration;
  const size_t budget = 227 * 1024 - fixed;
  // Resident weights: every (n tile, 64-channel block, tap group) stage has its own slot and is loaded once.
  const int stages_per_tile = q.ncb * q.ngroups;
  const size_t resident_bytes = (size_t)q.n_tiles * stages_per_tile * q.w_stage_bytes;
  const int items_per_cluster = ((q.m_tiles_total + cs - 1) / cs * q.n_tiles) / std::max(1, 148 / cs);
  q.w_resident = resident_env && q.n_tiles * stages_per_tile <= T2_MAX_NW && items_per_cluster >= 4 &&
                 resident_bytes + 64 * 1024 <= budget;   // and still >= 64 KB of A tiles in flight
  if (q.w_resident) {
    q.nw = q.n_tiles * stages_per_tile;
    q.na = (int)std::min<size_t>(T2_MAX_NA, (budget - resident_bytes) / q.a_stage_bytes);
  } else {
    // Little's law: bytes in flight must cover HBM/L2 latency, so the budget is split between the two rings in
    // proportion to what a tile consumes from each (a 1x1 conv streams mostly A, a k=7 conv mostly W).
    const double a_tile = (double)q.ncb * q.a_stage_bytes, w_tile = (double)q.ncb * q.taps * q.w_tap_bytes;
    int na = (int)((double)budget * a_tile / (a_tile + w_tile) / q.a_stage_bytes + 0.5);
    na = std::max(2, std::min(T2_MAX_NA, na));
    while (na > 2 && (size_t)na * q.a_stage_bytes + 3 * (size_t)q.w_stage_bytes > budget) --na;
    q.na = na;
    q.nw = (int)std::min<size_t>(T2_MAX_NW, (budget - (size_t)q.na * q.a_stage_bytes) / q.w_stage_bytes);
    static const int max_nw_env = env_int("Q3TTS_TC_MAX_NW", T2_MAX_NW);   // experiments: cap the W ring depth
    q.nw = std::min(q.nw, std::max(2, max_nw_env));
  }
  // Small-K, HBM-bound GEMMs whose weights do not fit in smem even halved (block 1's 1x1 conv: K = 384, two N tiles) run
  // 10 % faster as independent 128-row CTAs (no cross-CTA commits / arrives per item) than as CTA pairs; measured, gemm_bench.
  if (pair && !fuse && !q.w_resident && p.taps * p.Cin <= 384 && q.n_tiles >= 2)
    return launch_tc2_impl(p, g, op_dtype, y_dtype, s, fuse, false);
  static const int na_env = env_int("Q3TTS_TC_NA", 0);   // experiments
  if (na_env > 0 && !q.w_resident) {
    q.na = std::min(T2_MAX_NA, na_env);
    if ((size_t)q.na * q.a_stage_bytes + 2 * (size_t)q.w_stage_bytes > budget) return cudaErrorInvalidConfiguration;
    q.nw = (int)std::min<size_t>(T2_MAX_NW, (budget - (size_t)q.na * q.a_stage_bytes) / q.w_stage_bytes);
  }
  // Two MMA issuer warps on alternate tiles: a warp then waits on ring slots up to one tile ahead of the other's, and
  // an mbarrier parity wait cannot tell phase n from phase n-2, so both rings must hold more than one tile.
  static const int nmma_env = env_int("Q3TTS_TC_NMMA", 2);
  q.nmma = (!fuse && nmma_env == 2 && q.na >= q.ncb + 1 && (q.w_resident || q.nw >= stages_per_tile + 1)) ? 2 : 1;
  if (q.nw < 2 && !q.w_resident) return cudaErrorInvalidConfiguration;
  if (q.na < 2) return cudaErrorInvalidConfiguration;
  size_t smem = (size_t)q.na * q.a_stage_bytes + (size_t)q.nw * q.w_stage_bytes + fixed;
  smem = std::max<size_t>(smem, 128 * 1024);   // one CTA per SM: it owns all 512 TMEM columns
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = (q.m_tiles_total + cs - 1) / cs;
  int grid = std::min(groups * q.n_tiles * cs, sms / cs * cs);
  grid = std::max(grid / cs * cs, cs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(epi_block ? T2_THREADS_BLOCK : T2_THREADS_GENERIC);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (op_dtype == DT_F16) return launch_variant<__half>(cfg, pair, epi_block, map_a, map_w, map_y, map_o, map_w2, q, yf);
  return launch_variant<__nv_bfloat16>(cfg, pair, epi_block, map_a, map_w, map_y, map_o, map_w2, q, yf);
}

}  // namespace q3
