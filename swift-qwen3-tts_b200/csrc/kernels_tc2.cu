// tcgen05 multi-tap GEMM, second generation: HALO tiles + cluster-multicast weights.
//
// v1 (kernels_tc.cu) re-reads the activation tile once per tap and the weight tile once per 128-row tile; ncu showed
// block3's conv7 moving 18 GB L2->SM for 1.1 GB of input (profiles/r1_v1_*).  Here:
//  * the A operand of ALL taps of a 64-channel block comes from ONE TMA box of 128 + (taps-1)*dil rows (the halo
//    tile).  Tap j is the same smem tile viewed from row j*dil: its UMMA descriptor simply starts j*dil*128 bytes
//    later (the 128B swizzle is a function of the absolute smem address, which TMA and the MMA unit share);
//  * the CTAs of a cluster work on consecutive M tiles of the same N tile; each loads 1/cs of every weight tile and
//    TMA-multicasts it to the whole cluster, so L2 reads of W drop by the cluster size;
//  * partial K blocks (Cin % 64 != 0, e.g. 96) issue only the k-steps that hold data instead of multiplying zeros;
//  * the epilogue prefetches its 16-bit residual rows while the MMAs of the tile are still running.
// Roles, rings and the fused epilogue are otherwise those of v1.
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

constexpr int T2_BM = 128, T2_BK = 64, T2_MAX_NA = 8, T2_MAX_NW = 12, T2_MAX_BN = 256;
constexpr uint32_t T2_CHUNK_BYTES = 128 * 32;            // one staged output chunk: 128 rows x 16 columns x 2 B
constexpr uint32_t T2_STAGING_BYTES = 48 * 1024;   // generic: 2 groups x 2 outputs x 2 buffers x 4 KB; block mode: 12 warps x 2 outputs x 2 KB
constexpr int T2_THREADS = 512, T2_EPI_WARPS = 8, T2_EPI_WARPS_BLOCK = 12;
constexpr uint32_t T2_TMEM_COLS = 512;

struct Tc2Params {
  int B, Tmax, rows_per_frame;
  const int* len_frames;
  int N, BN, Cin, taps, dil, ncb, halo;          // halo = (taps-1)*dil rows in front of every tile
  int tiles_per_utt, n_tiles, m_tiles_total, cs; // cs = cluster size (1 or 2)
  uint32_t idesc, a_stage_bytes, a_tx_bytes, w_stage_bytes;
  int desc_mode;                                 // 1: base_offset 0, 2: base_offset = (addr >> 7) & 7
  int na, nw;                                    // ring depths (A halo tiles, W stages)
  int wg;                                        // taps per W stage: one barrier wait / commit per wg*nk MMAs
  uint32_t w_tap_bytes;                          // bytes of one tap's weight tile in this CTA (BN/cs rows x 128 B)
  int tma_y, tma_a;                              // 16-bit outputs leave through smem staging + TMA store
  int epi_block;                                 // 0: generic epilogue (8 warps); 1: specialised block epilogue (12 warps)
  const float* bias; int act;
  const void* res; int ldres; long long res_bstride;
  const float* scale;
  void* out_y; int ldy; long long y_bstride;
  void* out_a; int lda_out; long long ao_bstride;
  const float* snake_ea; const float* snake_ib;
  float* out_tap; int ldt; long long tap_bstride;
};

__device__ __forceinline__ void ld4(const float* p, float (&v)[16], int i) {
  const float4 t = __ldg((const float4*)p + i);
  v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
}


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

// Specialised epilogue of the decoder blocks' GEMMs (transposed conv, conv7, conv1 -- ~95 % of the step's elements):
//   v = acc + bias [+ res16];   y = v (16-bit, kY);   a = v + ib * sin^2(v * ea) (16-bit)
// 12 warps (3 per TMEM lane quarter) take the tile's 32-column chunks round-robin.  Each WARP stages its own 32 rows
// (64-byte rows, 64B swizzle) and issues its own TMA stores: no cross-warp barrier, no scattered global stores, and no
// per-element branches (everything the generic epilogue decides at run time is a template parameter here).
template <typename T16, bool kRes, bool kY>
__device__ __forceinline__ void epi_block_chunk(const uint32_t (&r)[32], const float* __restrict__ bias, const float* __restrict__ ea,
                                                const float* __restrict__ ib, const uint4 (&rres)[4], uint8_t* buf_y,
                                                uint8_t* buf_a, int lane) {
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);   // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // 8 columns per step
    float v[8];
    const float4 b0 = __ldg((const float4*)(bias + 8 * c)), b1 = __ldg((const float4*)(bias + 8 * c + 4));
    v[0] = __uint_as_float(r[8 * c + 0]) + b0.x; v[1] = __uint_as_float(r[8 * c + 1]) + b0.y;
    v[2] = __uint_as_float(r[8 * c + 2]) + b0.z; v[3] = __uint_as_float(r[8 * c + 3]) + b0.w;
    v[4] = __uint_as_float(r[8 * c + 4]) + b1.x; v[5] = __uint_as_float(r[8 * c + 5]) + b1.y;
    v[6] = __uint_as_float(r[8 * c + 6]) + b1.z; v[7] = __uint_as_float(r[8 * c + 7]) + b1.w;
    if (kRes) {
      const uint4 u = rres[c];
      const float2 r0 = Cvt<T16>::unpack(u.x), r1 = Cvt<T16>::unpack(u.y), r2 = Cvt<T16>::unpack(u.z), r3 = Cvt<T16>::unpack(u.w);
      v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y; v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
    }
    if (kY)
      *(uint4*)(buf_y + lane * 64 + ((c ^ sw) << 4)) = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]),
                                                                  Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
    const float4 e0 = __ldg((const float4*)(ea + 8 * c)), e1 = __ldg((const float4*)(ea + 8 * c + 4));
    const float4 i0 = __ldg((const float4*)(ib + 8 * c)), i1 = __ldg((const float4*)(ib + 8 * c + 4));
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w}, ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sn = __sinf(v[e] * ee[e]);
      v[e] = fmaf(ii[e], sn * sn, v[e]);
    }
    *(uint4*)(buf_a + lane * 64 + ((c ^ sw) << 4)) = make_uint4(Cvt<T16>::pack(v[0], v[1]), Cvt<T16>::pack(v[2], v[3]),
                                                                Cvt<T16>::pack(v[4], v[5]), Cvt<T16>::pack(v[6], v[7]));
  }
}

// kPair: the two CTAs of the cluster form a cta_group::2 pair -- ONE tcgen05.mma spans both SMs (M = 256), each CTA
// stages its own 128(+halo) rows of A and HALF of the weight tile, so weight ingest and B-operand smem reads per SM
// halve.  Only the leader (rank 0) issues MMAs; both CTAs run producers and epilogues (each on its own TMEM half).
template <typename T16, bool kPair>
__global__ void __launch_bounds__(T2_THREADS, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_o, Tc2Params p,
                     int y_is_f32) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  const int T2_NA = p.na, T2_NW = p.nw;
  uint8_t* w_ring = smem + (size_t)T2_NA * p.a_stage_bytes;
  uint8_t* staging = w_ring + (size_t)T2_NW * p.w_stage_bytes;          // 1024-aligned (all stage sizes are)
  uint64_t* bars = (uint64_t*)(staging + ((p.tma_y | p.tma_a) ? T2_STAGING_BYTES : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + T2_MAX_NA;
  uint64_t* w_full = a_empty + T2_MAX_NA;
  uint64_t* w_empty = w_full + T2_MAX_NW;
  uint64_t* tmem_full = w_empty + T2_MAX_NW;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cs = p.cs;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1);
  const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
  const int groups = (p.m_tiles_total + cs - 1) / cs;
  const int items = groups * p.n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (p.tma_y) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    if (p.tma_a) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < T2_NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < T2_NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], kPair ? 1u : (uint32_t)cs); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], (uint32_t)((kPair ? 2 : 1) * (p.epi_block ? T2_EPI_WARPS_BLOCK : T2_EPI_WARPS))); }
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
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive / multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // item -> (n tile, this CTA's M tile); valid_any = some CTA of the cluster has rows to produce
  auto coords = [&](int item, int& b, int& t0, int& n0, bool& mine) -> bool {
    const int nt = item % p.n_tiles, g = item / p.n_tiles;
    n0 = nt * p.BN;
    bool any = false;
    mine = false;
    b = p.B; t0 = 0;                 // out-of-range utterance: TMA zero-fills, epilogue stores nothing
    for (int r = 0; r < cs; ++r) {
      const int mg = g * cs + r;
      if (mg >= p.m_tiles_total) continue;
      const int bb = mg / p.tiles_per_utt, tt = (mg % p.tiles_per_utt) * T2_BM;
      const bool ok = tt < p.len_frames[bb] * p.rows_per_frame;
      any |= ok;
      if (r == (int)rank) { b = bb; t0 = tt; mine = ok; }
    }
    return any;
  };

  if (warp == 0) {
    // ================= TMA producer (lane 0) + residual L2 prefetch (all lanes) =================
    // The producer runs one to two tiles ahead of the epilogue, so pulling this item's residual rows into L2 here
    // turns the epilogue's residual reads from HBM-latency loads into L2 hits.
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0;
    const int wrows = p.BN / cs;
    const int res_es = y_is_f32 ? 4 : 2;
    for (int item = cid; item < items; item += ncl) {
      int b, t0, n0; bool mine;
      if (!coords(item, b, t0, n0, mine)) continue;
      if (p.res != nullptr && mine) {
        const int valid = p.len_frames[b] * p.rows_per_frame;
        const int line_cnt = (p.BN * res_es + 127) / 128;
        for (int r = lane; r < T2_BM; r += 32) {
          if (t0 + r >= valid) break;
          const char* ptr = (const char*)p.res + ((long long)b * p.res_bstride + (long long)(t0 + r) * p.ldres + n0) * res_es;
          for (int l = 0; l < line_cnt; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + l * 128));
        }
      }
      if (lane == 0) {
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (kPair) {   // both CTAs' boxes complete on the LEADER's barrier
            if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * p.a_tx_bytes);
            tma_load_3d_2sm(a_ring + (size_t)sa * p.a_stage_bytes, &map_a, &a_full[sa], cb * T2_BK, t0 - p.halo, b);
          } else {
            mbar_expect_tx(&a_full[sa], p.a_tx_bytes);
            tma_load_3d(a_ring + (size_t)sa * p.a_stage_bytes, &map_a, &a_full[sa], cb * T2_BK, t0 - p.halo, b);
          }
          if (++sa == T2_NA) { sa = 0; pa ^= 1; }
          for (int tap0 = 0; tap0 < p.taps; tap0 += p.wg) {
            const int ntap = min(p.wg, p.taps - tap0);
            mbar_wait(&w_empty[sw], pw ^ 1);            // every CTA of the cluster has drained this slot
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
            if (++sw == T2_NW) { sw = 0; pw ^= 1; }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && (!kPair || rank == 0)) {
      int sa = 0, sw = 0, acc = 0;
      uint32_t pa = 0, pw = 0, pacc = 0;
      for (int item = cid; item < items; item += ncl) {
        int b, t0, n0; bool mine;
        if (!coords(item, b, t0, n0, mine)) continue;
        mbar_wait(&tmem_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T2_MAX_BN);
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_ring + (size_t)sa * p.a_stage_bytes);
          const int nk = min(T2_BK, p.Cin - cb * T2_BK) / 16;
          for (int tap0 = 0; tap0 < p.taps; tap0 += p.wg) {
            const int ntap = min(p.wg, p.taps - tap0);
            mbar_wait(&w_full[sw], pw);
            tc_fence_after();
            const uint32_t w_base = smem_u32(w_ring + (size_t)sw * p.w_stage_bytes);
            for (int j = 0; j < ntap; ++j) {
              const uint32_t a_tap = a_base + (uint32_t)((tap0 + j) * p.dil) * 128u;   // row-shifted view of the halo tile
              const uint32_t w_tap = w_base + (uint32_t)j * p.w_tap_bytes;
              for (int k = 0; k < nk; ++k) {
                if (kPair) tc_mma_f16_2sm(d_tmem, make_smem_desc_shifted(a_tap + 32u * k, false), make_smem_desc(w_tap + 32u * k), p.idesc, accumulate);
                else tc_mma_f16(d_tmem, make_smem_desc_shifted(a_tap + 32u * k, p.desc_mode == 2), make_smem_desc(w_tap + 32u * k), p.idesc, accumulate);
                accumulate = 1;
              }
            }
            if (kPair) tc_commit_2sm(&w_empty[sw], mc_mask);
            else if (cs == 1) tc_commit(&w_empty[sw]); else tc_commit_mc(&w_empty[sw], mc_mask);
            if (++sw == T2_NW) { sw = 0; pw ^= 1; }
          }
          if (kPair) tc_commit_2sm(&a_empty[sa], mc_mask); else tc_commit(&a_empty[sa]);
          if (++sa == T2_NA) { sa = 0; pa ^= 1; }
        }
        if (kPair) tc_commit_2sm(&tmem_full[acc], mc_mask); else tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
    }
  } else if (warp >= 4 && p.epi_block) {
    // ================= specialised block epilogue (12 warps) =================
    const int ew = warp - 4, quarter = warp & 3, grp = ew >> 2;          // grp 0..2 takes chunks grp, grp+3, ...
    const int nch = p.BN / 32;
    uint8_t* buf_y = staging + (size_t)ew * 4096;                         // this warp's 32 rows x 64 B, y then a
    uint8_t* buf_a = buf_y + 2048;
    const bool has_res = p.res != nullptr, has_y = p.out_y != nullptr;
    const int slot_rows = p.Tmax * p.rows_per_frame;
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
          tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * T2_MAX_BN + ch * 32), r);
          if (has_res && ch != grp) {
#pragma unroll
            for (int c = 0; c < 4; ++c) rres[c] = __ldg((const uint4*)(res_row + ch * 32 + 8 * c));
          }
          tc_wait_ld();
          if (lane == 0) tma_store_wait_read0();                          // my previous stores have drained the staging rows
          __syncwarp();
          const int n = n0 + ch * 32;
          if (has_res) {
            if (has_y) epi_block_chunk<T16, true, true>(r, p.bias + n, p.snake_ea + n, p.snake_ib + n, rres, buf_y, buf_a, lane);
            else epi_block_chunk<T16, true, false>(r, p.bias + n, p.snake_ea + n, p.snake_ib + n, rres, buf_y, buf_a, lane);
          } else {
            if (has_y) epi_block_chunk<T16, false, true>(r, p.bias + n, p.snake_ea + n, p.snake_ib + n, rres, buf_y, buf_a, lane);
            else epi_block_chunk<T16, false, false>(r, p.bias + n, p.snake_ea + n, p.snake_ib + n, rres, buf_y, buf_a, lane);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (has_y) tma_store_3d(&map_y, buf_y, n, t0 + quarter * 32, b);
            tma_store_3d(&map_o, buf_a, n, t0 + quarter * 32, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= 4 && warp < 12) {
    // ================= epilogue =================
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
      const bool row_ok = mine && t < p.len_frames[min(b, p.B - 1)] * p.rows_per_frame;
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
        tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * T2_MAX_BN + ch * 16), r);
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
        if (kPair && rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
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
    for (int bn = T2_MAX_BN; bn >= 64; bn -= 32)
      if (N % bn == 0) return bn;
  for (int bn = T2_MAX_BN; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}
// The specialised block epilogue: bias + [16-bit residual] + [16-bit stream out] + SnakeBeta operand out, nothing else.
bool block_epilogue_ok(const ConvGemmParams& p, int y_dtype) {
  return p.out_a && p.snake_ea && p.bias && p.act == ACT_NONE && !p.out_tap && !p.scale &&
         (!p.res || (p.out_y && p.res == p.out_y)) && (!p.out_y || y_dtype != DT_F32) && pick_bn2(p.N, true) % 32 == 0 &&
         pick_bn2(p.N, true) >= 64;
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
}  // namespace

// 0 = v1 (per-tap loads), 1 = v2 with base_offset 0 (default), 2 = v2 with base_offset from the start address
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

cudaError_t launch_conv_gemm_tc2(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn2();
  if (!enc) return cudaErrorNotSupported;
  static const int cs_env = env_int("Q3TTS_TC_CLUSTER", 2);
  static const int pair_env = env_int("Q3TTS_TC_PAIR", 1);
  static const int block_env = env_int("Q3TTS_TC_BLOCK_EPI", 1);
  const bool epi_block = block_env && block_epilogue_ok(p, y_dtype);
  const int BN = pick_bn2(p.N, epi_block);
  const int slot_rows = g.Tmax * p.rows_per_frame;
  const int halo = (p.taps - 1) * p.dil;
  Tc2Params q{};
  q.epi_block = epi_block;
  q.B = g.B; q.Tmax = g.Tmax; q.rows_per_frame = p.rows_per_frame; q.len_frames = g.len_frames;
  q.N = p.N; q.BN = BN; q.Cin = p.Cin; q.taps = p.taps; q.dil = p.dil; q.ncb = (p.Cin + T2_BK - 1) / T2_BK; q.halo = halo;
  q.tiles_per_utt = (slot_rows + T2_BM - 1) / T2_BM;
  q.n_tiles = p.N / BN;
  q.m_tiles_total = g.B * q.tiles_per_utt;
  int cs = (cs_env == 2 && BN % 16 == 0 && q.m_tiles_total >= 2) ? 2 : 1;
  const bool pair = pair_env && cs == 2;
  q.cs = cs;
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
  {  // several taps share one W stage when the tiles are small: the single MMA thread then waits / commits once per
     // wg*nk MMAs instead of once per nk (its per-stage overhead was what kept the tensor pipe at 30 % for N = 96)
    static const int wg_env = env_int("Q3TTS_TC_WG_KB", 40);
    const int wg_max = std::max(1, (wg_env * 1024) / (int)q.w_tap_bytes);
    const int ngroups = (p.taps + wg_max - 1) / wg_max;
    q.wg = (p.taps + ngroups - 1) / ngroups;
  }
  q.w_stage_bytes = q.w_tap_bytes * (uint32_t)q.wg;
  q.desc_mode = tc2_mode();
  q.bias = p.bias; q.act = p.act;
  q.res = p.res; q.ldres = p.ldres; q.res_bstride = p.res_bstride; q.scale = p.scale;
  q.out_y = p.out_y; q.ldy = p.ldy; q.y_bstride = p.y_bstride;
  q.out_a = p.out_a; q.lda_out = p.lda_out; q.ao_bstride = p.ao_bstride;
  q.snake_ea = p.snake_ea; q.snake_ib = p.snake_ib;
  q.out_tap = (float*)p.out_tap; q.ldt = p.ldt; q.tap_bstride = p.tap_bstride;
  // 16-bit outputs leave through smem staging + TMA stores ({16 cols, 128 rows} boxes, 32-byte swizzle)
  static const int tma_store_env = env_int("Q3TTS_TC_TMA_STORE", 1);
  const int yf = y_dtype == DT_F32;
  q.tma_y = tma_store_env && p.out_y && !yf && p.act != ACT_SWIGLU;
  q.tma_a = tma_store_env && p.out_a && p.act != ACT_SWIGLU;
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
  if (q.tma_y && !out_map(&map_y, p.out_y, p.ldy, p.y_bstride)) return cudaErrorInvalidValue;
  if (q.tma_a && !out_map(&map_o, p.out_a, p.lda_out, p.ao_bstride)) return cudaErrorInvalidValue;
  // ring depths from the smem budget: keep as many weight tiles in flight as fit (small-N convs need many)
  const size_t budget = 225 * 1024 - 2048 - ((q.tma_y || q.tma_a) ? T2_STAGING_BYTES : 0);
  // Little's law: bytes in flight must cover HBM/L2 latency, so the budget is split between the two rings in
  // proportion to what a tile consumes from each (a 1x1 conv streams mostly A, a k=7 conv mostly W).
  {
    const double a_tile = (double)q.ncb * q.a_stage_bytes, w_tile = (double)q.ncb * q.taps * q.w_tap_bytes;
    int na = (int)((double)budget * a_tile / (a_tile + w_tile) / q.a_stage_bytes + 0.5);
    na = std::max(2, std::min(T2_MAX_NA, na));
    while (na > 2 && (size_t)na * q.a_stage_bytes + 3 * (size_t)q.w_stage_bytes > budget) --na;
    q.na = na;
    q.nw = (int)std::min<size_t>(T2_MAX_NW, (budget - (size_t)q.na * q.a_stage_bytes) / q.w_stage_bytes);
  }
  if (q.nw < 2) return cudaErrorInvalidConfiguration;
  size_t smem = (size_t)q.na * q.a_stage_bytes + (size_t)q.nw * q.w_stage_bytes + ((q.tma_y || q.tma_a) ? T2_STAGING_BYTES : 0) + 512 + 1024;
  smem = std::max<size_t>(smem, 128 * 1024);   // one CTA per SM: it owns all 512 TMEM columns
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = (q.m_tiles_total + cs - 1) / cs;
  int grid = std::min(groups * q.n_tiles * cs, sms / cs * cs);
  grid = std::max(grid / cs * cs, cs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err;
  if (op_dtype == DT_F16) {
    static bool done = false;
    if (!done) {
      cudaFuncSetAttribute(conv_gemm_tc2_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(conv_gemm_tc2_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      done = true;
    }
    err = pair ? cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<__half, true>, map_a, map_w, map_y, map_o, q, yf)
               : cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<__half, false>, map_a, map_w, map_y, map_o, q, yf);
  } else {
    static bool done = false;
    if (!done) {
      cudaFuncSetAttribute(conv_gemm_tc2_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(conv_gemm_tc2_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      done = true;
    }
    err = pair ? cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<__nv_bfloat16, true>, map_a, map_w, map_y, map_o, q, yf)
               : cudaLaunchKernelEx(&cfg, conv_gemm_tc2_kernel<__nv_bfloat16, false>, map_a, map_w, map_y, map_o, q, yf);
  }
  return err;
}

}  // namespace q3
