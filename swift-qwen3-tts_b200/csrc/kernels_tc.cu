// tcgen05 / TMEM / TMA multi-tap GEMM for sm_100a: the tensor-core engine behind every conv, linear and
// transposed conv of the decoder in the 16-bit modes (see kernels.cuh for the multi-tap GEMM definition).
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (smem, K-major, 128B swizzle) * W[BN x 64]^T (smem, K-major)
//
// * A tile for tap j = TMA box {64 ch, 128 rows, 1 utterance} of the channels-last activation tensor at row
//   t0 - (taps-1-j)*dil: the causal left padding (and any row past the slot) is TMA out-of-bounds zero fill,
//   so there is no im2col buffer and no pad copy.  Cin that is not a multiple of 64 is zero-filled the same way.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane, tcgen05.mma cta_group::1 kind::f16),
//   warp 2 = TMEM allocator, warps 4..11 = epilogue (tcgen05.ld 32x32b, two warps per TMEM lane quarter).
// * smem ring of 4 stages (mbarrier full/empty), TMEM ring of 2 accumulators (tmem_full/tmem_empty): the fused
//   epilogue (bias, GELU / SwiGLU, layer-scale + residual, SnakeBeta for the NEXT conv's operand) of tile i
//   overlaps the MMAs of tile i+1.  Persistent CTAs, one per SM, static round-robin over (utterance, M, N) tiles.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace q3 {

namespace {

constexpr int TC_BM = 128;        // rows (time) per tile == UMMA M
constexpr int TC_BK = 64;         // K per stage == one 128-byte swizzle atom of 16-bit elements
constexpr int TC_STAGES = 4;
constexpr int TC_MAX_BN = 256;
constexpr int TC_THREADS = 384;   // 4 control warps + 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t TC_TMEM_COLS = 512;

struct TcParams {
  int B, Tmax, rows_per_frame;
  const int* len_frames;
  int N, BN, taps, dil, ncb;      // ncb = ceil(Cin/64)
  int tiles_per_utt, n_tiles;     // M tiles per utterance slot, N tiles
  uint32_t idesc;
  uint32_t b_bytes;               // BN * 128
  // epilogue (same meaning as ConvGemmParams)
  const float* bias; int act;
  const void* res; int ldres; long long res_bstride;
  const float* scale;
  void* out_y; int ldy; long long y_bstride;
  void* out_a; int lda_out; long long ao_bstride;
  const float* snake_ea; const float* snake_ib;
  float* out_tap; int ldt; long long tap_bstride;
};

using namespace tc;

// ---- the kernel ---------------------------------------------------------------------------------------------
template <typename T16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, TcParams p,
                    int y_is_f32) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-B alignment
  const uint32_t stage_bytes = TC_A_BYTES + p.b_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)TC_STAGES * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_STAGES;
  uint64_t* tmem_full = bars + 2 * TC_STAGES;
  uint64_t* tmem_empty = bars + 2 * TC_STAGES + 2;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.tiles_per_utt * p.n_tiles;
  const int nkb = p.taps * p.ncb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto tile_coords = [&](int tile, int& b, int& t0, int& n0) -> bool {
    const int nt = tile % p.n_tiles, mg = tile / p.n_tiles;
    b = mg / p.tiles_per_utt;
    t0 = (mg % p.tiles_per_utt) * TC_BM;
    n0 = nt * p.BN;
    return t0 < p.len_frames[b] * p.rows_per_frame;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b, t0, n0;
        if (!tile_coords(tile, b, t0, n0)) continue;
        for (int kb = 0; kb < nkb; ++kb) {
          const int tap = kb / p.ncb, c0 = (kb % p.ncb) * TC_BK;
          const int shift = (p.taps - 1 - tap) * p.dil;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_3d(sa, &map_a, &full_bar[stage], c0, t0 - shift, b);
          tma_load_2d(sa + TC_A_BYTES, &map_w, &full_bar[stage], c0, tap * p.N + n0);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b, t0, n0;
        if (!tile_coords(tile, b, t0, n0)) continue;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_MAX_BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)   // +32 B along K inside the swizzle atom == +2 in the address field
            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (kb | k) != 0);
          tc_commit(&empty_bar[stage]);          // frees the smem slot once these MMAs have read it
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = warp - 4, quarter = warp & 3, chalf = ew >> 2;   // TMEM lanes [32*quarter, +32); column half
    const bool yf32 = y_is_f32 != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int b, t0, n0;
      if (!tile_coords(tile, b, t0, n0)) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int t = t0 + quarter * 32 + lane;
      const bool row_ok = t < p.len_frames[b] * p.rows_per_frame;
      const int nchunks = p.BN / 16;
      const int c_begin = chalf == 0 ? 0 : (nchunks + 1) / 2, c_end = chalf == 0 ? (nchunks + 1) / 2 : nchunks;
      for (int ch = c_begin; ch < c_end; ++ch) {
        uint32_t r[16];
        __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the row_ok-predicated body below
        tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_MAX_BN + ch * 16), r);
        tc_wait_ld();
        if (!row_ok) continue;
        const int n = n0 + ch * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + n + i);
        }
        if (p.act == ACT_SWIGLU) {   // columns (2i, 2i+1) = (gate, up)  (ST.swift:560-562)
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float g0 = v[4 * i], u0 = v[4 * i + 1], g1 = v[4 * i + 2], u1 = v[4 * i + 3];
            o[i] = Cvt<T16>::pack(g0 / (1.0f + __expf(-g0)) * u0, g1 / (1.0f + __expf(-g1)) * u1);
          }
          *(uint4*)((T16*)p.out_a + (long long)b * p.ao_bstride + (long long)t * p.lda_out + (n >> 1)) = make_uint4(o[0], o[1], o[2], o[3]);
          continue;
        }
        if (p.act == ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu_exact(v[i]);
        }
        if (p.res) {
          float rr[16];
          load16<T16>(p.res, yf32, (long long)b * p.res_bstride + (long long)t * p.ldres + n, rr);
          if (p.scale) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(__ldg(p.scale + n + i), v[i], rr[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rr[i];
          }
        }
        if (p.out_y) store16<T16>(p.out_y, yf32, (long long)b * p.y_bstride + (long long)t * p.ldy + n, v);
        if (p.out_tap) store16<T16>(p.out_tap, true, (long long)b * p.tap_bstride + (long long)t * p.ldt + n, v);
        if (p.out_a) {
          if (p.snake_ea) {   // SnakeBeta of the consumer (ST.swift:246-253), applied once here instead of once per tap there
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s = __sinf(v[i] * __ldg(p.snake_ea + n + i));
              v[i] = fmaf(__ldg(p.snake_ib + n + i), s * s, v[i]);
            }
          }
          store16<T16>(p.out_a, false, (long long)b * p.ao_bstride + (long long)t * p.lda_out + n, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int pick_bn(int N) {
  for (int bn = TC_MAX_BN; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace

bool tc_supported(const ConvGemmParams& p, int op_dtype) {
  if (op_dtype != DT_F16 && op_dtype != DT_BF16) return false;
  static const bool disabled = [] { const char* e = getenv("Q3TTS_NO_TC"); return e && e[0] == '1'; }();
  if (disabled) return false;   // A/B switch: route 16-bit GEMMs through the CUDA-core kernel
  if (p.Cin % 8 || p.Cin < 64) return false;               // TMA: 16-byte global strides; at least one full K block
  if (p.N % 16 || pick_bn(p.N) < 32) return false;
  if (p.act == ACT_SWIGLU && (!p.out_a || p.out_y || p.res)) return false;
  if (p.lda != p.Cin) return false;
  return encode_fn() != nullptr;
}

cudaError_t launch_conv_gemm_tc(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return cudaErrorNotSupported;
  const int BN = pick_bn(p.N);
  const int slot_rows = g.Tmax * p.rows_per_frame;
  const CUtensorMapDataType dt = op_dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap map_a, map_w;
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)slot_rows, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)p.lda * 2, (cuuint64_t)p.a_bstride * 2};
    cuuint32_t box[3] = {TC_BK, TC_BM, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&map_a, dt, 3, const_cast<void*>(p.A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)p.taps * p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    if (enc(&map_w, dt, 2, const_cast<void*>(p.W), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  TcParams q{};
  q.B = g.B; q.Tmax = g.Tmax; q.rows_per_frame = p.rows_per_frame; q.len_frames = g.len_frames;
  q.N = p.N; q.BN = BN; q.taps = p.taps; q.dil = p.dil; q.ncb = (p.Cin + TC_BK - 1) / TC_BK;
  q.tiles_per_utt = (slot_rows + TC_BM - 1) / TC_BM;
  q.n_tiles = p.N / BN;
  // instruction descriptor (kind::f16): D = F32, A/B = F16 or BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  const uint32_t fmt = op_dtype == DT_F16 ? 0u : 1u;
  q.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  q.b_bytes = (uint32_t)BN * 128u;
  q.bias = p.bias; q.act = p.act;
  q.res = p.res; q.ldres = p.ldres; q.res_bstride = p.res_bstride; q.scale = p.scale;
  q.out_y = p.out_y; q.ldy = p.ldy; q.y_bstride = p.y_bstride;
  q.out_a = p.out_a; q.lda_out = p.lda_out; q.ao_bstride = p.ao_bstride;
  q.snake_ea = p.snake_ea; q.snake_ib = p.snake_ib;
  q.out_tap = (float*)p.out_tap; q.ldt = p.ldt; q.tap_bstride = p.tap_bstride;
  // >= 128 KB so that exactly one CTA (which allocates all 512 TMEM columns) is resident per SM
  const size_t smem = std::max<size_t>((size_t)TC_STAGES * (TC_A_BYTES + q.b_bytes) + 256 + 1024, 128 * 1024);
  const int total_tiles = g.B * q.tiles_per_utt * q.n_tiles;
  const int grid = std::min(total_tiles, num_sms());
  cudaError_t err;
  if (op_dtype == DT_F16) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_gemm_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
    conv_gemm_tc_kernel<__half><<<grid, TC_THREADS, smem, s>>>(map_a, map_w, q, y_dtype == DT_F32);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_gemm_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = true; }
    conv_gemm_tc_kernel<__nv_bfloat16><<<grid, TC_THREADS, smem, s>>>(map_a, map_w, q, y_dtype == DT_F32);
  }
  err = cudaGetLastError();
  return err;
}

}  // namespace q3
