// fp32 reference-mode multi-tap GEMM (CUDA cores) and the row kernels shared by both precision
// modes (RVQ gather-sum, RMSNorm, depthwise-conv+LayerNorm, attention, tail conv, lengths, taps).
// Semantics follow /root/reference/Sources/Qwen3TTS/Models/SpeechTokenizer.swift (ST.swift);
// each kernel cites the lines it implements.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.cuh"

namespace q3 {

static thread_local int g_pdl_suspended = 0;
void pdl_suspend(bool on) { g_pdl_suspended += on ? 1 : -1; }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("Q3TTS_PDL"); return !(e && e[0] == '0'); }();
  return on && g_pdl_suspended == 0;
}


// ---- typed element access ---------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __half* p, int64_t i) { return __half2float(p[i]); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(__half* p, int64_t i, float v) { p[i] = __float2half_rn(v); }
__device__ __forceinline__ void stf(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

#define Q3_DISPATCH_DT(dt, T, ...)                                   \
  do {                                                               \
    if ((dt) == DT_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dt) == DT_F16) { using T = __half; __VA_ARGS__; }      \
    else { using T = __nv_bfloat16; __VA_ARGS__; }                   \
  } while (0)

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float snake_f(float v, float ea, float ib) {
  float s = sinf(v * ea);   // ST.swift:246-253
  return v + ib * (s * s);
}

// ================================================================================================
// CUDA-core multi-tap GEMM (fp32 accumulate)
// ================================================================================================
constexpr int F_BM = 128, F_BN = 64, F_BK = 16, F_TM = 8, F_TN = 4, F_PAD = 4;

__device__ __forceinline__ float4 load4(const float* p) { return *(const float4*)p; }
__device__ __forceinline__ float4 load4(const __half* p) {
  const uint2 u = *(const uint2*)p;
  const float2 a = __half22float2(*(const __half2*)&u.x), b = __half22float2(*(const __half2*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *(const uint2*)p;
  const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&u.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// TA: operand type (A, W, out_a); TS: stream type (res, out_y).  fp32 accumulate on CUDA cores.
// In fp32 mode this IS the engine; in 16-bit mode it is the fallback for shapes the tcgen05 kernel
// does not cover.
template <typename TA, typename TS>
__global__ void __launch_bounds__(256)
conv_gemm_simt_kernel(ConvGemmParams p, BatchGeom g, int tiles_per_utt) {
  __shared__ __align__(16) float As[F_BK][F_BM + F_PAD];
  __shared__ __align__(16) float Bs[F_BK][F_BN + F_PAD];
  const int b = blockIdx.x / tiles_per_utt;
  const int t0 = (blockIdx.x % tiles_per_utt) * F_BM;
  const int n0 = blockIdx.y * F_BN;
  const int slot_rows = g.Tmax * p.rows_per_frame;
  const int valid_rows = g.len_frames[b] * p.rows_per_frame;
  if (t0 >= valid_rows) return;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const TA* A = (const TA*)p.A + (int64_t)b * p.a_bstride;
  const TA* W = (const TA*)p.W;
  const int ncb = (p.Cin + F_BK - 1) / F_BK;
  const int n_it = p.taps * ncb;

  float acc[F_TM][F_TN];
#pragma unroll
  for (int i = 0; i < F_TM; ++i)
#pragma unroll
    for (int j = 0; j < F_TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;
  auto gload = [&](int it) {
    const int j = it / ncb, c0 = (it % ncb) * F_BK;
    const int shift = (p.taps - 1 - j) * p.dil;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, row = idx >> 2, c = c0 + (idx & 3) * 4;
      const int tin = t0 + row - shift;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tin >= 0 && tin < slot_rows && c < p.Cin) v = load4(A + (int64_t)tin * p.lda + c);
      ra[i] = v;
    }
    {
      const int row = tid >> 2, c = c0 + (tid & 3) * 4, n = n0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.N && c < p.Cin) v = load4(W + ((int64_t)j * p.N + n) * p.Cin + c);
      rb = v;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, row = idx >> 2, k = (idx & 3) * 4;
      As[k + 0][row] = ra[i].x; As[k + 1][row] = ra[i].y; As[k + 2][row] = ra[i].z; As[k + 3][row] = ra[i].w;
    }
    const int row = tid >> 2, k = (tid & 3) * 4;
    Bs[k + 0][row] = rb.x; Bs[k + 1][row] = rb.y; Bs[k + 2][row] = rb.z; Bs[k + 3][row] = rb.w;
  };

  gload(0);
  sstore();
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (it + 1 < n_it) gload(it + 1);
#pragma unroll
    for (int k = 0; k < F_BK; ++k) {
      float a[F_TM], bb[F_TN];
      *(float4*)&a[0] = *(const float4*)&As[k][ty * F_TM];
      *(float4*)&a[4] = *(const float4*)&As[k][ty * F_TM + 4];
      *(float4*)&bb[0] = *(const float4*)&Bs[k][tx * F_TN];
#pragma unroll
      for (int i = 0; i < F_TM; ++i)
#pragma unroll
        for (int j = 0; j < F_TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
    if (it + 1 < n_it) {
      sstore();
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const int ncol0 = n0 + tx * F_TN;
#pragma unroll
  for (int i = 0; i < F_TM; ++i) {
    const int t = t0 + ty * F_TM + i;
    if (t >= valid_rows) continue;
    if (p.act == ACT_SWIGLU) {
#pragma unroll
      for (int j = 0; j < F_TN; j += 2) {
        const int n = ncol0 + j;
        if (n >= p.N) continue;
        float gte = acc[i][j], up = acc[i][j + 1];
        if (p.bias) { gte += p.bias[n]; up += p.bias[n + 1]; }
        const float v = silu_f(gte) * up;   // ST.swift:560-562
        const int oc = n >> 1;
        if (p.out_a) stf((TA*)p.out_a, (int64_t)b * p.ao_bstride + (int64_t)t * p.lda_out + oc, v);
        if (p.out_y) stf((TS*)p.out_y, (int64_t)b * p.y_bstride + (int64_t)t * p.ldy + oc, v);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < F_TN; ++j) {
      const int n = ncol0 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.act == ACT_GELU) v = gelu_exact(v);
      else if (p.act == ACT_GELU_TANH) v = v * 0.5f * (1.0f + tanhf(0.7978845608f * (v + 0.044715f * (v * v * v))));
      if (p.res) {
        const float r = ldf((const TS*)p.res, (int64_t)b * p.res_bstride + (int64_t)t * p.ldres + n);
        v = r + (p.scale ? p.scale[n] * v : v);
      }
      if (p.out_y) stf((TS*)p.out_y, (int64_t)b * p.y_bstride + (int64_t)t * p.ldy + n, v);
      if (p.out_tap) ((float*)p.out_tap)[(int64_t)b * p.tap_bstride + (int64_t)t * p.ldt + n] = v;
      if (p.out_a) {
        const float a = p.snake_ea ? snake_f(v, p.snake_ea[n], p.snake_ib[n]) : (p.a_elu ? (v > 0.f ? v : expf(v) - 1.0f) : v);
        stf((TA*)p.out_a, (int64_t)b * p.ao_bstride + (int64_t)t * p.lda_out + n, a);
      }
    }
  }
}

void launch_conv_gemm_simt(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s) {
  const int slot_rows = g.Tmax * p.rows_per_frame;
  const int tiles = (slot_rows + F_BM - 1) / F_BM;
  dim3 grid((unsigned)(g.B * tiles), (unsigned)((p.N + F_BN - 1) / F_BN));
  if (op_dtype == DT_F32) conv_gemm_simt_kernel<float, float><<<grid, 256, 0, s>>>(p, g, tiles);
  else if (op_dtype == DT_F16) {
    if (y_dtype == DT_F32) conv_gemm_simt_kernel<__half, float><<<grid, 256, 0, s>>>(p, g, tiles);
    else conv_gemm_simt_kernel<__half, __half><<<grid, 256, 0, s>>>(p, g, tiles);
  } else {
    if (y_dtype == DT_F32) conv_gemm_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(p, g, tiles);
    else conv_gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(p, g, tiles);
  }
}

// ================================================================================================
// RVQ gather-and-sum (ST.swift:28-30, 50-55, 81-96): one warp per frame, 128-bit row reads.
// ================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) rvq_kernel(RvqParams p, BatchGeom g) {
  pdl_trigger();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = g.B * g.Tmax;
  if (warp >= rows) return;
  const int b = warp / g.Tmax, t = warp % g.Tmax;
  if (t >= g.len_frames[b]) return;
  const int32_t* cb = p.codes + p.code_base[b] + (int64_t)t * p.st;
  TO* out = (TO*)p.out + (int64_t)warp * (2 * p.half);
  for (int part = 0; part < 2; ++part) {
    const int q0 = part == 0 ? 0 : p.num_sem, q1 = part == 0 ? p.num_sem : p.num_q;
    for (int d0 = lane * 4; d0 < p.half; d0 += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = q0; q < q1; ++q) {
        int code = cb[(int64_t)q * p.sq];
        if (code < 0 || code >= p.table_sizes[q]) {
          if (d0 == lane * 4) atomicOr(p.err_flag, 1);
          code = 0;
        }
        const float4 e = __ldg((const float4*)(p.tables[q] + (int64_t)code * p.half + d0));
        if (q == q0) acc = e;   // first term is taken as is, then sequential left-to-right adds
        else { acc.x = __fadd_rn(acc.x, e.x); acc.y = __fadd_rn(acc.y, e.y); acc.z = __fadd_rn(acc.z, e.z); acc.w = __fadd_rn(acc.w, e.w); }
      }
      const int o = part * p.half + d0;
      stf(out, o + 0, acc.x); stf(out, o + 1, acc.y); stf(out, o + 2, acc.z); stf(out, o + 3, acc.w);
    }
  }
}

void launch_rvq(const RvqParams& p, const BatchGeom& g, cudaStream_t s) {
  const int64_t rows = (int64_t)g.B * g.Tmax;
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  Q3_DISPATCH_DT(p.out_dtype, T, (rvq_kernel<T><<<blocks, 256, 0, s>>>(p, g)));
}

// ================================================================================================
// RMSNorm (x * rsqrt(mean(x^2)+eps) * w), one warp per row.  ST.swift:589, 595, 639.
// ================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* x, const float* w, float eps, TO* out, int64_t rows, int C) {
  pdl_trigger();
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) { float v = xr[c]; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / (float)C + eps);
  for (int c = lane; c < C; c += 32) stf(out, row * C + c, xr[c] * r * w[c]);
}

void launch_rmsnorm(const float* x, const float* w, float eps, void* out, int out_dtype, int64_t rows, int C, cudaStream_t s) {
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  Q3_DISPATCH_DT(out_dtype, T, (rmsnorm_kernel<T><<<blocks, 256, 0, s>>>(x, w, eps, (T*)out, rows, C)));
}

// ================================================================================================
// Depthwise causal conv k=7 + bias, then LayerNorm over C (ST.swift:389-393, 372-379).
// One CTA per output row; each thread owns channels tid, tid+256, ...
// ================================================================================================
// One WARP per output row (no block barriers): a lane owns channels 4*lane + 128*k .. +3 (k < C/128), so every access is a
// 16-byte load / store that the warp coalesces into 512 contiguous bytes; the 8 warps of a CTA take 8 consecutive rows, whose
// 7-row windows overlap in L1.  w7 is [7][C] (tap-major).  C % 128 == 0, C <= 2048; other widths use the generic kernel below.
__device__ __forceinline__ void st4(float* p, float4 v) { *(float4*)p = v; }
__device__ __forceinline__ void st4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  *(uint2*)p = make_uint2(*(uint32_t*)&a, *(uint32_t*)&b);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *(uint2*)p = make_uint2(*(uint32_t*)&a, *(uint32_t*)&b);
}
template <typename TO, int KC>
__global__ void __launch_bounds__(256)
dwconv_ln_warp_kernel(const float* __restrict__ x, const float* __restrict__ w7, const float* __restrict__ wb,
                      const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, TO* __restrict__ out, BatchGeom g,
                      int rows_per_frame) {
  pdl_trigger();
  constexpr int C = KC * 128;
  const int slot_rows = g.Tmax * rows_per_frame;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (int64_t)g.B * slot_rows) return;
  const int b = (int)(row / slot_rows), t = (int)(row % slot_rows), lane = threadIdx.x & 31;
  if (t >= g.len_frames[b] * rows_per_frame) return;
  const float* xb = x + (int64_t)b * slot_rows * C;
  float4 acc[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) acc[k] = __ldg((const float4*)(wb + 4 * lane + 128 * k));
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int tin = t - 6 + j;
    if (tin < 0) continue;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 xv = __ldg((const float4*)(xb + (int64_t)tin * C + 4 * lane + 128 * k));
      const float4 wv = __ldg((const float4*)(w7 + j * C + 4 * lane + 128 * k));
      acc[k].x = fmaf(wv.x, xv.x, acc[k].x); acc[k].y = fmaf(wv.y, xv.y, acc[k].y);
      acc[k].z = fmaf(wv.z, xv.z, acc[k].z); acc[k].w = fmaf(wv.w, xv.w, acc[k].w);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KC; ++k) sum += (acc[k].x + acc[k].y) + (acc[k].z + acc[k].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const float dx = acc[k].x - mean, dy = acc[k].y - mean, dz = acc[k].z - mean, dw = acc[k].w - mean;
    sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float r = rsqrtf(sq / (float)C + eps);
  TO* orow = out + ((int64_t)b * slot_rows + t) * C;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const float4 lw = __ldg((const float4*)(ln_w + 4 * lane + 128 * k)), lb = __ldg((const float4*)(ln_b + 4 * lane + 128 * k));
    st4(orow + 4 * lane + 128 * k, make_float4((acc[k].x - mean) * r * lw.x + lb.x, (acc[k].y - mean) * r * lw.y + lb.y,
                                               (acc[k].z - mean) * r * lw.z + lb.z, (acc[k].w - mean) * r * lw.w + lb.w));
  }
}

template <typename TO>
__global__ void __launch_bounds__(256)
dwconv_ln_kernel(const float* x, const float* w7, const float* wb, const float* ln_w, const float* ln_b, float eps,
                 TO* out, BatchGeom g, int rows_per_frame, int C) {
  const int slot_rows = g.Tmax * rows_per_frame;
  const int b = blockIdx.x / slot_rows, t = blockIdx.x % slot_rows;
  if (t >= g.len_frames[b] * rows_per_frame) return;
  const float* xb = x + (int64_t)b * slot_rows * C;
  float vals[8];
  float sum = 0.f;
  int cnt = 0;
  for (int c = threadIdx.x; c < C; c += 256, ++cnt) {
    float a = wb[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int tin = t - 6 + j;
      if (tin >= 0) a = fmaf(w7[j * C + c], xb[(int64_t)tin * C + c], a);
    }
    vals[cnt] = a;
    sum += a;
  }
  __shared__ float red[8];
  auto block_sum = [&](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    __syncthreads();
    return tot;
  };
  const float mean = block_sum(sum) / (float)C;
  float sq = 0.f;
  for (int i = 0; i < cnt; ++i) { float d = vals[i] - mean; sq = fmaf(d, d, sq); }
  const float var = block_sum(sq) / (float)C;
  const float r = rsqrtf(var + eps);
  TO* orow = out + ((int64_t)b * slot_rows + t) * C;
  cnt = 0;
  for (int c = threadIdx.x; c < C; c += 256, ++cnt) stf(orow, c, (vals[cnt] - mean) * r * ln_w[c] + ln_b[c]);
}

void launch_dwconv_ln(const float* x, const float* w7, const float* wb, const float* ln_w, const float* ln_b, float eps,
                      void* out, int out_dtype, const BatchGeom& g, int rows_per_frame, int C, cudaStream_t s) {
  const int64_t rows = (int64_t)g.B * g.Tmax * rows_per_frame;
  if (C == 1024) {
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    Q3_DISPATCH_DT(out_dtype, T, (dwconv_ln_warp_kernel<T, 8><<<blocks, 256, 0, s>>>(x, w7, wb, ln_w, ln_b, eps, (T*)out, g, rows_per_frame)));
    return;
  }
  if (C == 512) {
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    Q3_DISPATCH_DT(out_dtype, T, (dwconv_ln_warp_kernel<T, 4><<<blocks, 256, 0, s>>>(x, w7, wb, ln_w, ln_b, eps, (T*)out, g, rows_per_frame)));
    return;
  }
  Q3_DISPATCH_DT(out_dtype, T, (dwconv_ln_kernel<T><<<(unsigned)rows, 256, 0, s>>>(x, w7, wb, ln_w, ln_b, eps, (T*)out, g, rows_per_frame, C)));
}

// ================================================================================================
// Attention (ST.swift:512-528): softmax(scale * Q K^T [+ causal sliding-window mask]) V per (utterance, head).
// One warp per query row, keys streamed through shared memory in chunks of 32 with an online softmax.
// head_dim <= 128 and a multiple of 32.  Keys are limited to the utterance's own frames, so padded
// slots never leak into valid frames.
// ================================================================================================
constexpr int ATT_QB = 16;   // queries per CTA (4 warps x 4 queries)
template <typename TI, typename TO, int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const TI* qkv, TO* out, BatchGeom g, int nh, int nkv, float scale, int window) {
  constexpr int DPL = HD / 32;   // dims per lane
  __shared__ float Ks[32][HD + 1];
  __shared__ float Vs[32][HD + 1];
  __shared__ float Qs[ATT_QB][HD];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_QB;
  const int len = g.len_frames[b], beg = g.row_begin ? g.row_begin[b] : 0;
  if (q0 >= len) return;
  const int hk = h / (nh / nkv);
  const int ld = (nh + 2 * nkv) * HD;
  const TI* base = qkv + (int64_t)b * g.Tmax * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < ATT_QB * HD; i += 128) {
    const int qi = i / HD, d = i % HD;
    Qs[qi][d] = (q0 + qi < len) ? ldf(base, (int64_t)(q0 + qi) * ld + h * HD + d) * scale : 0.f;
  }
  float m[4], l[4], o[4][DPL];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[i][d] = 0.f;
  }
  const int qlast = min(q0 + ATT_QB, len) - 1;
  int k_begin = 0, k_end = len;
  if (window > 0) { k_begin = max(0, q0 - window + 1); k_end = qlast + 1; }
  k_begin = max(k_begin, beg) & ~31;
  for (int kc = k_begin; kc < k_end; kc += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * HD; i += 128) {
      const int kj = i / HD, d = i % HD;
      const bool ok = kc + kj < len && kc + kj >= beg;
      Ks[kj][d] = ok ? ldf(base, (int64_t)(kc + kj) * ld + (nh + hk) * HD + d) : 0.f;
      Vs[kj][d] = ok ? ldf(base, (int64_t)(kc + kj) * ld + (nh + nkv + hk) * HD + d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = warp * 4 + i, tq = q0 + qi;
      if (tq >= len) continue;
      const int kj = kc + lane;
      float sc = 0.f;
#pragma unroll 16
      for (int d = 0; d < HD; ++d) sc = fmaf(Qs[qi][d], Ks[lane][d], sc);
      bool ok = kj < len && kj >= beg;
      if (window > 0) ok = ok && kj <= tq && (tq - kj) < window;
      sc = ok ? sc : -INFINITY;
      float cm = sc;
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, of));
      const float mn = fmaxf(m[i], cm);
      if (mn == -INFINITY) continue;   // whole chunk masked for this query
      const float corr = __expf(m[i] - mn);
      const float pj = ok ? __expf(sc - mn) : 0.f;
      float ps = pj;
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, of);
      l[i] = l[i] * corr + ps;
      m[i] = mn;
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[i][d] *= corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pjj = __shfl_sync(0xffffffffu, pj, j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[i][d] = fmaf(pjj, Vs[j][lane + 32 * d], o[i][d]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tq = q0 + warp * 4 + i;
    if (tq >= len) continue;
    const float inv = 1.0f / l[i];
    TO* orow = out + ((int64_t)b * g.Tmax + tq) * (nh * HD) + h * HD;
#pragma unroll
    for (int d = 0; d < DPL; ++d) stf(orow, lane + 32 * d, o[i][d] * inv);
  }
}

template <typename TI, typename TO>
static void attention_dispatch(const void* qkv, void* out, const BatchGeom& g, int nh, int nkv, int hd, float scale,
                               int window, cudaStream_t s) {
  dim3 grid((unsigned)((g.Tmax + ATT_QB - 1) / ATT_QB), (unsigned)nh, (unsigned)g.B);
  if (hd == 64) attention_kernel<TI, TO, 64><<<grid, 128, 0, s>>>((const TI*)qkv, (TO*)out, g, nh, nkv, scale, window);
  else if (hd == 32) attention_kernel<TI, TO, 32><<<grid, 128, 0, s>>>((const TI*)qkv, (TO*)out, g, nh, nkv, scale, window);
  else if (hd == 128) attention_kernel<TI, TO, 128><<<grid, 128, 0, s>>>((const TI*)qkv, (TO*)out, g, nh, nkv, scale, window);
}

void launch_attention(const void* qkv, int qkv_dtype, void* out, int out_dtype, const BatchGeom& g, int nh, int nkv,
                      int hd, float scale, int causal_window, cudaStream_t s) {
  // fp32 softmax accumulation in every mode (MLX's fused SDPA does the same)
  static const bool no_mma = [] { const char* e = getenv("Q3TTS_NO_MMA_ATTN"); return e && e[0] == '1'; }();
  if (!no_mma && attention_mma_supported(qkv_dtype, hd)) {
    launch_attention_mma(qkv, qkv_dtype, out, g, nh, nkv, scale, causal_window, s);
    return;
  }
  if (qkv_dtype == DT_F32) attention_dispatch<float, float>(qkv, out, g, nh, nkv, hd, scale, causal_window, s);
  else if (qkv_dtype == DT_F16) attention_dispatch<__half, __half>(qkv, out, g, nh, nkv, hd, scale, causal_window, s);
  else attention_dispatch<__nv_bfloat16, __nv_bfloat16>(qkv, out, g, nh, nkv, hd, scale, causal_window, s);
  (void)out_dtype;
}

// ================================================================================================
// Tail: outConv (C -> 1, k=7, causal) + bias + clip (ST.swift:674-678, 688, 781).  One warp per sample.
// ================================================================================================
template <typename TI>
__global__ void __launch_bounds__(256)
tail_kernel(const TI* a, int64_t a_bstride, const float* w, float bias, int C, float* pcm, const int64_t* pcm_base,
            float* tap, int64_t tap_bstride, BatchGeom g, int rows_per_frame) {
  const int64_t slot_rows = (int64_t)g.Tmax * rows_per_frame;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= slot_rows * g.B) return;
  const int b = (int)(gw / slot_rows);
  const int64_t t = gw % slot_rows;
  if (t >= (int64_t)g.len_frames[b] * rows_per_frame) return;
  const TI* ab = a + (int64_t)b * a_bstride;
  float acc = 0.f;
  for (int j = 0; j < 7; ++j) {
    const int64_t tin = t - 6 + j;
    if (tin < 0) continue;
    for (int c = lane; c < C; c += 32) acc = fmaf(w[j * C + c], ldf(ab, tin * C + c), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float v = acc + bias;
    if (tap) tap[(int64_t)b * tap_bstride + t] = v;
    store_pcm(pcm, pcm_base[b] + t, v, g.pcm_i16);
  }
}

// 16-bit operand, C in {72, 96}: one thread per output sample, the 7-row x C window staged once per CTA in shared
// memory with 128-bit loads (row stride C+8 halves = conflict-free LDS.128), weights in shared memory (broadcast reads).
constexpr int TAIL_TILE = 192;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], __half) {
  const float2 a = __half22float2(*(const __half2*)&u.x), b = __half22float2(*(const __half2*)&u.y),
               c = __half22float2(*(const __half2*)&u.z), d = __half22float2(*(const __half2*)&u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], __nv_bfloat16) {
  const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&u.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&u.y),
               c = __bfloat1622float2(*(const __nv_bfloat162*)&u.z), d = __bfloat1622float2(*(const __nv_bfloat162*)&u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <typename T16, int C>
__global__ void __launch_bounds__(TAIL_TILE)
tail16_kernel(const T16* a, int64_t a_bstride, const float* __restrict__ w /*[7][C], the model's own buffer*/, float bias, float* pcm,
              const int64_t* pcm_base, float* tap, int64_t tap_bstride, BatchGeom g, int rows_per_frame, int tiles_per_utt) {
  constexpr int STRIDE = C + 8, ROWS = TAIL_TILE + 6, V = C / 8;
  __shared__ __align__(16) T16 tile[ROWS * STRIDE];
  __shared__ float ws[7 * C];   // per-CTA copy of this model's weights (a process-wide __constant__ would be shared by all models)
  const int b = blockIdx.x / tiles_per_utt;
  const int64_t t0 = (int64_t)(blockIdx.x % tiles_per_utt) * TAIL_TILE;
  const int64_t slot_rows = (int64_t)g.Tmax * rows_per_frame, valid = (int64_t)g.len_frames[b] * rows_per_frame;
  if (t0 >= valid) return;
  const T16* ab = a + (int64_t)b * a_bstride;
  for (int i = threadIdx.x; i < 7 * C; i += TAIL_TILE) ws[i] = __ldg(w + i);
  for (int idx = threadIdx.x; idx < ROWS * V; idx += TAIL_TILE) {
    const int row = idx / V, c8 = idx % V;
    const int64_t tin = t0 - 6 + row;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tin >= 0 && tin < slot_rows) v = __ldg((const uint4*)(ab + tin * C + c8 * 8));
    *(uint4*)&tile[row * STRIDE + c8 * 8] = v;
  }
  __syncthreads();
  const int64_t t = t0 + threadIdx.x;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
#pragma unroll
    for (int c8 = 0; c8 < V; ++c8) {
      float f[8];
      unpack8(*(const uint4*)&tile[(threadIdx.x + j) * STRIDE + c8 * 8], f, T16());
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(ws[j * C + c8 * 8 + e], f[e], acc);
    }
  }
  if (t < valid) {
    const float v = acc + bias;
    if (tap) tap[(int64_t)b * tap_bstride + t] = v;
    store_pcm(pcm, pcm_base[b] + t, v, g.pcm_i16);
  }
}

void launch_tail(const void* a, int a_dtype, int64_t a_bstride, const float* w, float bias, int C, float* pcm,
                 const int64_t* pcm_base, float* tap, int64_t tap_bstride, const BatchGeom& g, int rows_per_frame,
                 cudaStream_t s) {
  static const bool no_mma_tail = [] { const char* e = getenv("Q3TTS_NO_MMA_TAIL"); return e && e[0] == '1'; }();
  if (!no_mma_tail && tail_mma_supported(a_dtype, C)) {
    launch_tail_mma(a, a_dtype, a_bstride, w, bias, C, pcm, pcm_base, tap, tap_bstride, g, rows_per_frame, s);
    return;
  }
  if (a_dtype != DT_F32 && (C == 96 || C == 72)) {
    const int tiles = (int)(((int64_t)g.Tmax * rows_per_frame + TAIL_TILE - 1) / TAIL_TILE);
    const unsigned blocks = (unsigned)(g.B * tiles);
#define Q3_TAIL(T, CC) tail16_kernel<T, CC><<<blocks, TAIL_TILE, 0, s>>>((const T*)a, a_bstride, w, bias, pcm, pcm_base, tap, tap_bstride, g, rows_per_frame, tiles)
    if (a_dtype == DT_F16) { if (C == 96) Q3_TAIL(__half, 96); else Q3_TAIL(__half, 72); }
    else { if (C == 96) Q3_TAIL(__nv_bfloat16, 96); else Q3_TAIL(__nv_bfloat16, 72); }
#undef Q3_TAIL
    return;
  }
  const int64_t warps = (int64_t)g.B * g.Tmax * rows_per_frame;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  Q3_DISPATCH_DT(a_dtype, T, (tail_kernel<T><<<blocks, 256, 0, s>>>((const T*)a, a_bstride, w, bias, C, pcm, pcm_base, tap, tap_bstride, g, rows_per_frame)));
}

// ================================================================================================
// Codec-embedding sum: one CTA per frame, a thread owns 8 consecutive channels (16-byte row reads for the
// 16-bit tables); 1 + 15 table rows are read once each, the running sum stays in registers in the tables' dtype.
// ================================================================================================
template <typename T> struct Emb8;
template <> struct Emb8<float> {
  float v[8];
  __device__ void load(const float* p) { const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)p + 1); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
  __device__ void add(const Emb8& o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], o.v[i]);
  }
  __device__ void store(float* p) const { *(float4*)p = make_float4(v[0], v[1], v[2], v[3]); *((float4*)p + 1) = make_float4(v[4], v[5], v[6], v[7]); }
};
template <> struct Emb8<__half> {
  __half2 v[4];
  __device__ void load(const __half* p) { const uint4 u = __ldg((const uint4*)p); memcpy(v, &u, 16); }
  __device__ void add(const Emb8& o) {      // fp32 add, one rounding to half per add (= MLX's 16-bit add)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __half22float2(v[i]), b = __half22float2(o.v[i]);
      v[i] = __floats2half2_rn(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
    }
  }
  __device__ void store(__half* p) const { uint4 u; memcpy(&u, v, 16); *(uint4*)p = u; }
};
template <> struct Emb8<__nv_bfloat16> {
  __nv_bfloat162 v[4];
  __device__ void load(const __nv_bfloat16* p) { const uint4 u = __ldg((const uint4*)p); memcpy(v, &u, 16); }
  __device__ void add(const Emb8& o) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __bfloat1622float2(v[i]), b = __bfloat1622float2(o.v[i]);
      v[i] = __floats2bfloat162_rn(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
    }
  }
  __device__ void store(__nv_bfloat16* p) const { uint4 u; memcpy(&u, v, 16); *(uint4*)p = u; }
};

template <typename T>
__global__ void __launch_bounds__(256) codec_embed_sum_kernel(CodecEmbedParams p) {
  __shared__ int s_code[kMaxCodeGroups];
  for (long long t = blockIdx.x; t < p.n; t += gridDim.x) {
    if (threadIdx.x < p.groups) {
      int c = p.codes[t * p.groups + threadIdx.x];
      if (c < 0 || c >= p.vocab[threadIdx.x]) { atomicOr(p.err_flag, 1); c = 0; }
      s_code[threadIdx.x] = c;
    }
    __syncthreads();
    for (int d0 = threadIdx.x * 8; d0 < p.H; d0 += blockDim.x * 8) {
      Emb8<T> row[4], acc;
      acc.load((const T*)p.tables[0] + (long long)s_code[0] * p.H + d0);
      for (int g0 = 1; g0 < p.groups; g0 += 4) {       // four independent row reads in flight, adds in order
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (g0 + j < p.groups) row[j].load((const T*)p.tables[g0 + j] + (long long)s_code[g0 + j] * p.H + d0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (g0 + j < p.groups) acc.add(row[j]);
      }
      acc.store((T*)p.out + t * p.H + d0);
    }
    __syncthreads();
  }
}

void launch_codec_embed_sum(const CodecEmbedParams& p, cudaStream_t s) {
  if (p.n <= 0) return;
  const unsigned blocks = (unsigned)std::min<long long>(p.n, 148LL * 8 * 16);
  if (p.dtype == DT_F32) codec_embed_sum_kernel<float><<<blocks, 256, 0, s>>>(p);
  else if (p.dtype == DT_F16) codec_embed_sum_kernel<__half><<<blocks, 256, 0, s>>>(p);
  else codec_embed_sum_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(p);
}

// ================================================================================================
// Batched block copy: one CTA per item, 16-byte chunks (the streaming API's state shuffles).
// ================================================================================================
__global__ void __launch_bounds__(256) block_copy_kernel(const CopyItem* items) {
  pdl_trigger();
  const CopyItem it = items[blockIdx.x];
  const long long n = it.bytes >> 4;
  uint4* d = (uint4*)it.dst;
  const uint4* sp = (const uint4*)it.src;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) d[i] = sp ? sp[i] : make_uint4(0, 0, 0, 0);
}
void launch_block_copy(const CopyItem* d_items, int n_items, cudaStream_t s) {
  if (n_items > 0) block_copy_kernel<<<(unsigned)n_items, 256, 0, s>>>(d_items);
}

// ================================================================================================
// audioLengths (ST.swift:831-833): count of NON-ZERO first-codebook entries, not a prefix length.
// ================================================================================================
__global__ void lengths_kernel(const int32_t* codes, const int64_t* code_base, int64_t st, const int* len_frames,
                               int rate, int32_t* out) {
  const int b = blockIdx.x;
  int cnt = 0;
  for (int t = threadIdx.x; t < len_frames[b]; t += blockDim.x) cnt += codes[code_base[b] + (int64_t)t * st] > 0;
  __shared__ int red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    out[b] = tot * rate;
  }
}

void launch_lengths(const int32_t* codes, const int64_t* code_base, int64_t st, const int* len_frames, int B, int rate,
                    int32_t* out, cudaStream_t s) {
  lengths_kernel<<<B, 256, 0, s>>>(codes, code_base, st, len_frames, rate, out);
}

// ================================================================================================
// Stage tap: channels-last [B, rows, C] -> fp32 NCT [B, C, L] (the reference's inter-module layout).
// ================================================================================================
template <typename TI>
__global__ void tap_copy_kernel(const TI* src, int64_t bstride, int ld, float* dst, int C, int64_t L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t t = t0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < L && c < C) ? ldf(src, (int64_t)b * bstride + t * ld + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t t = t0 + threadIdx.x;
    if (t < L && c < C) dst[((int64_t)b * C + c) * L + t] = tile[threadIdx.x][i];
  }
}

void launch_tap_copy(const void* src, int dtype, int64_t bstride, int ld, float* dst, int B, int C, int64_t L, cudaStream_t s) {
  dim3 grid((unsigned)((L + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B), block(32, 8);
  Q3_DISPATCH_DT(dtype, T, (tap_copy_kernel<T><<<grid, block, 0, s>>>((const T*)src, bstride, ld, dst, C, L)));
}

template <typename TO>
__global__ void convert_kernel(const float* src, TO* dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) stf(dst, i, src[i]);
}
void launch_convert(const float* src, void* dst, int dtype, int64_t n, cudaStream_t s) {
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  Q3_DISPATCH_DT(dtype, T, (convert_kernel<T><<<blocks, 256, 0, s>>>(src, (T*)dst, n)));
}

}  // namespace q3
