// Warp-level mma.sync kernels for the two small ops that are not worth a tcgen05 pipeline: attention and the tail conv.
//
// Tensor-core attention for the decoder's pre-transformer (ST.swift:512-528): softmax(scale * Q K^T [+ mask]) V per
// (utterance, head), head_dim 64, 16-bit operands, fp32 softmax statistics and output accumulators.
// Attention is 0.1-0.25 % of the decoder's FLOPs and T <= a few thousand, so this is a compact
// flash-style kernel on the warp-level mma.sync path (HMMA m16n8k16), not a tcgen05 pipeline: 4 warps x 16 query rows
// per CTA, 64-key K/V tiles staged in padded shared memory, ldmatrix(.trans) fragment loads, online softmax.
// Keys are limited to the utterance's own frames, so padded batch slots never leak into valid frames.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"

namespace q3 {
namespace {

constexpr int AQ = 64, AK = 64, AD = 64, APAD = 72;   // query tile, key tile, head_dim, padded smem row (halves)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> struct Mma;
template <> struct Mma<__half> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __half2 h = __floats2half2_rn(x, y); return *(uint32_t*)&h; }
};
template <> struct Mma<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __nv_bfloat162 h = __floats2bfloat162_rn(x, y); return *(uint32_t*)&h; }
};

template <typename T16>
__global__ void __launch_bounds__(128)
attention_mma_kernel(const T16* __restrict__ qkv, T16* __restrict__ out, BatchGeom g, int nh, int nkv, float scale_log2e, int window) {
  __shared__ __align__(16) T16 Qs[AQ * APAD];
  __shared__ __align__(16) T16 Ks[AK * APAD];
  __shared__ __align__(16) T16 Vs[AK * APAD];
  pdl_trigger();
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
  const int len = g.len_frames[b], beg = g.row_begin ? g.row_begin[b] : 0;
  if (q0 >= len) return;
  const int hk = h / (nh / nkv), ld = (nh + 2 * nkv) * AD;
  const T16* base = qkv + (int64_t)b * g.Tmax * ld;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto load_tile = [&](T16* dst, int row0, int col0) {   // 64 rows x 64 halves, 16-byte loads, zero outside [beg, len)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128, r = idx >> 3, c8 = idx & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < len && row0 + r >= beg) v = __ldg((const uint4*)(base + (int64_t)(row0 + r) * ld + col0 + c8 * 8));
      *(uint4*)&dst[r * APAD + c8 * 8] = v;
    }
  };
  load_tile(Qs, q0, h * AD);
  __syncthreads();
  // Q fragments: 16 rows of this warp x 64 d = 4 k-steps
  uint32_t qa[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(qa[kk], smem_addr(&Qs[(warp * 16 + (lane & 15)) * APAD + kk * 16 + (lane >> 4) * 8]));

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;        // the two query rows this thread holds

  int k_begin = 0, k_end = len;
  if (window > 0) { k_begin = max(beg, max(0, q0 - window + 1)) / AK * AK; k_end = min(len, q0 + AQ); }
  else k_begin = beg / AK * AK;
  for (int kt = k_begin; kt < k_end; kt += AK) {
    __syncthreads();                       // previous tile fully consumed
    load_tile(Ks, kt, (nh + hk) * AD);
    load_tile(Vs, kt, (nh + nkv + hk) * AD);
    __syncthreads();
    // S = Q K^T : 8 key n-tiles x 4 k-steps
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {   // d 0..31, 32..63
        uint32_t kb[4];
        ldsm_x4(kb, smem_addr(&Ks[(j * 8 + (lane & 7)) * APAD + half * 32 + (lane >> 3) * 8]));
        Mma<T16>::run(s[j], qa[half * 2], kb[0], kb[1]);
        Mma<T16>::run(s[j], qa[half * 2 + 1], kb[2], kb[3]);
      }
    }
    // mask + online softmax (base-2, pre-scaled)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kt + j * 8 + 2 * (lane & 3) + (e & 1), tq = (e < 2) ? r0 : r1;
        bool ok = key < len && key >= beg;
        if (window > 0) ok = ok && key <= tq && (tq - key) < window;
        s[j][e] = ok ? s[j][e] * scale_log2e : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;   // fully masked so far
    const float c0 = exp2f(m0 - ms0), c1 = exp2f(m1 - ms1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] - ms0); s[j][1] = exp2f(s[j][1] - ms0);
      s[j][2] = exp2f(s[j][2] - ms1); s[j][3] = exp2f(s[j][3] - ms1);
      sum0 += s[j][0] + s[j][1];
      sum1 += s[j][2] + s[j][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    l0 = l0 * c0 + sum0; l1 = l1 * c1 + sum1;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
    // O += P V : 4 key k-steps x 8 d n-tiles; P comes straight from the S accumulators
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = Mma<T16>::pack(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = Mma<T16>::pack(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = Mma<T16>::pack(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = Mma<T16>::pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jd = 0; jd < 8; jd += 2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, smem_addr(&Vs[(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * APAD + (jd + (lane >> 4)) * 8]));
        Mma<T16>::run(o[jd], pa, vb[0], vb[1]);
        Mma<T16>::run(o[jd + 1], pa, vb[2], vb[3]);
      }
    }
  }
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  T16* ob = out + (int64_t)b * g.Tmax * (nh * AD) + h * AD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = j * 8 + 2 * (lane & 3);
    if (r0 < len) *(uint32_t*)&ob[(int64_t)r0 * (nh * AD) + d] = Mma<T16>::pack(o[j][0] * i0, o[j][1] * i0);
    if (r1 < len) *(uint32_t*)&ob[(int64_t)r1 * (nh * AD) + d] = Mma<T16>::pack(o[j][2] * i1, o[j][3] * i1);
  }
}


// ---- tail: outConv (C -> 1, k = 7, causal) + bias + clip (ST.swift:674-678, 688, 781) -------------------------------------
// out[t] = b + sum_j sum_c w[j][c] a[t-6+j][c] = sum_j P[t-6+j][j] with P = A[rows x C] . W^T[C x 8] (7 taps + one zero
// column): a GEMM with N = 8, i.e. exactly one HMMA m16n8k16 n-tile.  The scalar version spent ~1400 instructions per output
// sample (672 conversions + 672 FMAs) and ran at 1.7 TB/s; here a 16-row group costs C/16 ldmatrix + 2*C/16 HMMA.  The fp32
// weights are split into hi + lo 16-bit halves (two MMAs) so the result keeps fp32-weight accuracy.
constexpr int TT_ROWS = 128, TT_HALO = 6, TT_IN = 144;   // outputs per CTA; input rows = 134 rounded up to 9 groups of 16

template <typename T16, int C>
__global__ void __launch_bounds__(128)
tail_mma_kernel(const T16* __restrict__ a, int64_t a_bstride, const float* __restrict__ w /*[7][C]*/, float bias, float* __restrict__ pcm,
                const int64_t* __restrict__ pcm_base, float* __restrict__ tap, int64_t tap_bstride, BatchGeom g, int rows_per_frame,
                int tiles_per_utt, int tiles_per_cta) {
  constexpr int STRIDE = C + 8, V = C / 8, KS = C / 16;
  __shared__ __align__(16) T16 tile[TT_IN * STRIDE];
  __shared__ float P[TT_IN][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tg = lane & 3;
  // B fragments (col-major [K][8]): b0 = W[k = 2tg, 2tg+1][n = gq], b1 = W[k = 2tg+8, +9][n = gq]; tap 7 is a zero column.
  // Built ONCE per CTA: a CTA walks `tiles_per_cta` consecutive tiles (the set-up is ~150 instructions per thread, as much as the
  // work on one tile).
  uint32_t bh[KS][2], bl[KS][2];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = kk * 16 + 2 * tg + 8 * h;
      const float w0 = gq < 7 ? __ldg(w + gq * C + c) : 0.f, w1 = gq < 7 ? __ldg(w + gq * C + c + 1) : 0.f;
      const uint32_t hi = Mma<T16>::pack(w0, w1);
      float2 hf;
      if (sizeof(T16) == 2 && std::is_same<T16, __half>::value) hf = __half22float2(*(const __half2*)&hi);
      else hf = __bfloat1622float2(*(const __nv_bfloat162*)&hi);
      bh[kk][h] = hi;
      bl[kk][h] = Mma<T16>::pack(w0 - hf.x, w1 - hf.y);
    }
  const int64_t slot_rows = (int64_t)g.Tmax * rows_per_frame;
  const long long total_tiles = (long long)g.B * tiles_per_utt;
  for (int it = 0; it < tiles_per_cta; ++it) {
    const long long tile_id = (long long)blockIdx.x * tiles_per_cta + it;
    if (tile_id >= total_tiles) break;
    const int b = (int)(tile_id / tiles_per_utt);
    const int64_t t0 = (int64_t)(tile_id % tiles_per_utt) * TT_ROWS;
    const int64_t valid = (int64_t)g.len_frames[b] * rows_per_frame;
    if (t0 >= valid) continue;                                   // block-uniform
    const T16* ab = a + (int64_t)b * a_bstride;
    // all of a thread's 16-byte loads are issued before the first store: one HBM round trip per tile instead of one per chunk
    constexpr int NLD = (TT_IN * V + 127) / 128;
    uint4 ld[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + i * 128, row = idx / V, c8 = idx % V;
      const int64_t tin = t0 - TT_HALO + row;
      ld[i] = make_uint4(0, 0, 0, 0);
      if (idx < TT_IN * V && tin >= 0 && tin < slot_rows && row < TT_ROWS + TT_HALO) ld[i] = __ldg((const uint4*)(ab + tin * C + c8 * 8));
    }
    __syncthreads();                                             // the previous tile's reads of `tile` and `P` are over
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + i * 128, row = idx / V, c8 = idx % V;
      if (idx < TT_IN * V) *(uint4*)&tile[row * STRIDE + c8 * 8] = ld[i];
    }
    __syncthreads();
    for (int grp = warp; grp < TT_IN / 16; grp += 4) {
      float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        uint32_t af[4];
        ldsm_x4(af, smem_addr(&tile[(grp * 16 + (lane & 15)) * STRIDE + kk * 16 + (lane >> 4) * 8]));
        Mma<T16>::run(c4, af, bh[kk][0], bh[kk][1]);
        Mma<T16>::run(c4, af, bl[kk][0], bl[kk][1]);
      }
      *(float2*)&P[grp * 16 + gq][2 * tg] = make_float2(c4[0], c4[1]);
      *(float2*)&P[grp * 16 + gq + 8][2 * tg] = make_float2(c4[2], c4[3]);
    }
    __syncthreads();
    const int64_t t = t0 + tid;
    if (t < valid) {
      float v = bias;
#pragma unroll
      for (int j = 0; j < 7; ++j) v += P[tid + j][j];      // input row t-6+j sits at tile row (t - t0) + j
      if (tap) tap[(int64_t)b * tap_bstride + t] = v;
      store_pcm(pcm, pcm_base[b] + t, v, g.pcm_i16);
    }
  }
}

// ---- outConv from the partial products the last residual unit left (kernels_res96.cu, tail mode) -----------------------------
// P[row][16] fp32: columns 0..6 = a[row] . hi(w[j]), 8..14 = a[row] . lo(w[j]) (w = hi + lo in the operand type: the 16-bit
// tensor-core product keeps fp32 weights).  out[t] = bias + sum_j (P[t-6+j][j] + P[t-6+j][8+j]); rows before the utterance are
// the causal zero padding.  64 B per row come in, 4 B (or 2) go out: the 192-byte activation row never exists in HBM.
constexpr int TP_ROWS = 256, TP_IN = TP_ROWS + 6, TP_LD = 17;
__global__ void __launch_bounds__(TP_ROWS)
tail_from_partials_kernel(const float* __restrict__ P, int64_t p_bstride, float bias, float* __restrict__ pcm, const int64_t* __restrict__ pcm_base,
                          float* __restrict__ tap, int64_t tap_bstride, BatchGeom g, int rows_per_frame, int tiles_per_utt) {
  __shared__ float sp[TP_IN * TP_LD];
  pdl_trigger();
  const int b = blockIdx.x / tiles_per_utt, tid = threadIdx.x;
  const int64_t t0 = (int64_t)(blockIdx.x % tiles_per_utt) * TP_ROWS;
  const int64_t valid = (int64_t)g.len_frames[b] * rows_per_frame;
  if (t0 >= valid) return;
  const float* pb = P + (int64_t)b * p_bstride;
  for (int idx = tid; idx < TP_IN * 4; idx += TP_ROWS) {
    const int row = idx >> 2, c4 = idx & 3;
    const int64_t tin = t0 - 6 + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tin >= 0 && tin < valid) v = __ldg((const float4*)(pb + tin * 16) + c4);
    float* d = &sp[row * TP_LD + 4 * c4];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int64_t t = t0 + tid;
  if (t < valid) {
    float v = bias;
#pragma unroll
    for (int j = 0; j < 7; ++j) v += sp[(tid + j) * TP_LD + j] + sp[(tid + j) * TP_LD + 8 + j];
    if (tap) tap[(int64_t)b * tap_bstride + t] = v;
    store_pcm(pcm, pcm_base[b] + t, v, g.pcm_i16);
  }
}

// [16][C] operand tile for that product: rows 0..6 the taps rounded to the operand type, rows 8..14 the rounding residues, 7 and 15 zero
template <typename T16>
__global__ void tail_tile_kernel(const float* __restrict__ w, int C, T16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 8 * C) return;
  const int j = i / C, c = i % C;
  float hi = 0.f, lo = 0.f;
  if (j < 7) {
    const float v = w[j * C + c];
    const uint32_t pk = Mma<T16>::pack(v, 0.f);
    float2 hf;
    if (std::is_same<T16, __half>::value) hf = __half22float2(*(const __half2*)&pk);
    else hf = __bfloat1622float2(*(const __nv_bfloat162*)&pk);
    hi = hf.x; lo = v - hf.x;
  }
  const uint32_t ph = Mma<T16>::pack(hi, 0.f), pl = Mma<T16>::pack(lo, 0.f);
  out[j * C + c] = *(const T16*)&ph;
  out[(8 + j) * C + c] = *(const T16*)&pl;
}
}  // namespace

void launch_tail_tile(const float* w, int C, void* out16, int dtype, cudaStream_t s) {
  const unsigned blocks = (unsigned)((8 * C + 127) / 128);
  if (dtype == DT_F16) tail_tile_kernel<__half><<<blocks, 128, 0, s>>>(w, C, (__half*)out16);
  else tail_tile_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(w, C, (__nv_bfloat16*)out16);
}

void launch_tail_from_partials(const float* P, int64_t p_bstride, float bias, float* pcm, const int64_t* pcm_base, float* tap,
                               int64_t tap_bstride, const BatchGeom& g, int rows_per_frame, cudaStream_t s) {
  const int tiles = (int)(((int64_t)g.Tmax * rows_per_frame + TP_ROWS - 1) / TP_ROWS);
  tail_from_partials_kernel<<<(unsigned)((long long)g.B * tiles), TP_ROWS, 0, s>>>(P, p_bstride, bias, pcm, pcm_base, tap, tap_bstride, g,
                                                                                     rows_per_frame, tiles);
}

bool attention_mma_supported(int dtype, int hd) { return (dtype == DT_F16 || dtype == DT_BF16) && hd == 64; }

void launch_attention_mma(const void* qkv, int dtype, void* out, const BatchGeom& g, int nh, int nkv, float scale,
                          int causal_window, cudaStream_t s) {
  dim3 grid((unsigned)((g.Tmax + AQ - 1) / AQ), (unsigned)nh, (unsigned)g.B);
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == DT_F16) attention_mma_kernel<__half><<<grid, 128, 0, s>>>((const __half*)qkv, (__half*)out, g, nh, nkv, sl2, causal_window);
  else attention_mma_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, g, nh, nkv, sl2, causal_window);
}

bool tail_mma_supported(int dtype, int C) { return (dtype == DT_F16 || dtype == DT_BF16) && (C == 96 || C == 64 || C == 128); }

void launch_tail_mma(const void* a, int dtype, int64_t a_bstride, const float* w, float bias, int C, float* pcm, const int64_t* pcm_base,
                     float* tap, int64_t tap_bstride, const BatchGeom& g, int rows_per_frame, cudaStream_t s) {
  const int tiles = (int)(((int64_t)g.Tmax * rows_per_frame + TT_ROWS - 1) / TT_ROWS);
  const long long total = (long long)g.B * tiles;
  static const int per_env = [] { const char* e = getenv("Q3TTS_TAIL_TILES"); return e ? atoi(e) : 8; }();
  const int per_cta = (int)std::max<long long>(1, std::min<long long>(per_env, total / (148 * 12)));   // keep >= 12 CTAs per SM's worth of blocks
  const unsigned blocks = (unsigned)((total + per_cta - 1) / per_cta);
#define Q3_TAILM(T, CC) tail_mma_kernel<T, CC><<<blocks, 128, 0, s>>>((const T*)a, a_bstride, w, bias, pcm, pcm_base, tap, tap_bstride, g, rows_per_frame, tiles, per_cta)
  if (dtype == DT_F16) { if (C == 96) Q3_TAILM(__half, 96); else if (C == 64) Q3_TAILM(__half, 64); else Q3_TAILM(__half, 128); }
  else { if (C == 96) Q3_TAILM(__nv_bfloat16, 96); else if (C == 64) Q3_TAILM(__nv_bfloat16, 64); else Q3_TAILM(__nv_bfloat16, 128); }
#undef Q3_TAILM
}

}  // namespace q3
