// Tensor-core attention for the decoder's pre-transformer (ST.swift:512-528): softmax(scale * Q K^T [+ mask]) V per
// (utterance, head), head_dim 64, 16-bit operands, fp32 softmax statistics and output accumulators.
// Attention is 0.1-0.25 % of the decoder's FLOPs and T <= a few thousand, so this is a compact
// flash-style kernel on the warp-level mma.sync path (HMMA m16n8k16), not a tcgen05 pipeline: 4 warps x 16 query rows
// per CTA, 64-key K/V tiles staged in padded shared memory, ldmatrix(.trans) fragment loads, online softmax.
// Keys are limited to the utterance's own frames, so padded batch slots never leak into valid frames.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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

}  // namespace

bool attention_mma_supported(int dtype, int hd) { return (dtype == DT_F16 || dtype == DT_BF16) && hd == 64; }

void launch_attention_mma(const void* qkv, int dtype, void* out, const BatchGeom& g, int nh, int nkv, float scale,
                          int causal_window, cudaStream_t s) {
  dim3 grid((unsigned)((g.Tmax + AQ - 1) / AQ), (unsigned)nh, (unsigned)g.B);
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == DT_F16) attention_mma_kernel<__half><<<grid, 128, 0, s>>>((const __half*)qkv, (__half*)out, g, nh, nkv, sl2, causal_window);
  else attention_mma_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, g, nh, nkv, sl2, causal_window);
}

}  // namespace q3
