// Row kernels of the speech-tokenizer ENCODER (reference: Sources/Qwen3TTS/Models/SpeechTokenizerEncoder.swift,
// "STE.swift").  fp32 throughout: the reference evaluates the nearest-codebook search in float32 on purpose (STE.swift:750-757), and
// a code is an argmin -- there is no "close enough".  The convolutions / linears run on the CUDA-core multi-tap GEMM of
// kernels_f32.cu (a strided conv with k = 2s is a 2-tap GEMM on the input viewed as [frames, s * Cin]); what is here is the rest.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.cuh"

namespace q3 {
namespace {

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expf(v) - 1.0f; }
__device__ __forceinline__ void split1(float a, __half& hi, __half& lo) {
  hi = __float2half_rn(a);
  lo = __float2half_rn((a - __half2float(hi)) * kSplitScale);
}

// y[b, t, c] = bias[c] + sum_j w[j][c] * x[b, t - (k-1) + j]   (zeros before the utterance); one thread per (t, c), c fastest
__global__ void __launch_bounds__(256)
enc_init_conv_kernel(const float* __restrict__ audio, int64_t audio_bstride, const float* __restrict__ w, const float* __restrict__ bias, int k,
                     int C, float* __restrict__ out_y, float* __restrict__ out_a, __half* __restrict__ out_h3, int64_t out_bstride, BatchGeom g) {
  // one thread per (sample, group of 4 channels): 16-byte stores, 32-bit index arithmetic (the host checks Tmax * C < 2^31)
  const int b = blockIdx.y;
  const int len = g.len_frames[b];
  const int c4n = C >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = (int)(i / (unsigned)c4n);
  const int c = (int)(i - (unsigned)t * (unsigned)c4n) * 4;
  if (t >= len) return;
  const float* x = audio + (int64_t)b * audio_bstride;
  float4 acc = *(const float4*)(bias + c);
  for (int j = 0; j < k; ++j) {
    const int src = t - (k - 1) + j;
    if (src >= 0) {
      const float xv = x[src];
      const float4 wv = *(const float4*)(w + j * C + c);
      acc.x = fmaf(wv.x, xv, acc.x); acc.y = fmaf(wv.y, xv, acc.y); acc.z = fmaf(wv.z, xv, acc.z); acc.w = fmaf(wv.w, xv, acc.w);
    }
  }
  const int64_t o = (int64_t)b * out_bstride + (int64_t)t * C + c;
  *(float4*)(out_y + o) = acc;
  const float a[4] = {elu1(acc.x), elu1(acc.y), elu1(acc.z), elu1(acc.w)};
  if (out_a) *(float4*)(out_a + o) = make_float4(a[0], a[1], a[2], a[3]);
  if (out_h3) {   // the split form of the operand for a tensor-core consumer (kernels.cuh): [lo' | hi | hi] per row
    __half h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split1(a[q], h[q], l[q]);
    __half* d = out_h3 + 3 * ((int64_t)b * out_bstride + (int64_t)t * C) + c;
    *(uint2*)d = *(const uint2*)l;
    *(uint2*)(d + C) = *(const uint2*)h;
    *(uint2*)(d + 2 * C) = *(const uint2*)h;
  }
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bs, float eps, float* __restrict__ out, BatchGeom g,
                 int C) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)g.B * g.Tmax) return;
  const int b = (int)(row / g.Tmax), t = (int)(row % g.Tmax);
  if (t >= g.len_frames[b]) return;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float r = rsqrtf(v / (float)C + eps);
  for (int c = lane; c < C; c += 32) out[row * C + c] = (xr[c] - mean) * r * w[c] + bs[c];
}

// one thread per (row, head, i < hd/2)
__global__ void __launch_bounds__(256)
rope_kernel(float* __restrict__ qkv, int ld, int heads, int hd, const float* __restrict__ inv_freq, BatchGeom g) {
  const int half = hd >> 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = heads * half;
  const int64_t row = i / per_row;
  if (row >= (int64_t)g.B * g.Tmax) return;
  const int b = (int)(row / g.Tmax), t = (int)(row % g.Tmax);
  if (t >= g.len_frames[b]) return;
  const int r = (int)(i - row * per_row), h = r / half, j = r - h * half;
  float* p = qkv + row * ld + h * hd;
  float sn, cs;
  sincosf((float)t * inv_freq[j], &sn, &cs);
  const float x1 = p[j], x2 = p[j + half];
  p[j] = x1 * cs - x2 * sn;
  p[j + half] = x1 * sn + x2 * cs;
}

// one warp per valid row: first index of the maximum (== the reference's argMin of c2 - x.E^T: IEEE negation is exact), then the
// residual update in float32 (STE.swift:824-826)
__global__ void __launch_bounds__(256)
vq_select_kernel(const float* __restrict__ score, int K, const float* __restrict__ E, int D, float* __restrict__ resid, int32_t* __restrict__ codes,
                 int64_t code_bstride, BatchGeom g) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)g.B * g.Tmax) return;
  const int b = (int)(row / g.Tmax), t = (int)(row % g.Tmax);
  if (t >= g.len_frames[b]) return;
  const float* sr = score + row * K;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < K; c += 32) {
    const float v = sr[c];
    if (v > best) { best = v; bi = c; }          // ascending c per lane: keeps the lane's first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;                  // a row of NaNs: argMin would also return an index, keep the decode defined
  if (lane == 0) codes[(int64_t)b * code_bstride + t] = bi;
  const float* e = E + (int64_t)bi * D;
  float* rr = resid + row * D;
  for (int c = lane; c < D; c += 32) rr[c] = rr[c] - e[c];
}

// one thread per (row, 4 channels)
__global__ void __launch_bounds__(256)
enc_split_kernel(const float* __restrict__ y, float y_scale, const float* __restrict__ res, const float* __restrict__ scale, int act,
                 float* __restrict__ out_x, float* __restrict__ out_a32, __half* __restrict__ out_h3, int grp, int C, BatchGeom g) {
  const int c4n = C >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = i / c4n;
  if (row >= (int64_t)g.B * g.Tmax) return;
  const int b = (int)(row / g.Tmax), t = (int)(row % g.Tmax);
  if (t >= g.len_frames[b]) return;
  const int c = (int)(i - row * c4n) * 4;
  const int64_t o = row * C + c;
  float4 v = *(const float4*)(y + o);
  v.x *= y_scale; v.y *= y_scale; v.z *= y_scale; v.w *= y_scale;
  if (res) {
    const float4 r = *(const float4*)(res + o);
    if (scale) {
      const float4 sc = *(const float4*)(scale + c);
      v.x = r.x + sc.x * v.x; v.y = r.y + sc.y * v.y; v.z = r.z + sc.z * v.z; v.w = r.w + sc.w * v.w;
    } else {
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
  }
  if (out_x) *(float4*)(out_x + o) = v;
  float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (act == 1) a[k] = elu1(a[k]);
    else if (act == 2) a[k] = a[k] * 0.5f * (1.0f + tanhf(0.7978845608f * (a[k] + 0.044715f * (a[k] * a[k] * a[k]))));
  }
  if (out_a32) *(float4*)(out_a32 + o) = make_float4(a[0], a[1], a[2], a[3]);
  if (out_h3) {
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split1(a[k], h[k], l[k]);
    // rows are grouped `grp` at a time (the stride of the consuming conv): [lo' of the group | hi of the group | hi of the group]
    const int64_t grow = (int64_t)b * g.Tmax + (t / grp) * grp;          // first row of this row's group (Tmax is a multiple of grp)
    __half* d = out_h3 + grow * 3 * C + (int64_t)(t % grp) * C + c;
    const int64_t part = (int64_t)grp * C;
    *(uint2*)d = *(const uint2*)l;
    *(uint2*)(d + part) = *(const uint2*)h;
    *(uint2*)(d + 2 * part) = *(const uint2*)h;
  }
}

__global__ void __launch_bounds__(256)
expand_w3_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t rows, int Cin, int inner, int* __restrict__ overflow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Cin) return;
  if (!(fabsf(w[i]) * kSplitScale < 60000.0f)) atomicOr(overflow, 1);   // 2048 x hi would leave fp16's range (or w is not finite)
  const int64_t r = i / Cin;
  const int k = (int)(i - r * Cin), j = k / inner, c = k - j * inner;
  __half hi, lo;
  split1(w[i], hi, lo);
  __half* d = out + r * 3 * Cin + (int64_t)j * 3 * inner + c;
  d[0] = hi;                                                   // meets the operand's lo'
  d[inner] = lo;                                               // meets hi
  d[2 * inner] = __float2half_rn(__half2float(hi) * kSplitScale);   // meets hi; exact (a power of two, range checked above)
}
}  // namespace

void launch_enc_split(const float* y, float y_scale, const float* res, const float* scale, int act, float* out_x, float* out_a32,
                      __half* out_h3, int grp, int C, const BatchGeom& g, cudaStream_t s) {
  const int64_t n = (int64_t)g.B * g.Tmax * (C / 4);
  enc_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, y_scale, res, scale, act, out_x, out_a32, out_h3, grp < 1 ? 1 : grp, C, g);
}

void launch_expand_w3(const float* w, __half* out, int64_t rows, int Cin, int inner, int* overflow, cudaStream_t s) {
  const int64_t n = rows * Cin;
  expand_w3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, out, rows, Cin, inner, overflow);
}

void launch_enc_init_conv(const float* audio, int64_t audio_bstride, const float* w, const float* bias, int k, int C, float* out_y,
                          float* out_a, __half* out_h3, int64_t out_bstride, const BatchGeom& g, cudaStream_t s) {
  const int64_t n = (int64_t)g.Tmax * (C / 4);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)g.B);
  enc_init_conv_kernel<<<grid, 256, 0, s>>>(audio, audio_bstride, w, bias, k, C, out_y, out_a, out_h3, out_bstride, g);
}

void launch_layernorm(const float* x, const float* w, const float* b, float eps, float* out, const BatchGeom& g, int C, cudaStream_t s) {
  const int64_t rows = (int64_t)g.B * g.Tmax;
  layernorm_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(x, w, b, eps, out, g, C);
}

void launch_rope(float* qkv, int ld, int heads, int hd, const float* inv_freq, const BatchGeom& g, cudaStream_t s) {
  const int64_t n = (int64_t)g.B * g.Tmax * heads * (hd / 2);
  rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(qkv, ld, heads, hd, inv_freq, g);
}

void launch_vq_select(const float* score, int K, const float* E, int D, float* resid, int32_t* codes, int64_t code_bstride,
                      const BatchGeom& g, cudaStream_t s) {
  const int64_t rows = (int64_t)g.B * g.Tmax;
  vq_select_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(score, K, E, D, resid, codes, code_bstride, g);
}

}  // namespace q3
