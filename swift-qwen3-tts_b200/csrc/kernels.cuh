// Kernel-level interface of libqwen3tts_cuda (internal).  All activations are channels-last:
// a stage tensor is [B, rows_per_utt (stride), C] with C contiguous, so every conv / linear /
// transposed conv of the decoder is a GEMM whose A operand is a row-shifted view of its input
// (tap j of a causal conv with dilation d reads row t - (taps-1-j)*d; rows < 0 are zero).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace q3 {

enum ActKind : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SWIGLU = 2, ACT_GELU_TANH = 3 };   // GELU_TANH: the encoder's geluApprox (STE.swift:1080-1082), CUDA-core engine only
enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

// Batch geometry shared by every kernel of one launch chain (one micro-batch).
struct BatchGeom {
  int B;                   // utterances in this micro-batch
  int Tmax;                // frames per utterance slot (row stride at rate 1)
  const int* len_frames;   // [B] device: valid frames per utterance
  long long valid_frames;  // host copy of sum(len_frames) (work-list sizing of the strip-walking kernels)
  const int* row_begin;    // [B] device or null: first valid row of a slot (attention only: streaming keeps its KV history
                           //   right-aligned in front of the new frames, so a young stream's slot starts late)
  int pcm_i16 = 0;         // the tail writes 16-bit PCM (Int16(clamp(x,-1,1) * 32767), main.swift:158-160) instead of float
};

#ifdef __CUDACC__
// Final sample store of the tail kernels: clip (ST.swift:781), then float or the demo's WAV conversion (truncation toward zero).
__device__ __forceinline__ void store_pcm(float* pcm, long long i, float v, int i16) {
  v = fminf(fmaxf(v, -1.0f), 1.0f);
  if (i16) ((short*)pcm)[i] = (short)__float2int_rz(v * 32767.0f);
  else pcm[i] = v;
}
#endif

// Generic "multi-tap GEMM": out[b,t,n] = epi( sum_{j<taps} sum_c A[b, t-(taps-1-j)*dil, c] * W[j,n,c] ).
//  - plain conv:        W[j,n,c] = w_mlx[n,j,c]
//  - linear:            taps = 1
//  - transposed conv k=2r,s=r: taps = 2, N = r*Cout, W[1,p*Cout+co,c] = w_mlx[co,p,c], W[0,..] = w_mlx[co,p+r,c];
//    GEMM row t of [T, r*Cout] is rows t*r..t*r+r-1 of the [T*r, Cout] output (ST.swift:339-353).
struct ConvGemmParams {
  const void* A;  int lda;  int64_t a_bstride;   // elements; rows of utterance b start at A + b*a_bstride
  const void* W;                                  // [taps][N][Cin], Cin contiguous
  int rows_per_frame;                             // GEMM rows per codec frame at this stage
  int N, Cin, taps, dil;
  // epilogue: v = acc (+ bias[n]); v = act(v); v = res[b,t,n] + scale[n]*v (if res); y = v; a = snake(v)
  const float* bias;                              // [N] or null
  int act;                                        // ActKind (SWIGLU: columns (2i,2i+1) = (gate,up) -> out col i)
  const void* res; int ldres; int64_t res_bstride;// residual (stream dtype), or null
  const float* scale;                             // [N] layer-scale / gamma or null (only with res)
  void* out_y; int ldy; int64_t y_bstride;        // stream output or null
  void* out_a; int lda_out; int64_t ao_bstride;   // operand output or null
  const float* snake_ea;                          // [N] exp(alpha) or null: out_a = v + snake_ib*sin^2(v*ea)
  const float* snake_ib;                          // [N] 1/(exp(beta)+1e-9)
  void* out_tap;  int ldt; int64_t tap_bstride;   // optional fp32 copy of v (pre-snake) for stage taps
  int a_elu;                                      // CUDA-core engine only: out_a = elu(v) (the encoder's pre-activation, STE.swift:1075-1077)
};

// ---- CUDA-core GEMM: the fp32 parity engine, and the 16-bit fallback for shapes tcgen05 skips ----
// op_dtype: type of A, W, out_a.  y_dtype: type of res / out_y (fp32, or == op_dtype).  out_tap is fp32.
void launch_conv_gemm_simt(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype, cudaStream_t s);

// ---- row kernels, templated on storage types via DType tags -----------------------------------
// RVQ gather-and-sum: codes -> [rows, 2*half] = [ sum of semantic rows | sum of acoustic rows ]
// (sequential left-to-right fp32 adds, ST.swift:84-93).  code address = codes + code_base[b] + q*sq + t*st.
struct RvqParams {
  const int32_t* codes; const int64_t* code_base; int64_t sq, st;
  const float* const* tables;   // [num_quantizers] device pointers to [size_q, half] fp32 codebooks
  const int32_t* table_sizes;   // [num_quantizers]
  int num_q, num_sem, half;
  void* out; int out_dtype;     // [B*Tmax, 2*half]
  int* err_flag;                // set to 1 when a code id is out of range
};
void launch_rvq(const RvqParams& p, const BatchGeom& g, cudaStream_t s);

void launch_rmsnorm(const float* x, const float* w, float eps, void* out, int out_dtype, int64_t rows, int C,
                    cudaStream_t s);

// depthwise causal conv k=7 (+bias) followed by LayerNorm(eps) over C (ST.swift:389-393); x fp32 stream.
void launch_dwconv_ln(const float* x, const float* w7 /*[7][C]*/, const float* wb, const float* ln_w,
                      const float* ln_b, float eps, void* out, int out_dtype, const BatchGeom& g,
                      int rows_per_frame, int C, cudaStream_t s);

// Attention over one utterance's frames.  qkv: [B*Tmax, (nh+2*nkv)*hd] (q | k | v), out [B*Tmax, nh*hd].
void launch_attention(const void* qkv, int qkv_dtype, void* out, int out_dtype, const BatchGeom& g, int nh,
                      int nkv, int hd, float scale, int causal_window /*0 = full*/, cudaStream_t s);

// ---- speech-tokenizer ENCODER (row N3), fp32, kernels_enc.cu --------------------------------------------------------------
// First conv of the Seanet encoder: 1 -> C channels, k taps, causal (STE.swift:405-415): y = x' (stream), a = elu(x') (operand).
// out_a (float32) and / or out_h3 (split form, [lo' | hi | hi] per row) may be null
void launch_enc_init_conv(const float* audio, int64_t audio_bstride, const float* w /*[k][C]*/, const float* bias, int k, int C,
                          float* out_y, float* out_a, __half* out_h3, int64_t out_bstride, const BatchGeom& g, cudaStream_t s);
// LayerNorm over C (biased variance, eps inside the sqrt, affine), one warp per valid row of [B, Tmax, C]
void launch_layernorm(const float* x, const float* w, const float* b, float eps, float* out, const BatchGeom& g, int C, cudaStream_t s);
// MLX RoPE(dimensions = hd, traditional = false): pairs (i, i + hd/2) of the first `heads` heads of every valid row of
// qkv [B, Tmax, ld], angle = t * inv_freq[i] (inv_freq [hd/2], host-computed), position t = row index in the utterance
void launch_rope(float* qkv, int ld, int heads, int hd, const float* inv_freq, const BatchGeom& g, cudaStream_t s);
// One layer of EncoderResidualVectorQuantization.encode (STE.swift:816-829) after the GEMM score = x E^T - c2:
// idx = first argmax of score[row, :K]; codes[b*code_bstride + t] = idx; resid[row, :D] -= E[idx, :D]
void launch_vq_select(const float* score, int K, const float* E, int D, float* resid, int32_t* codes, int64_t code_bstride,
                      const BatchGeom& g, cudaStream_t s);

// Split-precision operands for the tensor-core engine of the encoder.  A float32 value a is carried as two fp16 numbers,
// hi = fp16(a) and lo' = fp16((a - hi) * 2048) (the factor keeps lo' out of fp16's subnormal range).  One GEMM with K' = 3 K,
//   A' = [a_lo' | a_hi | a_hi],   W' = [w_hi | w_lo' | 2048 w_hi]   =>   A'.W' = 2048 (a_lo w_hi + a_hi w_lo + a_hi w_hi)
// accumulates in float32 inside tensor memory: ~22 mantissa bits per operand, every scaling an exact power of two.  The small terms
// come FIRST in K: the tensor core truncates at every accumulation step, relative to the accumulator's magnitude, so they must be
// summed while the accumulator is still small (measured: 3x the error otherwise).  The output (and the bias) carry the factor
// 2048, which the next element-wise pass removes (`y_scale`).
//  v = y * y_scale (+ res, scaled per column if `scale`);  out_x = v;  a = act(v) (0 none, 1 elu, 2 tanh-GELU);  out_a32 = a;
//  out_h3: rows in groups of `grp` (the consuming conv's stride; Tmax % grp == 0): [lo' of the group's rows | hi | hi], so that the
//  conv's view [Tmax / grp, 3 grp C] of the same memory keeps the three parts contiguous.
// All tensors are [B, Tmax, C] (out_h3: [B, Tmax, 3 C]); only rows < len are touched.  Any output may be null.
constexpr float kSplitScale = 2048.0f;
void launch_enc_split(const float* y, float y_scale, const float* res, const float* scale, int act, float* out_x, float* out_a32,
                      __half* out_h3, int grp, int C, const BatchGeom& g, cudaStream_t s);
// weights [rows][Cin] float32 -> [rows][3 Cin] fp16 in blocks of `inner` channels ([hi | lo' | 2048 hi] per block; inner = Cin here)
// *overflow (device int) is set when some |w| x 2048 does not fit fp16: such a GEMM must stay on the CUDA cores
void launch_expand_w3(const float* w, __half* out, int64_t rows, int Cin, int inner, int* overflow, cudaStream_t s);

// Tensor-core (mma.sync) flash-style attention for 16-bit operands, head_dim 64 (kernels_attn.cu).
bool attention_mma_supported(int dtype, int hd);
void launch_attention_mma(const void* qkv, int dtype, void* out, const BatchGeom& g, int nh, int nkv, float scale,
                          int causal_window, cudaStream_t s);

// Tail: causal conv C->1, k=7 on the (already outSnake-activated) operand, + bias, clip to [-1,1].
// pcm address = pcm + pcm_base[b] + t.  Optional unclipped fp32 copy for the "out_conv" tap.
// Codec-embedding sum for the Talker's next-step input (Q3.swift:485-491, 720-728, 927-935, 1157-1162):
//   out[t,:] = (((E0[c[t,0]] + E1[c[t,1]]) + E2[c[t,2]]) + ...), every add rounded to the tables' dtype (MLX adds 16-bit arrays in
//   fp32 and rounds the result: the sequential order and the intermediate roundings are part of the reference's result).
constexpr int kMaxCodeGroups = 32;
struct CodecEmbedParams {
  const void* tables[kMaxCodeGroups];   // [vocab[g], H] each, dtype `dtype`
  int vocab[kMaxCodeGroups];
  int groups, H, dtype;                 // DT_F32 / DT_F16 / DT_BF16
  const int32_t* codes;                 // [n, groups] frame-major
  void* out;                            // [n, H], dtype `dtype`
  long long n;
  int* err_flag;                        // set when a code is outside its table (the reference leaves it unchecked)
};
void launch_codec_embed_sum(const CodecEmbedParams& p, cudaStream_t s);

void launch_tail(const void* a, int a_dtype, int64_t a_bstride, const float* w /*[7][C]*/, float bias, int C,
                 float* pcm, const int64_t* pcm_base, float* tap, int64_t tap_bstride, const BatchGeom& g,
                 int rows_per_frame, cudaStream_t s);

// Same op on the warp-level tensor path for 16-bit operands (kernels_attn.cu): [rows x C] . [C x 8] with HMMA, then the 7-tap diagonal sum.
// Programmatic dependent launch (one decode = ~95 launches in one stream).  Every kernel of the chain calls pdl_trigger() early, which
// lets the NEXT kernel of the stream -- if it was launched with the programmatic-serialization attribute (pdl_enabled()) -- get its
// CTAs resident and run its prologue (barrier init, tensor-memory allocation, tensor-map prefetch, constants) under this kernel's
// tail; such a kernel calls pdl_wait() before it first touches memory an earlier kernel wrote (pdl_wait returns when the preceding
// grid has completed and its writes are visible; without the attribute both are no-ops).  Q3TTS_PDL=0 switches the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
void pdl_suspend(bool on);   // this thread's launches: plain stream order while suspended (per-launch event timing measures serialised kernels)
bool tail_mma_supported(int dtype, int C);
// outConv split in two: the last residual unit (tail mode) multiplies its on-chip output with the [16][C] hi/lo tap tile and stores
// 16 partial products per row; this kernel adds the seven that belong to each sample, then bias + clip (ST.swift:687-688, 781)
void launch_tail_tile(const float* w /*[7][C]*/, int C, void* out16 /*[16][C]*/, int dtype, cudaStream_t s);
void launch_tail_from_partials(const float* P, int64_t p_bstride, float bias, float* pcm, const int64_t* pcm_base, float* tap,
                               int64_t tap_bstride, const BatchGeom& g, int rows_per_frame, cudaStream_t s);
void launch_tail_mma(const void* a, int dtype, int64_t a_bstride, const float* w, float bias, int C, float* pcm,
                     const int64_t* pcm_base, float* tap, int64_t tap_bstride, const BatchGeom& g, int rows_per_frame, cudaStream_t s);

// audioLengths = count(code[b,t,0] > 0) * rate (ST.swift:831-833).
void launch_lengths(const int32_t* codes, const int64_t* code_base, int64_t st, const int* len_frames, int B,
                    int rate, int32_t* out, cudaStream_t s);

// Stage tap: [B, rows(stride), C] (any dtype) -> fp32 NCT [B, C, L].
void launch_tap_copy(const void* src, int dtype, int64_t bstride, int ld, float* dst, int B, int C, int64_t L,
                     cudaStream_t s);

// Batched block copy (streaming state shuffles): item i copies `bytes` (multiple of 16) from src to dst; a null src zero-fills.
struct CopyItem { const void* src; void* dst; long long bytes; };
void launch_block_copy(const CopyItem* d_items, int n_items, cudaStream_t s);

// float <-> 16-bit conversion of packed weights.
void launch_convert(const float* src, void* dst, int dtype, int64_t n, cudaStream_t s);

// ---- tcgen05 tensor-core path (kernels_tc2.cu) -------------------------------------------------
// A/W/out_a are `op_dtype` (F16/BF16) operands, fp32 accumulate in TMEM; res/out_y are `y_dtype`.
// Halo tiles (one TMA box serves every tap of a 64-channel block), CTA-pair MMA, smem-resident weights, fused epilogues.
bool tc2_supported(const ConvGemmParams& p, int op_dtype);
// Optional second stage fused into the same kernel (residual unit, ST.swift:430-437): `p` is the unit's k=7 conv with bias + SnakeBeta
// (p.out_a is not written: the activated result stays in smem as the operand of the 1x1 conv W1 [N][N]), then
//   v = W1 . that + bias + res;   out_y = v;   out_a = v + ib * sin^2(v * ea)   (out_a / ea / ib may be null: stream only).
// out_a must not alias p.A (later tiles still read halo rows of p.A); res == out_y (in place) is fine.
struct FusedConv1 {
  const void* W1; const float* bias;
  const void* res; void* out_y; void* out_a;
  const float* ea; const float* ib;
};
bool tc2_fuse_supported(const ConvGemmParams& p, int op_dtype);
cudaError_t launch_conv_gemm_tc2(const ConvGemmParams& p, const BatchGeom& g, int op_dtype, int y_dtype,
                                 cudaStream_t s, const FusedConv1* fuse = nullptr);

// ---- fused residual unit of the 96-channel block (kernels_res96.cu) ------------------------------------------------
//   out = X + conv1x1(snake2(conv7_dil(snake1(X)) + b7)) + b1        (ST.swift:430-437), or snake3(that) when ea3 != null.
// X / out are [B, Tmax*rows_per_frame, C] 16-bit channels-last tensors (out != X: tiles of other CTAs read X's halo rows).
struct ResUnitParams {
  const void* x_in; void* out;
  const void* w7; const void* w1;       // packed [7][C][C] and [1][C][C], 16-bit, Cin contiguous
  const float *b7, *b1;
  const float *ea1, *ib1, *ea2, *ib2;   // snake before conv7, between the convs
  const float *ea3, *ib3;               // optional: activation of the consumer, applied to the output
  int C, dil, rows_per_frame;
  // tail mode (the decoder's last unit, ea3 = outSnake): instead of `out`, the unit writes the outConv partial products
  // P[row][16] fp32 = snake3(X')[row] . wout[n] (wout: [16][C] 16-bit, launch_tail_tile) -- `out` is not written
  const void* wout = nullptr; float* p_out = nullptr;
  void* dbg;                            // optional device buffer [16][32] of clock64 stamps (pipeline debugging)
};
bool resunit96_supported(const ResUnitParams& p, int op_dtype);
cudaError_t launch_resunit96(const ResUnitParams& p, const BatchGeom& g, int op_dtype, cudaStream_t s);

}  // namespace q3
