// Speech-tokenizer ENCODER (audio -> codes), SURVEY 8(f) row N3: host engine.  Reference: Qwen3TTSSpeechTokenizerEncoder.encode,
// Sources/Qwen3TTS/Models/SpeechTokenizerEncoder.swift:1031-1056 ("STE.swift") and the modules it drives (cited per stage below).
// Everything is float32 and channels-last [B, rows, C]; see encoder.hpp for the data layout.
#include "encoder.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

#define ENC_CUDA_OK(x)                                                                                          \
  do {                                                                                                          \
    cudaError_t e_ = (x);                                                                                       \
    if (e_ != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));           \
  } while (0)

namespace q3 {
namespace {

float* upload(EncoderModel& m, const std::vector<float>& v) {
  float* d = nullptr;
  ENC_CUDA_OK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
  m.allocs.push_back(d);
  ENC_CUDA_OK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, m.stream));
  ENC_CUDA_OK(cudaStreamSynchronize(m.stream));
  return d;
}

const HostTensor& T(const EncoderCheckpoint& ck, const std::string& k) {
  auto it = ck.tensors.find(k);
  if (it == ck.tensors.end()) throw Error(Q3TTS_EFORMAT, "missing encoder tensor " + k);
  return it->second;
}

// MLX conv weight [Cout, K, Cin] -> multi-tap GEMM weights [taps][N][Cin'].
//  stride 1: taps = K, W[j][n][c] = w[n][j][c]                                   (cross-correlation, STE.swift:262-284)
//  stride s, K = 2s (STE.swift:369-379, 688-698): the input is viewed as [frames, s*Cin]; left padding K - s = s is exactly one
//  row of that view, so y[t] = W0 . row(t-1) + W1 . row(t) with W0[n][j*Cin + c] = w[n][j][c], W1[n][j*Cin + c] = w[n][s + j][c].
EncGemm pack_conv(EncoderModel& m, const HostTensor& w, const HostTensor* bias, int stride) {
  const int N = (int)w.shape[0], K = (int)w.shape[1], Cin = (int)w.shape[2];
  EncGemm g;
  g.N = N;
  std::vector<float> packed;
  if (stride == 1) {
    g.taps = K; g.Cin = Cin;
    packed.resize((size_t)K * N * Cin);
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < K; ++j)
        for (int c = 0; c < Cin; ++c) packed[((size_t)j * N + n) * Cin + c] = w.data[((size_t)n * K + j) * Cin + c];
  } else {
    if (K != 2 * stride) throw Error(Q3TTS_EFORMAT, "encoder: a strided conv must have kernel = 2 * stride");
    g.taps = 2; g.Cin = stride * Cin;
    packed.resize((size_t)2 * N * g.Cin);
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < K; ++j)
        for (int c = 0; c < Cin; ++c)
          packed[((size_t)(j / stride) * N + n) * g.Cin + (size_t)(j % stride) * Cin + c] = w.data[((size_t)n * K + j) * Cin + c];
  }
  if (g.Cin % 4) throw Error(Q3TTS_EFORMAT, "encoder: channel counts must be multiples of 4");
  g.w = upload(m, packed);
  g.bias = bias ? upload(m, bias->data) : nullptr;
  return g;
}

EncGemm pack_linear(EncoderModel& m, const std::vector<const HostTensor*>& rows) {   // y = x W^T, W [out, in]; several matrices stacked
  EncGemm g;
  g.taps = 1; g.Cin = (int)rows[0]->shape[1];
  std::vector<float> packed;
  for (auto* r : rows) {
    if ((int)r->shape[1] != g.Cin) throw Error(Q3TTS_EFORMAT, "encoder: stacked linears differ in width");
    packed.insert(packed.end(), r->data.begin(), r->data.end());
    g.N += (int)r->shape[0];
  }
  if (g.Cin % 4) throw Error(Q3TTS_EFORMAT, "encoder: channel counts must be multiples of 4");
  g.w = upload(m, packed);
  return g;
}

void run_gemm(EncoderModel& m, const EncGemm& w, const BatchGeom& g, const float* A, int lda, int64_t a_bstride, const ConvGemmParams& epi) {
  ConvGemmParams p = epi;
  p.A = A; p.lda = lda; p.a_bstride = a_bstride;
  p.W = w.w; p.rows_per_frame = 1; p.N = w.N; p.Cin = w.Cin; p.taps = w.taps; p.dil = 1;
  if (!p.bias) p.bias = w.bias;
  launch_conv_gemm_simt(p, g, DT_F32, DT_F32, m.stream);
  ++m.launches;
}

}  // namespace

EncoderModel* encoder_create(const std::string& dir, const q3tts_options& opts) {
  EncoderCheckpoint ck;
  load_encoder_checkpoint(dir, &ck);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) throw Error(Q3TTS_ECUDA, "no CUDA device (this library has no CPU path)");
  int device = opts.device;
  if (device < 0) ENC_CUDA_OK(cudaGetDevice(&device));                      // -1 = the calling thread's current device, as q3tts_model_load
  if (device >= ndev) throw Error(Q3TTS_EINVAL, "device index out of range");
  if (opts.precision != Q3TTS_PREC_FP32 && opts.precision != Q3TTS_PREC_FP16)
    throw Error(Q3TTS_EINVAL, "encoder precision: Q3TTS_PREC_FP32 (CUDA cores) or Q3TTS_PREC_FP16 (tensor cores, split fp16 operands)");
  ENC_CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<EncoderModel> mp(new EncoderModel());
  EncoderModel& m = *mp;
  m.cfg = ck.cfg;
  m.device = device;
  m.tc = opts.precision == Q3TTS_PREC_FP16;
  m.num_parameters = ck.num_parameters;
  ENC_CUDA_OK(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
  const EncoderConfig& c = m.cfg;
  {  // init conv: [nf, k, 1] -> [k][nf]
    const HostTensor& w = T(ck, "encoder.encoder.init_conv1d.conv.conv.weight");
    std::vector<float> pk((size_t)c.kernel_size * c.num_filters);
    for (int n = 0; n < c.num_filters; ++n)
      for (int j = 0; j < c.kernel_size; ++j) pk[(size_t)j * c.num_filters + n] = w.data[(size_t)n * c.kernel_size + j];
    m.init_w = upload(m, pk);
    m.init_b = upload(m, T(ck, "encoder.encoder.init_conv1d.conv.conv.bias").data);
  }
  int mult = 1;
  for (int li = 0; li < c.n_ratios; ++li) {
    EncStage st;
    st.ratio = c.ratios[c.n_ratios - 1 - li];                                                 // ratios.reversed(), STE.swift:420
    st.dim = mult * c.num_filters;
    const std::string p = "encoder.encoder.layers." + std::to_string(li);
    st.res3 = pack_conv(m, T(ck, p + ".residuals.0.block.0.conv.conv.weight"), &T(ck, p + ".residuals.0.block.0.conv.conv.bias"), 1);
    st.res1 = pack_conv(m, T(ck, p + ".residuals.0.block.1.conv.conv.weight"), &T(ck, p + ".residuals.0.block.1.conv.conv.bias"), 1);
    st.down = pack_conv(m, T(ck, p + ".downsample.conv.conv.weight"), &T(ck, p + ".downsample.conv.conv.bias"), st.ratio);
    m.stages.push_back(st);
    mult *= 2;
  }
  m.final_conv = pack_conv(m, T(ck, "encoder.encoder.final_conv1d.conv.conv.weight"), &T(ck, "encoder.encoder.final_conv1d.conv.conv.bias"), 1);
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const std::string p = "encoder.encoder_transformer.transformer.layers." + std::to_string(i);
    EncLayer L;
    L.n1w = upload(m, T(ck, p + ".norm1.weight").data); L.n1b = upload(m, T(ck, p + ".norm1.bias").data);
    L.n2w = upload(m, T(ck, p + ".norm2.weight").data); L.n2b = upload(m, T(ck, p + ".norm2.bias").data);
    L.qkv = pack_linear(m, {&T(ck, p + ".self_attn.q_proj.weight"), &T(ck, p + ".self_attn.k_proj.weight"), &T(ck, p + ".self_attn.v_proj.weight")});
    L.o = pack_linear(m, {&T(ck, p + ".self_attn.o_proj.weight")});
    L.fc1 = pack_linear(m, {&T(ck, p + ".gating.linear1.weight")});
    L.fc2 = pack_linear(m, {&T(ck, p + ".gating.linear2.weight")});
    L.ls1 = upload(m, T(ck, p + ".layer_scale_1.scale").data);
    L.ls2 = upload(m, T(ck, p + ".layer_scale_2.scale").data);
    m.layers.push_back(L);
  }
  {  // MLX RoPE(dimensions: head_dim, traditional: false, base: rope_theta), STE.swift:494: inv_freq[i] = base^(-i / (hd/2))
    const int hd = c.hidden_size / c.num_attention_heads, half = hd / 2;
    std::vector<float> f((size_t)half);
    for (int i = 0; i < half; ++i) f[(size_t)i] = std::exp(-(float)i * (std::log(c.rope_theta) / (float)half));
    m.inv_freq = upload(m, f);
  }
  m.downsample = pack_conv(m, T(ck, "encoder.downsample.conv.conv.conv.weight"), nullptr, c.downsample_stride());
  for (int part = 0; part < 2; ++part) {
    const std::string q = std::string("encoder.quantizer.") + (part == 0 ? "rvq_first" : "rvq_rest");
    const HostTensor& pw = T(ck, q + ".input_proj.weight");                                   // [cb, 1, H]
    HostTensor lin;
    lin.shape = {pw.shape[0], pw.shape[2]};
    lin.data = pw.data;
    m.proj[part] = pack_linear(m, {&lin});
    const int n = part == 0 ? 1 : c.valid_quantizers - 1;
    for (int i = 0; i < n; ++i) {
      const std::string b = q + ".vq.layers." + std::to_string(i) + ".codebook";
      const HostTensor& sum = T(ck, b + ".embeddingSum");
      const HostTensor& usage = T(ck, b + ".clusterUsage");
      const int K = c.codebook_size, D = c.codebook_dim;
      std::vector<float> E((size_t)K * D), nc2((size_t)K);
      for (int r = 0; r < K; ++r) {                                                           // STE.swift:738-743
        const float den = std::max(usage.data[(size_t)r], 1e-5f);
        float s2 = 0.f;
        for (int d = 0; d < D; ++d) {
          const float e = sum.data[(size_t)r * D + d] / den;
          E[(size_t)r * D + d] = e;
          s2 += e * e;
        }
        nc2[(size_t)r] = -(s2 / 2.0f);
      }
      EncBook bk;
      bk.part = part;
      bk.score.taps = 1; bk.score.Cin = D; bk.score.N = K;
      bk.score.w = upload(m, E);
      bk.score.bias = upload(m, nc2);
      m.books.push_back(bk);
    }
  }
  if (m.tc) {
    int maxN = std::max(std::max(c.codebook_size, c.intermediate_size), (c.num_attention_heads + 2 * c.num_key_value_heads) * (c.hidden_size / c.num_attention_heads));
    for (auto& st : m.stages) maxN = std::max(maxN, 2 * st.dim);
    m.zeros = upload(m, std::vector<float>((size_t)maxN, 0.f));
    m.inv_split = upload(m, std::vector<float>((size_t)maxN, 1.0f / kSplitScale));
    for (size_t i = 0; i < m.layers.size(); ++i) {   // layer scale with the tensor-core GEMM's factor folded in: h += (ls / 2048) * (2048 y)
      const std::string p = "encoder.encoder_transformer.transformer.layers." + std::to_string(i);
      std::vector<float> a = T(ck, p + ".layer_scale_1.scale").data, b = T(ck, p + ".layer_scale_2.scale").data;
      for (auto& v : a) v /= kSplitScale;
      for (auto& v : b) v /= kSplitScale;
      m.layers[i].ls1_s = upload(m, a);
      m.layers[i].ls2_s = upload(m, b);
    }
    int* d_flag = nullptr;
    ENC_CUDA_OK(cudaMalloc(&d_flag, sizeof(int))); m.allocs.push_back(d_flag);
    auto split_w = [&](EncGemm& g) {
      const int64_t rows = (int64_t)g.taps * g.N;
      ENC_CUDA_OK(cudaMalloc(&g.w3, (size_t)rows * 3 * (size_t)g.Cin * 2)); m.allocs.push_back(g.w3);
      ENC_CUDA_OK(cudaMemsetAsync(d_flag, 0, sizeof(int), m.stream));
      launch_expand_w3(g.w, g.w3, rows, g.Cin, g.Cin, d_flag, m.stream);
      int flag = 0;
      ENC_CUDA_OK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, m.stream));
      ENC_CUDA_OK(cudaStreamSynchronize(m.stream));
      if (flag) { g.w3 = nullptr; return; }            // a weight too large for the scaled fp16 form: this GEMM stays on the CUDA cores
      if (g.bias) {
        std::vector<float> hb((size_t)g.N);
        ENC_CUDA_OK(cudaMemcpyAsync(hb.data(), g.bias, sizeof(float) * (size_t)g.N, cudaMemcpyDeviceToHost, m.stream));
        ENC_CUDA_OK(cudaStreamSynchronize(m.stream));
        for (auto& v : hb) v *= kSplitScale;
        g.bias_s = upload(m, hb);
      } else {
        g.bias_s = m.zeros;
      }
    };
    for (auto& st : m.stages) { split_w(st.res3); split_w(st.res1); split_w(st.down); }
    split_w(m.final_conv); split_w(m.downsample); split_w(m.proj[0]); split_w(m.proj[1]);
    for (auto& L : m.layers) { split_w(L.qkv); split_w(L.o); split_w(L.fc1); split_w(L.fc2); }
    for (auto& b : m.books) split_w(b.score);
    ENC_CUDA_OK(cudaStreamSynchronize(m.stream));
  }
  return mp.release();
}

void encoder_destroy(EncoderModel* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (void* p : m->allocs) cudaFree(p);
  if (m->arena) cudaFree(m->arena);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int64_t encoder_frames(const EncoderConfig& c, int64_t samples) {
  int64_t L = samples;
  for (int li = 0; li < c.n_ratios; ++li) {
    const int r = c.ratios[c.n_ratios - 1 - li];
    L = (L + r - 1) / r;                       // getExtraPaddingForConv1d pads the input to ceil(L / stride) frames, STE.swift:115-119
  }
  const int ds = c.downsample_stride();
  return (L + ds - 1) / ds;
}

namespace {
// The strided convs read whole strides: rows between a level's valid length and its padded length must be zero (the reference's
// "extra padding").  No kernel ever writes such a row, so the workspace is cleared once per (batch, length, taps) shape -- not per
// call: a memset of the 20-35 GB a 64 x 10 s batch touches would cost 7-12 ms -- and again whenever the shape (hence the plan) changes.
void zero_if_new_shape(EncoderModel& m, char* from, size_t bytes, int B, int64_t samples, cudaStream_t s) {
  if (m.zeroed_B == B && m.zeroed_samples == samples && m.zeroed_taps == m.taps_enabled) return;
  ENC_CUDA_OK(cudaMemsetAsync(from, 0, bytes, s));
  m.zeroed_B = B; m.zeroed_samples = samples; m.zeroed_taps = m.taps_enabled;
}

// ---- tensor-core engine -------------------------------------------------------------------------------------------------
// Same graph, but a GEMM is three tcgen05 products of split fp16 operands accumulated in float32 (kernels.cuh: kSplitScale),
//   Y = b + A_hi.W_hi;  Y += (A_hi.W_lo) / 2048;  Y += (A_lo.W_hi) / 2048
// through the decoder's multi-tap GEMM (kernels_tc2.cu, fp32 stream output, residual + per-column scale epilogue), and everything
// element-wise around it (residual, layer scale, elu / GELU, the split of the next operand) is ONE pass of enc_split_kernel.
// A GEMM the tensor-core kernel does not take (Cin < 64: the first stage's 1x1 conv, the tiny test architecture) runs on the
// CUDA-core engine from a float32 operand.
struct Opnd { float* f32 = nullptr; __half* h3 = nullptr; };   // float32 form [rows][C] and / or split form [rows][3 C]

bool tc_takes(const EncoderModel& m, const EncGemm& w) {
  ConvGemmParams p{};
  p.N = w.N; p.Cin = 3 * w.Cin; p.lda = 3 * w.Cin; p.taps = w.taps; p.dil = 1;
  return m.tc && w.w3 && tc2_supported(p, DT_F16);
}

// Y = (bias + conv(a)) x out_scale; returns out_scale (kSplitScale from the tensor-core GEMM, 1 from the CUDA-core fallback).
// lda / a_bstride are in elements of the float32 form; the split form has three times as many halves per row.
// With `res`: Y = res + s * (bias + conv(a)) at true scale (s = scale_f32 per column, or 1), res may be Y itself; returns 1.
float gemm_y(EncoderModel& m, const EncGemm& w, const BatchGeom& g, const Opnd& a, int lda, int64_t a_bstride, float* Y, int ldy, int64_t y_bstride,
             const float* res = nullptr, const float* scale_tc = nullptr, const float* scale_f32 = nullptr) {
  if (!tc_takes(m, w)) {
    if (!a.f32) throw Error(Q3TTS_EINVAL, "internal: CUDA-core GEMM without a float32 operand");
    ConvGemmParams e{};
    e.out_y = Y; e.ldy = ldy; e.y_bstride = y_bstride;
    if (res) { e.res = res; e.ldres = ldy; e.res_bstride = y_bstride; e.scale = scale_f32; }
    run_gemm(m, w, g, a.f32, lda, a_bstride, e);
    return 1.0f;
  }
  if (!a.h3) throw Error(Q3TTS_EINVAL, "internal: tensor-core GEMM without a split operand");
  ConvGemmParams p{};
  p.A = a.h3; p.lda = 3 * lda; p.a_bstride = 3 * a_bstride; p.W = w.w3;
  p.rows_per_frame = 1; p.N = w.N; p.Cin = 3 * w.Cin; p.taps = w.taps; p.dil = 1;
  p.bias = w.bias_s;
  p.out_y = Y; p.ldy = ldy; p.y_bstride = y_bstride;
  if (res) { p.res = res; p.ldres = ldy; p.res_bstride = y_bstride; p.scale = scale_tc ? scale_tc : m.inv_split; }   // the epilogue removes the factor
  cudaError_t err = launch_conv_gemm_tc2(p, g, DT_F16, DT_F32, m.stream);
  if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("encoder tcgen05 GEMM launch: ") + cudaGetErrorString(err));
  ++m.launches;
  return res ? 1.0f : kSplitScale;
}

void encode_tc(EncoderModel& m, const float* audio, int B, int64_t samples, int32_t* codes_out) {
  const EncoderConfig& c = m.cfg;
  ENC_CUDA_OK(cudaSetDevice(m.device));
  cudaStream_t s = m.stream;
  const int nst = c.n_ratios, ds = c.downsample_stride(), H = c.hidden_size, I = c.intermediate_size;
  const int nh = c.num_attention_heads, nkv = c.num_key_value_heads, hd = H / nh, QW = (nh + 2 * nkv) * hd;
  const int K = c.codebook_size, CB = c.codebook_dim, NQ = c.valid_quantizers;
  std::vector<int64_t> L((size_t)nst + 2), P((size_t)nst + 1);
  L[0] = samples;
  for (int li = 0; li < nst; ++li) {
    const int r = m.stages[(size_t)li].ratio;
    L[(size_t)li + 1] = (L[(size_t)li] + r - 1) / r;
    P[(size_t)li] = L[(size_t)li + 1] * r;
  }
  const int64_t Tq = (L[(size_t)nst] + ds - 1) / ds;
  L[(size_t)nst + 1] = Tq;
  P[(size_t)nst] = Tq * ds;

  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_len = take(sizeof(int) * (size_t)B * (size_t)(nst + 2));
  const size_t o_audio = take(sizeof(float) * (size_t)B * (size_t)samples);
  // per level: X (stream), Y (GEMM output / float32 operand), A3 (split operand, 3 C halves per row); per stage the hidden activation likewise
  // (the strided conv's operand D3 has its own buffer: its rows are grouped by the stride, and the padding rows of the last group must
  //  never have been written in another layout)
  std::vector<size_t> oX((size_t)nst + 1), oY((size_t)nst + 1), oA3((size_t)nst + 1), oD3((size_t)nst), oHY((size_t)nst), oH3((size_t)nst);
  for (int li = 0; li <= nst; ++li) {
    const size_t dim = (size_t)(li < nst ? m.stages[(size_t)li].dim : 2 * m.stages[(size_t)nst - 1].dim), n = (size_t)B * (size_t)P[(size_t)li];
    oX[(size_t)li] = take(4 * n * dim); oY[(size_t)li] = take(4 * n * dim); oA3[(size_t)li] = take(6 * n * dim);
    if (li < nst) { const size_t hid = dim / (size_t)c.compress; oHY[(size_t)li] = take(4 * n * hid); oH3[(size_t)li] = take(6 * n * hid); oD3[(size_t)li] = take(6 * n * dim); }
  }
  const size_t rowsT = (size_t)B * (size_t)P[(size_t)nst], WT = (size_t)std::max(std::max(H, I), QW);
  const size_t o_hs = take(4 * rowsT * (size_t)H), o_nb = take(4 * rowsT * (size_t)H), o_qkv = take(4 * rowsT * (size_t)QW), o_ao = take(4 * rowsT * (size_t)H);
  const size_t o_yt = take(4 * rowsT * WT), o_t3 = take(6 * rowsT * WT);
  const size_t o_ds3 = take(6 * rowsT * (size_t)H);               // the strided downsample's operand: its padding rows must stay zero
  const size_t o_tap = take(m.taps_enabled ? 4 * rowsT * (size_t)H : 0);
  const size_t rowsQ = (size_t)B * (size_t)Tq;
  const size_t o_d = take(4 * rowsQ * (size_t)H), o_r0 = take(4 * rowsQ * (size_t)CB), o_r1 = take(4 * rowsQ * (size_t)CB), o_sc = take(4 * rowsQ * (size_t)K);
  const size_t o_q3 = take(6 * rowsQ * (size_t)std::max(H, CB));
  const size_t o_codes = take(sizeof(int32_t) * rowsQ * (size_t)NQ);
  if (off > m.arena_cap) {
    ENC_CUDA_OK(cudaStreamSynchronize(s));
    if (m.arena) cudaFree(m.arena);
    m.arena = nullptr; m.arena_cap = 0; m.zeroed_B = -1;
    size_t free_b = 0, total_b = 0;
    ENC_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    if (off > free_b) throw Error(Q3TTS_ENOMEM, "encode: the batch needs " + std::to_string(off >> 20) + " MiB of device memory; encode it in smaller batches");
    if (cudaMalloc(&m.arena, off) != cudaSuccess) { cudaGetLastError(); throw Error(Q3TTS_ENOMEM, "encode: device allocation failed"); }
    m.arena_cap = off;
  }
  char* base = (char*)m.arena;
  auto F = [&](size_t o) { return (float*)(base + o); };
  auto Hp = [&](size_t o) { return (__half*)(base + o); };
  zero_if_new_shape(m, base + o_audio, off - o_audio, B, samples, s);
  m.launches = 0;
  std::vector<int> hlen((size_t)B * (size_t)(nst + 2));
  for (int lv = 0; lv < nst + 2; ++lv)
    for (int b = 0; b < B; ++b) hlen[(size_t)lv * (size_t)B + (size_t)b] = (int)L[(size_t)lv];
  int* d_len = (int*)(base + o_len);
  ENC_CUDA_OK(cudaMemcpyAsync(d_len, hlen.data(), hlen.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  ENC_CUDA_OK(cudaMemcpyAsync(F(o_audio), audio, sizeof(float) * (size_t)B * (size_t)samples, cudaMemcpyHostToDevice, s));
  auto geom = [&](int64_t slot_rows, int level) {
    BatchGeom g{};
    g.B = B; g.Tmax = (int)slot_rows; g.len_frames = d_len + (size_t)level * (size_t)B; g.valid_frames = (long long)B * L[(size_t)level]; g.row_begin = nullptr;
    return g;
  };
  // element-wise pass: v = y / y_scale (+ res, scaled); x_out = v; the operand act(v) in the form the consuming GEMM takes
  auto split = [&](const BatchGeom& g, int C, const float* y, float y_scale, const float* res, const float* scale, int act, float* x_out,
                   const EncGemm* consumer, float* a32, __half* h3, bool tap_a32 = false) {
    const bool tcg = consumer && tc_takes(m, *consumer);
    const bool want32 = (consumer && !tcg) || (tap_a32 && m.taps_enabled);     // (a stage tap reads the float32 form)
    const int grp = consumer ? consumer->Cin / C : 1;                          // the consumer's stride: its Cin is stride x C
    launch_enc_split(y, 1.0f / y_scale, res, scale, act, x_out, want32 ? a32 : nullptr, tcg ? h3 : nullptr, grp, C, g, s);
    ++m.launches;
  };
  m.tap_index.clear();

  // ---- Seanet ----
  {  // init conv on the CUDA cores (1 input channel); its elu output lands in Y[0], split for stage 0's first conv if that runs on tensor cores
    const int nf = c.num_filters;
    const bool tcg = tc_takes(m, m.stages[0].res3);
    launch_enc_init_conv(F(o_audio), samples, m.init_w, m.init_b, c.kernel_size, nf, F(oX[0]), tcg ? nullptr : F(oY[0]), tcg ? Hp(oA3[0]) : nullptr,
                         P[0] * (int64_t)nf, geom(P[0], 0), s);
    ++m.launches;
  }
  for (int li = 0; li < nst; ++li) {
    const EncStage& st = m.stages[(size_t)li];
    const int dim = st.dim, hid = dim / c.compress;
    const int64_t Pl = P[(size_t)li], Pn = P[(size_t)li + 1];
    const BatchGeom g = geom(Pl, li);
    const Opnd a{F(oY[(size_t)li]), Hp(oA3[(size_t)li])};
    const Opnd h{F(oHY[(size_t)li]), Hp(oH3[(size_t)li])};
    float sc = gemm_y(m, st.res3, g, a, dim, Pl * dim, F(oHY[(size_t)li]), hid, Pl * hid);
    split(g, hid, F(oHY[(size_t)li]), sc, nullptr, nullptr, 1, nullptr, &st.res1, F(oHY[(size_t)li]), h.h3, true);
    m.tap_index["hid" + std::to_string(li)] = EncTap{oHY[(size_t)li], Pl, L[(size_t)li], hid};
    gemm_y(m, st.res1, g, h, hid, Pl * hid, F(oX[(size_t)li]), dim, Pl * dim, F(oX[(size_t)li]));     // x += conv1(.) in the GEMM's epilogue
    const Opnd d{F(oY[(size_t)li]), Hp(oD3[(size_t)li])};
    split(g, dim, F(oX[(size_t)li]), 1.0f, nullptr, nullptr, 1, nullptr, &st.down, d.f32, d.h3);
    m.tap_index["res" + std::to_string(li)] = EncTap{oX[(size_t)li], Pl, L[(size_t)li], dim};
    sc = gemm_y(m, st.down, geom(L[(size_t)li + 1], li + 1), d, st.ratio * dim, Pl * dim, F(oX[(size_t)li + 1]), 2 * dim, Pn * 2 * dim);
    const EncGemm* next = li + 1 < nst ? &m.stages[(size_t)li + 1].res3 : &m.final_conv;
    split(geom(Pn, li + 1), 2 * dim, F(oX[(size_t)li + 1]), sc, nullptr, nullptr, 1, F(oX[(size_t)li + 1]), next, F(oY[(size_t)li + 1]), Hp(oA3[(size_t)li + 1]));
    if (li == nst - 1) m.tap_index["layer" + std::to_string(li)] = EncTap{oX[(size_t)li + 1], Pn, L[(size_t)li + 1], 2 * dim};
  }
  const int64_t PT = P[(size_t)nst], LT = L[(size_t)nst];
  const BatchGeom gT = geom(PT, nst);
  {
    const int dimL = 2 * m.stages[(size_t)nst - 1].dim;
    const float sc = gemm_y(m, m.final_conv, gT, Opnd{F(oY[(size_t)nst]), Hp(oA3[(size_t)nst])}, dimL, PT * dimL, F(o_hs), H, PT * H);
    if (sc != 1.0f) split(gT, H, F(o_hs), sc, nullptr, nullptr, 0, F(o_hs), nullptr, nullptr, nullptr);
  }
  if (m.taps_enabled) {
    ENC_CUDA_OK(cudaMemcpyAsync(F(o_tap), F(o_hs), 4 * rowsT * (size_t)H, cudaMemcpyDeviceToDevice, s));
    m.tap_index["seanet"] = EncTap{o_tap, PT, LT, H};
  }

  // ---- transformer ----
  const float scale = 1.0f / std::sqrt((float)hd);
  const Opnd tmp{F(o_yt), Hp(o_t3)};                              // operand scratch of the layer in flight
  for (const EncLayer& Ly : m.layers) {
    launch_layernorm(F(o_hs), Ly.n1w, Ly.n1b, 1e-5f, F(o_nb), gT, H, s); ++m.launches;
    if (tc_takes(m, Ly.qkv)) split(gT, H, F(o_nb), 1.0f, nullptr, nullptr, 0, nullptr, &Ly.qkv, nullptr, tmp.h3);
    const float sq = gemm_y(m, Ly.qkv, gT, Opnd{F(o_nb), tmp.h3}, H, PT * H, F(o_qkv), QW, PT * QW);
    // q, k, v all carry the factor sq: RoPE is linear, the scores carry sq^2 (folded into the softmax scale), the output carries sq
    launch_rope(F(o_qkv), QW, nh + nkv, hd, m.inv_freq, gT, s); ++m.launches;
    launch_attention(F(o_qkv), DT_F32, F(o_ao), DT_F32, gT, nh, nkv, hd, scale / (sq * sq), (int)PT + 1, s); ++m.launches;
    split(gT, H, F(o_ao), sq, nullptr, nullptr, 0, nullptr, &Ly.o, F(o_ao), tmp.h3);
    gemm_y(m, Ly.o, gT, Opnd{F(o_ao), tmp.h3}, H, PT * H, F(o_hs), H, PT * H, F(o_hs), Ly.ls1_s, Ly.ls1);   // h += ls1 * attn, in the epilogue
    launch_layernorm(F(o_hs), Ly.n2w, Ly.n2b, 1e-5f, F(o_nb), gT, H, s); ++m.launches;
    if (tc_takes(m, Ly.fc1)) split(gT, H, F(o_nb), 1.0f, nullptr, nullptr, 0, nullptr, &Ly.fc1, nullptr, tmp.h3);
    const float sc = gemm_y(m, Ly.fc1, gT, Opnd{F(o_nb), tmp.h3}, H, PT * H, F(o_yt), I, PT * I);
    split(gT, I, F(o_yt), sc, nullptr, nullptr, 2, nullptr, &Ly.fc2, F(o_yt), tmp.h3);               // tanh-GELU, in place / split
    gemm_y(m, Ly.fc2, gT, tmp, I, PT * I, F(o_hs), H, PT * H, F(o_hs), Ly.ls2_s, Ly.ls2);            // h += ls2 * mlp, in the epilogue
  }
  m.tap_index["transformer"] = EncTap{o_hs, PT, LT, H};

  // ---- downsample + quantizer ----
  const BatchGeom gQ = geom(Tq, nst + 1);
  if (tc_takes(m, m.downsample)) split(gT, H, F(o_hs), 1.0f, nullptr, nullptr, 0, nullptr, &m.downsample, nullptr, Hp(o_ds3));
  {
    const float sc = gemm_y(m, m.downsample, gQ, Opnd{F(o_hs), Hp(o_ds3)}, ds * H, PT * H, F(o_d), H, Tq * H);
    // true scale in place (the stage tap and the CUDA-core fallback read D), and the split operand of the two input projections
    split(gQ, H, F(o_d), sc, nullptr, nullptr, 0, F(o_d), &m.proj[0], F(o_d), Hp(o_q3));
  }
  m.tap_index["downsample"] = EncTap{o_d, Tq, Tq, H};
  const size_t o_res[2] = {o_r0, o_r1};
  for (int part = 0; part < 2; ++part) {
    const float sc = gemm_y(m, m.proj[part], gQ, Opnd{F(o_d), Hp(o_q3)}, H, Tq * H, F(o_res[part]), CB, Tq * CB);
    if (sc != 1.0f) split(gQ, CB, F(o_res[part]), sc, nullptr, nullptr, 0, F(o_res[part]), nullptr, nullptr, nullptr);   // the residual is updated with unscaled code vectors
  }
  int32_t* d_codes = (int32_t*)(base + o_codes);
  for (size_t q = 0; q < m.books.size(); ++q) {
    const EncBook& bk = m.books[q];
    float* R = F(o_res[bk.part]);
    if (tc_takes(m, bk.score)) split(gQ, CB, R, 1.0f, nullptr, nullptr, 0, nullptr, &bk.score, nullptr, Hp(o_q3));
    gemm_y(m, bk.score, gQ, Opnd{R, Hp(o_q3)}, CB, Tq * CB, F(o_sc), K, Tq * K);                     // the argmax does not see the common factor
    launch_vq_select(F(o_sc), K, bk.score.w, CB, R, d_codes + (int64_t)q * Tq, (int64_t)NQ * Tq, gQ, s);
    ++m.launches;
  }
  ENC_CUDA_OK(cudaMemcpyAsync(codes_out, d_codes, sizeof(int32_t) * rowsQ * (size_t)NQ, cudaMemcpyDeviceToHost, s));
  ENC_CUDA_OK(cudaStreamSynchronize(s));
  ENC_CUDA_OK(cudaGetLastError());
  m.last_B = B;
}
}  // namespace

void encoder_encode(EncoderModel& m, const float* audio, int B, int64_t samples, int32_t* codes_out) {
  std::lock_guard<std::mutex> lock(m.mu);
  if (m.tc) {
    if (B < 1 || samples < 1 || !audio || !codes_out) throw Error(Q3TTS_EINVAL, "encode: bad arguments");
    if (samples > (int64_t)1 << 26 || B > 65535) throw Error(Q3TTS_EINVAL, "encode: at most 2^26 samples (46 minutes) per utterance and 65535 utterances per call");
    encode_tc(m, audio, B, samples, codes_out);
    return;
  }
  const EncoderConfig& c = m.cfg;
  if (B < 1 || samples < 1 || !audio || !codes_out) throw Error(Q3TTS_EINVAL, "encode: bad arguments");
  if (samples > (int64_t)1 << 26 || B > 65535) throw Error(Q3TTS_EINVAL, "encode: at most 2^26 samples (46 minutes) per utterance and 65535 utterances per call");
  ENC_CUDA_OK(cudaSetDevice(m.device));
  cudaStream_t s = m.stream;
  const int nst = c.n_ratios, ds = c.downsample_stride(), H = c.hidden_size, I = c.intermediate_size;
  const int nh = c.num_attention_heads, nkv = c.num_key_value_heads, hd = H / nh, QW = (nh + 2 * nkv) * hd;
  const int K = c.codebook_size, CB = c.codebook_dim, NQ = c.valid_quantizers;
  // rows per utterance at every level: L valid, P allocated (the next strided conv views [P / r, r * C]; rows L..P-1 stay zero,
  // which is the reference's right-hand "extra padding")
  std::vector<int64_t> L((size_t)nst + 2), P((size_t)nst + 1);
  L[0] = samples;
  for (int li = 0; li < nst; ++li) {
    const int r = m.stages[(size_t)li].ratio;
    L[(size_t)li + 1] = (L[(size_t)li] + r - 1) / r;
    P[(size_t)li] = L[(size_t)li + 1] * r;
  }
  const int64_t Tq = (L[(size_t)nst] + ds - 1) / ds;
  L[(size_t)nst + 1] = Tq;
  P[(size_t)nst] = Tq * ds;
  if (P[0] * (int64_t)m.stages[0].dim > (int64_t)1 << 40) throw Error(Q3TTS_EINVAL, "encode: batch too large");

  // ---- arena plan (bytes, 256-aligned) ----
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_len = take(sizeof(int) * (size_t)B * (size_t)(nst + 2));
  const size_t o_audio = take(sizeof(float) * (size_t)B * (size_t)samples);
  std::vector<size_t> oX((size_t)nst + 1), oA((size_t)nst + 1), oH((size_t)nst);
  for (int li = 0; li <= nst; ++li) {
    const int dim = li < nst ? m.stages[(size_t)li].dim : 2 * m.stages[(size_t)nst - 1].dim;
    oX[(size_t)li] = take(sizeof(float) * (size_t)B * (size_t)P[(size_t)li] * (size_t)dim);
    oA[(size_t)li] = take(sizeof(float) * (size_t)B * (size_t)P[(size_t)li] * (size_t)dim);
    if (li < nst) oH[(size_t)li] = take(sizeof(float) * (size_t)B * (size_t)P[(size_t)li] * (size_t)(dim / c.compress));
  }
  const size_t rowsT = (size_t)B * (size_t)P[(size_t)nst];
  const size_t o_hs = take(sizeof(float) * rowsT * (size_t)H), o_nb = take(sizeof(float) * rowsT * (size_t)H);
  const size_t o_qkv = take(sizeof(float) * rowsT * (size_t)QW), o_ao = take(sizeof(float) * rowsT * (size_t)H);
  const size_t o_ff = take(sizeof(float) * rowsT * (size_t)I);
  const size_t o_tap = take(m.taps_enabled ? sizeof(float) * rowsT * (size_t)H : 0);
  const size_t rowsQ = (size_t)B * (size_t)Tq;
  const size_t o_d = take(sizeof(float) * rowsQ * (size_t)H);
  const size_t o_r0 = take(sizeof(float) * rowsQ * (size_t)CB), o_r1 = take(sizeof(float) * rowsQ * (size_t)CB);
  const size_t o_sc = take(sizeof(float) * rowsQ * (size_t)K);
  const size_t o_codes = take(sizeof(int32_t) * rowsQ * (size_t)NQ);
  if (off > m.arena_cap) {
    ENC_CUDA_OK(cudaStreamSynchronize(s));
    if (m.arena) cudaFree(m.arena);
    m.arena = nullptr; m.arena_cap = 0; m.zeroed_B = -1;
    size_t free_b = 0, total_b = 0;
    ENC_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    if (off > free_b) throw Error(Q3TTS_ENOMEM, "encode: the batch needs " + std::to_string(off >> 20) + " MiB of device memory; encode it in smaller batches");
    if (cudaMalloc(&m.arena, off) != cudaSuccess) { cudaGetLastError(); throw Error(Q3TTS_ENOMEM, "encode: device allocation failed"); }
    m.arena_cap = off;
  }
  char* base = (char*)m.arena;
  auto F = [&](size_t o) { return (float*)(base + o); };
  zero_if_new_shape(m, base + o_audio, off - o_audio, B, samples, s);
  m.launches = 0;

  // ---- lengths + audio to the device ----
  std::vector<int> hlen((size_t)B * (size_t)(nst + 2));
  for (int lv = 0; lv < nst + 2; ++lv)
    for (int b = 0; b < B; ++b) hlen[(size_t)lv * (size_t)B + (size_t)b] = (int)L[(size_t)lv];
  int* d_len = (int*)(base + o_len);
  ENC_CUDA_OK(cudaMemcpyAsync(d_len, hlen.data(), hlen.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  ENC_CUDA_OK(cudaMemcpyAsync(F(o_audio), audio, sizeof(float) * (size_t)B * (size_t)samples, cudaMemcpyHostToDevice, s));
  auto geom = [&](int64_t slot_rows, int level) {
    BatchGeom g{};
    g.B = B; g.Tmax = (int)slot_rows; g.len_frames = d_len + (size_t)level * (size_t)B; g.valid_frames = (long long)B * L[(size_t)level]; g.row_begin = nullptr;
    return g;
  };

  // ---- Seanet (STE.swift:436-443) ----
  launch_enc_init_conv(F(o_audio), samples, m.init_w, m.init_b, c.kernel_size, c.num_filters, F(oX[0]), F(oA[0]), nullptr,
                       P[0] * (int64_t)c.num_filters, geom(P[0], 0), s);
  ++m.launches;
  m.tap_index.clear();
  for (int li = 0; li < nst; ++li) {
    const EncStage& st = m.stages[(size_t)li];
    const int dim = st.dim, hid = dim / c.compress;
    const int64_t Pl = P[(size_t)li];
    const BatchGeom g = geom(Pl, li);
    {  // block.0: conv k3 on elu(x) -> elu(.) (STE.swift:337-340)
      ConvGemmParams e{};
      e.out_a = F(oH[(size_t)li]); e.lda_out = hid; e.ao_bstride = Pl * hid; e.a_elu = 1;
      run_gemm(m, st.res3, g, F(oA[(size_t)li]), dim, Pl * dim, e);
    }
    {  // block.1: conv k1, + x (true skip, STE.swift:345-348); the stage's downsample reads elu(x') (STE.swift:389)
      ConvGemmParams e{};
      e.res = F(oX[(size_t)li]); e.ldres = dim; e.res_bstride = Pl * dim;
      e.out_y = F(oX[(size_t)li]); e.ldy = dim; e.y_bstride = Pl * dim;
      e.out_a = F(oA[(size_t)li]); e.lda_out = dim; e.ao_bstride = Pl * dim; e.a_elu = 1;
      run_gemm(m, st.res1, g, F(oH[(size_t)li]), hid, Pl * hid, e);
    }
    {  // downsample: k = 2r, stride r, as a 2-tap GEMM over [frames, r * dim]
      const int64_t Pn = P[(size_t)li + 1];
      ConvGemmParams e{};
      e.out_y = F(oX[(size_t)li + 1]); e.ldy = 2 * dim; e.y_bstride = Pn * 2 * dim;
      e.out_a = F(oA[(size_t)li + 1]); e.lda_out = 2 * dim; e.ao_bstride = Pn * 2 * dim; e.a_elu = 1;
      run_gemm(m, st.down, geom(L[(size_t)li + 1], li + 1), F(oA[(size_t)li]), st.ratio * dim, Pl * dim, e);
      if (li == nst - 1) m.tap_index["layer" + std::to_string(li)] = EncTap{oX[(size_t)li + 1], Pn, L[(size_t)li + 1], 2 * dim};   // (earlier stages' outputs are updated in place by the next residual block)
      m.tap_index["res" + std::to_string(li)] = EncTap{oX[(size_t)li], Pl, L[(size_t)li], dim};          // x + block(x), the stage's stream
      m.tap_index["hid" + std::to_string(li)] = EncTap{oH[(size_t)li], Pl, L[(size_t)li], hid};          // elu(conv3(elu(x)))
    }
  }
  const int64_t PT = P[(size_t)nst], LT = L[(size_t)nst];
  const BatchGeom gT = geom(PT, nst);
  {  // final conv on elu(x) (STE.swift:441-442) -> the transformer's stream
    const int dimL = 2 * m.stages[(size_t)nst - 1].dim;
    ConvGemmParams e{};
    e.out_y = F(o_hs); e.ldy = H; e.y_bstride = PT * H;
    run_gemm(m, m.final_conv, gT, F(oA[(size_t)nst]), dimL, PT * dimL, e);
  }
  if (m.taps_enabled) {
    ENC_CUDA_OK(cudaMemcpyAsync(F(o_tap), F(o_hs), sizeof(float) * rowsT * (size_t)H, cudaMemcpyDeviceToDevice, s));
    m.tap_index["seanet"] = EncTap{o_tap, PT, LT, H};
  }

  // ---- transformer (STE.swift:571-590; full causal mask STE.swift:1038-1042; no input / output projection at 512 == d_model) ----
  const float scale = 1.0f / std::sqrt((float)hd);
  for (const EncLayer& Ly : m.layers) {
    launch_layernorm(F(o_hs), Ly.n1w, Ly.n1b, 1e-5f, F(o_nb), gT, H, s); ++m.launches;
    { ConvGemmParams e{}; e.out_y = F(o_qkv); e.ldy = QW; e.y_bstride = PT * QW; run_gemm(m, Ly.qkv, gT, F(o_nb), H, PT * H, e); }
    launch_rope(F(o_qkv), QW, nh + nkv, hd, m.inv_freq, gT, s); ++m.launches;
    launch_attention(F(o_qkv), DT_F32, F(o_ao), DT_F32, gT, nh, nkv, hd, scale, (int)PT + 1, s); ++m.launches;
    {
      ConvGemmParams e{};
      e.res = F(o_hs); e.ldres = H; e.res_bstride = PT * H; e.scale = Ly.ls1;
      e.out_y = F(o_hs); e.ldy = H; e.y_bstride = PT * H;
      run_gemm(m, Ly.o, gT, F(o_ao), H, PT * H, e);
    }
    launch_layernorm(F(o_hs), Ly.n2w, Ly.n2b, 1e-5f, F(o_nb), gT, H, s); ++m.launches;
    { ConvGemmParams e{}; e.act = ACT_GELU_TANH; e.out_y = F(o_ff); e.ldy = I; e.y_bstride = PT * I; run_gemm(m, Ly.fc1, gT, F(o_nb), H, PT * H, e); }
    {
      ConvGemmParams e{};
      e.res = F(o_hs); e.ldres = H; e.res_bstride = PT * H; e.scale = Ly.ls2;
      e.out_y = F(o_hs); e.ldy = H; e.y_bstride = PT * H;
      run_gemm(m, Ly.fc2, gT, F(o_ff), I, PT * I, e);
    }
  }
  m.tap_index["transformer"] = EncTap{o_hs, PT, LT, H};

  // ---- downsample to the code rate (STE.swift:684-705), then the split residual quantizer (STE.swift:816-829, 934-941) ----
  const BatchGeom gQ = geom(Tq, nst + 1);
  { ConvGemmParams e{}; e.out_y = F(o_d); e.ldy = H; e.y_bstride = Tq * H; run_gemm(m, m.downsample, gQ, F(o_hs), ds * H, PT * H, e); }
  m.tap_index["downsample"] = EncTap{o_d, Tq, Tq, H};
  const size_t o_res[2] = {o_r0, o_r1};
  for (int part = 0; part < 2; ++part) {
    ConvGemmParams e{};
    e.out_y = F(o_res[part]); e.ldy = CB; e.y_bstride = Tq * CB;
    run_gemm(m, m.proj[part], gQ, F(o_d), H, Tq * H, e);
  }
  int32_t* d_codes = (int32_t*)(base + o_codes);
  for (size_t q = 0; q < m.books.size(); ++q) {
    const EncBook& bk = m.books[q];
    { ConvGemmParams e{}; e.out_y = F(o_sc); e.ldy = K; e.y_bstride = Tq * K; run_gemm(m, bk.score, gQ, F(o_res[bk.part]), CB, Tq * CB, e); }
    launch_vq_select(F(o_sc), K, bk.score.w, CB, F(o_res[bk.part]), d_codes + (int64_t)q * Tq, (int64_t)NQ * Tq, gQ, s);
    ++m.launches;
  }
  ENC_CUDA_OK(cudaMemcpyAsync(codes_out, d_codes, sizeof(int32_t) * rowsQ * (size_t)NQ, cudaMemcpyDeviceToHost, s));
  ENC_CUDA_OK(cudaStreamSynchronize(s));
  ENC_CUDA_OK(cudaGetLastError());
  m.last_B = B;
}

void encoder_tap(EncoderModel& m, const std::string& name, float* out, int64_t cap, int64_t dims[3]) {
  std::lock_guard<std::mutex> lock(m.mu);
  auto it = m.tap_index.find(name);
  if (it == m.tap_index.end() || !m.arena || m.last_B < 1) throw Error(Q3TTS_EINVAL, "encoder tap '" + name + "' is not available (enable taps, then encode)");
  const EncTap& t = it->second;
  dims[0] = m.last_B; dims[1] = t.valid_rows; dims[2] = t.C;
  if (!out) return;
  if (cap < dims[0] * dims[1] * dims[2]) throw Error(Q3TTS_EINVAL, "encoder tap: output buffer too small");
  ENC_CUDA_OK(cudaSetDevice(m.device));
  ENC_CUDA_OK(cudaMemcpy2DAsync(out, sizeof(float) * (size_t)t.valid_rows * (size_t)t.C, (const char*)m.arena + t.offset,
                                sizeof(float) * (size_t)t.slot_rows * (size_t)t.C, sizeof(float) * (size_t)t.valid_rows * (size_t)t.C,
                                (size_t)m.last_B, cudaMemcpyDeviceToHost, m.stream));
  ENC_CUDA_OK(cudaStreamSynchronize(m.stream));
}

}  // namespace q3
