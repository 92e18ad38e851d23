// See engine.hpp.  Weight packing + the per-micro-batch launch chain.
#include "engine.hpp"

#include <algorithm>
#include <functional>
#include <cmath>
#include <cstring>
#include <memory>

namespace q3 {

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw Error(_e == cudaErrorMemoryAllocation ? Q3TTS_ENOMEM : Q3TTS_ECUDA,                         \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                  \
  } while (0)

static size_t dt_size(int dt) { return dt == DT_F32 ? 4 : 2; }

Model::~Model() {
  cudaSetDevice(device);
  for (void* p : allocs) cudaFree(p);
  if (arena) cudaFree(arena);
  if (d_codes) cudaFree(d_codes);
  if (d_pcm) cudaFree(d_pcm);
  for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (stream_hook_h) cudaFreeHost(stream_hook_h);
  if (stream_hook_d) cudaFree(stream_hook_d);
  if (d_lengths) cudaFree(d_lengths);
  if (d_meta) cudaFree(d_meta);
  if (h_meta) cudaFreeHost(h_meta);
  if (stream_ws) cudaFree(stream_ws);
  if (stream_meta_h) cudaFreeHost(stream_meta_h);
  if (d_err) cudaFree(d_err);
  if (h_err) cudaFreeHost(h_err);
  for (auto& t : taps) if (t.second.d) cudaFree(t.second.d);
  for (auto& p : prof) { if (p.ev0) cudaEventDestroy(p.ev0); if (p.ev1) cudaEventDestroy(p.ev1); }
  for (auto& lp : launch_prof) { if (lp.ev0) cudaEventDestroy(lp.ev0); if (lp.ev1) cudaEventDestroy(lp.ev1); }
  for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
  if (chain_done) cudaEventDestroy(chain_done);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (mb_done[i]) cudaEventDestroy(mb_done[i]);
    if (d2h_done[i]) cudaEventDestroy(d2h_done[i]);
    if (h_pcm[i]) cudaFreeHost(h_pcm[i]);
  }
  if (h_codes) cudaFreeHost(h_codes);
  if (h_len) cudaFreeHost(h_len);
  if (stream) cudaStreamDestroy(stream);
}

void chain_begin(Model& m, cudaStream_t s) {
  if (m.has_chain && m.last_stream != s) CUDA_OK(cudaStreamWaitEvent(s, m.chain_done, 0));
}
void chain_end(Model& m, cudaStream_t s) {
  CUDA_OK(cudaEventRecord(m.chain_done, s));
  m.last_stream = s;
  m.has_chain = true;
}

// ---- upload helpers -----------------------------------------------------------------------------
static float* upload(Model& m, const std::vector<float>& v) {
  float* d = nullptr;
  CUDA_OK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
  m.allocs.push_back(d);
  // Stream-ordered with the fp32->16-bit convert kernels that follow on m.stream.  (A plain cudaMemcpy from pageable
  // memory may return before the DMA lands, and m.stream is non-blocking: the convert could read stale bytes.)
  CUDA_OK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, m.stream));
  CUDA_OK(cudaStreamSynchronize(m.stream));
  return d;
}

static const HostTensor& T(const Checkpoint& ck, const std::string& k) {
  auto it = ck.tensors.find(k);
  if (it == ck.tensors.end()) throw Error(Q3TTS_EFORMAT, "missing decoder tensor " + k);
  return it->second;
}

static void finish_gemm(Model& m, GemmW& g, const std::vector<float>& packed, const std::vector<float>* bias) {
  if (g.Cin % 4) throw Error(Q3TTS_EFORMAT, "channel counts must be multiples of 4");
  g.w32 = upload(m, packed);
  if (m.op_dtype != DT_F32) {
    void* d = nullptr;
    CUDA_OK(cudaMalloc(&d, packed.size() * 2));
    m.allocs.push_back(d);
    launch_convert(g.w32, d, m.op_dtype, (int64_t)packed.size(), m.stream);
    g.w16 = d;
  }
  g.bias = bias ? upload(m, *bias) : nullptr;
}

// conv weight, MLX layout [Cout, K, Cin] -> [K][Cout][Cin]
static void pack_conv(Model& m, GemmW& g, const HostTensor& w, const HostTensor* bias, int dil) {
  const int64_t Co = w.shape[0], K = w.shape[1], Ci = w.shape[2];
  g.taps = (int)K; g.N = (int)Co; g.Cin = (int)Ci; g.dil = dil;
  std::vector<float> p((size_t)(K * Co * Ci));
  for (int64_t o = 0; o < Co; ++o)
    for (int64_t k = 0; k < K; ++k)
      std::memcpy(&p[(size_t)((k * Co + o) * Ci)], &w.data[(size_t)((o * K + k) * Ci)], (size_t)Ci * 4);
  finish_gemm(m, g, p, bias ? &bias->data : nullptr);
}

// linear weight [N, Cin] (several stacked row-wise) -> taps=1
static void pack_linear(Model& m, GemmW& g, const std::vector<const HostTensor*>& ws, const HostTensor* bias,
                        bool interleave2 = false) {
  const int64_t Ci = ws[0]->shape[1];
  int64_t N = 0;
  for (auto* w : ws) N += w->shape[0];
  g.taps = 1; g.N = (int)N; g.Cin = (int)Ci; g.dil = 1;
  std::vector<float> p((size_t)(N * Ci));
  if (interleave2) {  // rows (2i, 2i+1) = (ws[0][i], ws[1][i]) so the SwiGLU epilogue sees gate/up side by side
    const int64_t I = ws[0]->shape[0];
    for (int64_t i = 0; i < I; ++i) {
      std::memcpy(&p[(size_t)((2 * i) * Ci)], &ws[0]->data[(size_t)(i * Ci)], (size_t)Ci * 4);
      std::memcpy(&p[(size_t)((2 * i + 1) * Ci)], &ws[1]->data[(size_t)(i * Ci)], (size_t)Ci * 4);
    }
  } else {
    int64_t r = 0;
    for (auto* w : ws) {
      std::memcpy(&p[(size_t)(r * Ci)], w->data.data(), w->data.size() * 4);
      r += w->shape[0];
    }
  }
  finish_gemm(m, g, p, bias ? &bias->data : nullptr);
}

// transposed conv, MLX layout [Cout, K, Cin], stride r.  K == r: one tap; K == 2r: two taps
// (ST.swift:339-353: y[t*r+p] = W[:,p,:] x[t] + W[:,p+r,:] x[t-1] + b).
static void pack_tconv(Model& m, GemmW& g, const HostTensor& w, const HostTensor& bias, int r) {
  const int64_t Co = w.shape[0], K = w.shape[1], Ci = w.shape[2];
  if (K != r && K != 2 * r) throw Error(Q3TTS_EFORMAT, "transposed conv kernel must be stride or 2*stride");
  const int taps = (K == r) ? 1 : 2;
  g.taps = taps; g.N = (int)(r * Co); g.Cin = (int)Ci; g.dil = 1;
  std::vector<float> p((size_t)(taps * r * Co * Ci));
  for (int j = 0; j < taps; ++j)
    for (int64_t ph = 0; ph < r; ++ph)
      for (int64_t o = 0; o < Co; ++o) {
        // last tap (j = taps-1) multiplies x[t] -> kernel index ph; the earlier one multiplies x[t-1] -> ph + r
        const int64_t k = (j == taps - 1) ? ph : ph + r;
        std::memcpy(&p[(size_t)(((int64_t)j * r * Co + ph * Co + o) * Ci)], &w.data[(size_t)((o * K + k) * Ci)], (size_t)Ci * 4);
      }
  std::vector<float> b((size_t)(r * Co));
  for (int64_t ph = 0; ph < r; ++ph)
    for (int64_t o = 0; o < Co; ++o) b[(size_t)(ph * Co + o)] = bias.data[(size_t)o];
  finish_gemm(m, g, p, &b);
}

static SnakeW pack_snake(Model& m, const HostTensor& alpha, const HostTensor& beta, int rep) {
  const int64_t C = alpha.shape[0];
  std::vector<float> ea((size_t)(C * rep)), ib((size_t)(C * rep));
  for (int r = 0; r < rep; ++r)
    for (int64_t c = 0; c < C; ++c) {
      ea[(size_t)(r * C + c)] = expf(alpha.data[(size_t)c]);
      ib[(size_t)(r * C + c)] = 1.0f / (expf(beta.data[(size_t)c]) + 1e-9f);   // ST.swift:237, 252
    }
  SnakeW s;
  s.ea = upload(m, ea);
  s.ib = upload(m, ib);
  s.n = (int)(C * rep);
  return s;
}

Model* model_create(const Checkpoint& ck, const q3tts_options& opts) {
  std::unique_ptr<Model> mp(new Model());
  Model& m = *mp;
  m.cfg = ck.cfg;
  m.opts = opts;
  int dev = opts.device;
  if (dev < 0) CUDA_OK(cudaGetDevice(&dev));
  m.device = dev;
  CUDA_OK(cudaSetDevice(dev));
  cudaDeviceProp prop{};
  CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    throw Error(Q3TTS_ECUDA, std::string("device ") + prop.name + " is not sm_100: libqwen3tts_cuda is built for sm_100a only and has no fallback");
  m.op_dtype = opts.precision == Q3TTS_PREC_FP32 ? DT_F32 : (opts.precision == Q3TTS_PREC_FP16 ? DT_F16 : DT_BF16);
  m.st_dtype = m.op_dtype;
  if (const char* e = getenv("Q3TTS_STREAM_F32")) { if (e[0] == '1') m.st_dtype = DT_F32; }   // keep the blocks' residual stream in fp32
  CUDA_OK(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&m.chain_done, cudaEventDisableTiming));
  const q3tts_config& c = m.cfg;
  if (c.latent_dim > 2048) throw Error(Q3TTS_EFORMAT, "latent_dim > 2048 is not supported");
  if (c.head_dim != 32 && c.head_dim != 64 && c.head_dim != 128) throw Error(Q3TTS_EFORMAT, "head_dim must be 32, 64 or 128");
  if ((c.codebook_dim / 2) % 4) throw Error(Q3TTS_EFORMAT, "codebook_dim/2 must be a multiple of 4");
  for (auto& kv : ck.tensors) m.weight_shapes[kv.first] = kv.second.shape;

  // --- codebooks (fp32 tables; Q3.swift:1716-1724 folding was done by the loader) ---
  std::vector<const float*> tabs;
  std::vector<int32_t> sizes;
  for (int q = 0; q < c.num_quantizers; ++q) {
    const bool sem = q < c.num_semantic_quantizers;
    const std::string k = std::string("decoder.quantizer.") + (sem ? "rvq_first" : "rvq_rest") + ".vq.layers." +
                          std::to_string(sem ? q : q - c.num_semantic_quantizers) + ".codebook.embed.weight";
    const HostTensor& e = T(ck, k);
    float* d = upload(m, e.data);
    m.codebooks.push_back(d);
    tabs.push_back(d);
    sizes.push_back((int32_t)e.shape[0]);
  }
  CUDA_OK(cudaMalloc(&m.d_tables, tabs.size() * sizeof(float*)));
  m.allocs.push_back((void*)m.d_tables);
  CUDA_OK(cudaMemcpyAsync((void*)m.d_tables, tabs.data(), tabs.size() * sizeof(float*), cudaMemcpyHostToDevice, m.stream));
  CUDA_OK(cudaMalloc(&m.d_table_sizes, sizes.size() * 4));
  m.allocs.push_back(m.d_table_sizes);
  CUDA_OK(cudaMemcpyAsync(m.d_table_sizes, sizes.data(), sizes.size() * 4, cudaMemcpyHostToDevice, m.stream));
  CUDA_OK(cudaStreamSynchronize(m.stream));

  // --- RVQ output projections as ONE GEMM over [sum_first | sum_rest] (ST.swift:161-169, 214-226) ---
  {
    const HostTensor& p1 = T(ck, "decoder.quantizer.rvq_first.output_proj.weight");   // [dim,1,half]
    const HostTensor& p2 = T(ck, "decoder.quantizer.rvq_rest.output_proj.weight");
    const int64_t dim = p1.shape[0], half = p1.shape[2];
    std::vector<float> p((size_t)(dim * 2 * half));
    for (int64_t n = 0; n < dim; ++n) {
      std::memcpy(&p[(size_t)(n * 2 * half)], &p1.data[(size_t)(n * half)], (size_t)half * 4);
      std::memcpy(&p[(size_t)(n * 2 * half + half)], &p2.data[(size_t)(n * half)], (size_t)half * 4);
    }
    m.rvq_proj.taps = 1; m.rvq_proj.N = (int)dim; m.rvq_proj.Cin = (int)(2 * half); m.rvq_proj.dil = 1;
    finish_gemm(m, m.rvq_proj, p, nullptr);
  }
  pack_conv(m, m.pre_conv, T(ck, "decoder.pre_conv.conv.weight"), &T(ck, "decoder.pre_conv.conv.bias"), 1);
  const std::string pt = "decoder.pre_transformer.";
  pack_linear(m, m.in_proj, {&T(ck, pt + "input_proj.weight")}, &T(ck, pt + "input_proj.bias"));
  pack_linear(m, m.out_proj, {&T(ck, pt + "output_proj.weight")}, &T(ck, pt + "output_proj.bias"));
  m.final_norm = upload(m, T(ck, pt + "norm.weight").data);
  m.layers.resize((size_t)c.num_hidden_layers);
  for (int n = 0; n < c.num_hidden_layers; ++n) {
    const std::string p = pt + "layers." + std::to_string(n) + ".";
    LayerW& L = m.layers[(size_t)n];
    pack_linear(m, L.qkv, {&T(ck, p + "self_attn.q_proj.weight"), &T(ck, p + "self_attn.k_proj.weight"), &T(ck, p + "self_attn.v_proj.weight")}, nullptr);
    L.qkv.bias = upload(m, std::vector<float>((size_t)L.qkv.N, 0.f));   // bias-free in the reference (ST.swift:492-494); a zero bias selects the fast epilogue
    pack_linear(m, L.o, {&T(ck, p + "self_attn.o_proj.weight")}, nullptr);
    pack_linear(m, L.gate_up, {&T(ck, p + "mlp.gate_proj.weight"), &T(ck, p + "mlp.up_proj.weight")}, nullptr, true);
    pack_linear(m, L.down, {&T(ck, p + "mlp.down_proj.weight")}, nullptr);
    L.ln1 = upload(m, T(ck, p + "input_layernorm.weight").data);
    L.ln2 = upload(m, T(ck, p + "post_attention_layernorm.weight").data);
    L.ls_attn = upload(m, T(ck, p + "self_attn_layer_scale.scale").data);
    L.ls_mlp = upload(m, T(ck, p + "mlp_layer_scale.scale").data);
  }
  m.ups.resize((size_t)c.num_upsampling_ratios);
  for (int i = 0; i < c.num_upsampling_ratios; ++i) {
    const std::string u = "decoder.upsample." + std::to_string(i) + ".";
    UpsampleW& U = m.ups[(size_t)i];
    U.ratio = c.upsampling_ratios[i];
    pack_tconv(m, U.tconv, T(ck, u + "0.conv.weight"), T(ck, u + "0.conv.bias"), U.ratio);
    const HostTensor& dw = T(ck, u + "1.dwconv.conv.weight");   // [L,7,1]
    {  // [C][7] -> [7][C]: the row kernel reads four channels of one tap with one 16-byte load
      const int64_t Cc = dw.shape[0], K = dw.shape[1];
      std::vector<float> tw((size_t)(Cc * K));
      for (int64_t cc = 0; cc < Cc; ++cc)
        for (int64_t k = 0; k < K; ++k) tw[(size_t)(k * Cc + cc)] = dw.data[(size_t)(cc * K + k)];
      U.dw_w = upload(m, tw);
    }
    U.dw_b = upload(m, T(ck, u + "1.dwconv.conv.bias").data);
    U.ln_w = upload(m, T(ck, u + "1.norm.weight").data);
    U.ln_b = upload(m, T(ck, u + "1.norm.bias").data);
    pack_linear(m, U.pw1, {&T(ck, u + "1.pwconv1.weight")}, &T(ck, u + "1.pwconv1.bias"));
    pack_linear(m, U.pw2, {&T(ck, u + "1.pwconv2.weight")}, &T(ck, u + "1.pwconv2.bias"));
    U.gamma = upload(m, T(ck, u + "1.gamma").data);
  }
  const std::string dd = "decoder.decoder.";
  pack_conv(m, m.init_conv, T(ck, dd + "initConv.conv.weight"), &T(ck, dd + "initConv.conv.bias"), 1);
  static const int kDil[3] = {1, 3, 9};   // ST.swift:468-470
  for (int i = 0; i < 4; ++i) {
    const std::string b = dd + "block" + std::to_string(i) + ".";
    BlockW& B = m.blocks[i];
    B.rate = c.upsample_rates[i];
    B.cin = c.decoder_dim >> i;
    B.cout = c.decoder_dim >> (i + 1);
    m.block_in_snake[i] = pack_snake(m, T(ck, b + "snake.alpha"), T(ck, b + "snake.beta"), 1);
    pack_tconv(m, B.tconv, T(ck, b + "upsample.conv.weight"), T(ck, b + "upsample.conv.bias"), B.rate);
    for (int j = 0; j < 3; ++j) {
      const std::string r = b + "res" + std::to_string(j + 1) + ".";
      // act1 of res j is applied by the producer of the res unit's input: the transposed conv (j = 0,
      // replicated over the r phases) or the previous res unit's conv1 epilogue.
      B.act_in_next[j] = pack_snake(m, T(ck, r + "act1.alpha"), T(ck, r + "act1.beta"), j == 0 ? B.rate : 1);
      pack_conv(m, B.conv7[j], T(ck, r + "conv1.conv.weight"), &T(ck, r + "conv1.conv.bias"), kDil[j]);
      B.act2[j] = pack_snake(m, T(ck, r + "act2.alpha"), T(ck, r + "act2.beta"), 1);
      pack_conv(m, B.conv1[j], T(ck, r + "conv2.conv.weight"), &T(ck, r + "conv2.conv.bias"), 1);
    }
  }
  m.out_snake = pack_snake(m, T(ck, dd + "outSnake.alpha"), T(ck, dd + "outSnake.beta"), 1);
  {
    const HostTensor& w = T(ck, dd + "outConv.conv.weight");   // [1,7,C]
    m.tail_w = upload(m, w.data);
    m.tail_bias = T(ck, dd + "outConv.conv.bias").data[0];
    // outConv inside the last residual unit (kernels_res96.cu tail mode): possible when that unit runs as the fused kernel
    static const int tail_env = []() { const char* e = getenv("Q3TTS_FUSED_TAIL"); return e ? atoi(e) : 1; }();
    const BlockW& B3 = m.blocks[3];
    ResUnitParams probe{};
    probe.C = B3.cout; probe.dil = 1;
    if (tail_env && m.op_dtype != DT_F32 && m.st_dtype == m.op_dtype && B3.conv7[0].taps == 7 && resunit96_supported(probe, m.op_dtype)) {
      CUDA_OK(cudaMalloc(&m.tail_w16, (size_t)16 * B3.cout * 2));
      m.allocs.push_back(m.tail_w16);
      launch_tail_tile(m.tail_w, B3.cout, m.tail_w16, m.op_dtype, m.stream);
      CUDA_OK(cudaStreamSynchronize(m.stream));
      m.fused_tail = true;
    }
  }
  {  // default activation budget: 64 GiB (a batch of 64 x 30 s in one launch chain) or 45 % of what is free now
    size_t free_b = 0, total_b = 0;
    CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    m.default_workspace = std::max<uint64_t>(1ull << 30, std::min<uint64_t>(64ull << 30, (uint64_t)(free_b * 0.45)));
  }
  CUDA_OK(cudaMalloc(&m.d_err, sizeof(int)));
  CUDA_OK(cudaMemset(m.d_err, 0, sizeof(int)));
  CUDA_OK(cudaMallocHost(&m.h_err, sizeof(int)));
  *m.h_err = 0;
  static const char* kStages[] = {"rvq", "pre_conv", "transformer", "upsample", "init_conv",
                                  "block0", "block1", "block2", "block3", "tail"};
  for (const char* n : kStages) {
    StageProfile sp;
    sp.name = n;
    CUDA_OK(cudaEventCreate(&sp.ev0));
    CUDA_OK(cudaEventCreate(&sp.ev1));
    m.prof.push_back(sp);
  }
  CUDA_OK(cudaStreamSynchronize(m.stream));
  CUDA_OK(cudaGetLastError());
  return mp.release();
}

// ---- workspace plan ------------------------------------------------------------------------------
namespace {
struct Arena {
  char* base;
  size_t off = 0;
  template <typename Tp = void>
  Tp* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    Tp* p = base ? (Tp*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

struct Plan {
  void *Q, *QP, *PC, *NB, *QKV, *AO, *GU, *TO, *A0;
  float* H;
  struct Up { float* X; void *N, *G, *XO; } up[8];
  struct Blk { void *X, *A, *C; } blk[4];
  size_t total;
};

Plan make_plan(const Model& m, char* base, int B, int Tmax) {
  const q3tts_config& c = m.cfg;
  const size_t R = (size_t)B * Tmax, op = dt_size(m.op_dtype), st = dt_size(m.st_dtype);
  const size_t A = (size_t)c.num_attention_heads * c.head_dim, KV = (size_t)c.num_key_value_heads * c.head_dim;
  Arena a{base};
  Plan p{};
  p.Q = a.take(R * c.codebook_dim * op);
  p.QP = a.take(R * c.codebook_dim * op);
  p.PC = a.take(R * c.latent_dim * op);
  p.H = a.take<float>(R * c.hidden_size * 4);
  p.NB = a.take(R * c.hidden_size * op);
  p.QKV = a.take(R * (A + 2 * KV) * op);
  p.AO = a.take(R * A * op);
  p.GU = a.take(R * c.intermediate_size * op);
  p.TO = a.take(R * c.latent_dim * op);
  size_t rate = 1;
  for (int i = 0; i < c.num_upsampling_ratios; ++i) {
    rate *= (size_t)c.upsampling_ratios[i];
    p.up[i].X = a.take<float>(R * rate * c.latent_dim * 4);
    p.up[i].N = a.take(R * rate * c.latent_dim * op);
    p.up[i].G = a.take(R * rate * 4 * c.latent_dim * op);
    p.up[i].XO = a.take(R * rate * c.latent_dim * op);
  }
  p.A0 = a.take(R * rate * c.decoder_dim * op);
  for (int i = 0; i < 4; ++i) {
    rate *= (size_t)c.upsample_rates[i];
    const size_t C = (size_t)(c.decoder_dim >> (i + 1));
    p.blk[i].X = a.take(R * rate * C * st);
    p.blk[i].A = a.take(R * rate * C * op);
    p.blk[i].C = a.take(R * rate * C * op);
  }
  p.total = a.off + 256;
  return p;
}
}  // namespace

size_t plan_bytes(const Model& m, int B, int Tmax) { return make_plan(m, nullptr, B, Tmax).total; }

// ---- launch chain ----------------------------------------------------------------------------------
namespace {

struct Ctx {
  Model& m;
  BatchGeom g;
  cudaStream_t s;
  int64_t valid_frames;   // sum of len_frames over the micro-batch (host copy, for the work model)
  int cur_stage = -1;
  // streaming: called right before every consumer that reads rows in front of its tile (buf = its input, rows_per_frame, C, bytes
  // per element): the chunked decode restores the context rows of `buf` from the stream states and saves the new tail
  std::function<void(void*, int, int, int)>* halo_hook = nullptr;
};

inline void halo_in(Ctx& x, const void* buf, int rows_per_frame, int C, int es) {
  if (x.halo_hook) (*x.halo_hook)(const_cast<void*>(buf), rows_per_frame, C, es);
}

void account(Ctx& x, double flops, double bytes) {
  if (x.m.profile_enabled && x.cur_stage >= 0) {
    x.m.prof[(size_t)x.cur_stage].flops += flops;
    x.m.prof[(size_t)x.cur_stage].bytes += bytes;
  }
}

void stage_begin(Ctx& x, int idx) {
  x.cur_stage = idx;
  if (x.m.profile_enabled) {
    StageProfile& sp = x.m.prof[(size_t)idx];
    cudaEventRecord(sp.ev0, x.s);
    sp.used = true;
    sp.launches = 0;
  }
}
void stage_end(Ctx& x) {
  if (x.m.profile_enabled) cudaEventRecord(x.m.prof[(size_t)x.cur_stage].ev1, x.s);
}
void count_launch(Ctx& x, int n = 1) {
  x.m.launches += n;
  if (x.m.profile_enabled && x.cur_stage >= 0) x.m.prof[(size_t)x.cur_stage].launches += n;
}

cudaEvent_t pooled_event(Model& m) {
  if (!m.event_pool.empty()) { cudaEvent_t e = m.event_pool.back(); m.event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// Per-launch timing (profile mode): events bracket ONE kernel launch on the launch stream.
void launch_begin(Ctx& x, const char* op, double flops, double bytes) {
  if (!x.m.profile_enabled || x.cur_stage < 0) return;
  LaunchProfile lp;
  lp.label = x.m.prof[(size_t)x.cur_stage].name + "." + op;
  lp.flops = flops; lp.bytes = bytes;
  lp.ev0 = pooled_event(x.m); lp.ev1 = pooled_event(x.m);
  cudaEventRecord(lp.ev0, x.s);
  x.m.launch_prof.push_back(lp);
}
void launch_end(Ctx& x) {
  if (!x.m.profile_enabled || x.cur_stage < 0 || x.m.launch_prof.empty()) return;
  cudaEventRecord(x.m.launch_prof.back().ev1, x.s);
}

float* tap_buffer(Model& m, const std::string& name, int B, int C, int64_t L) {
  TapBuf& t = m.taps[name];
  const size_t need = (size_t)B * C * L * 4;
  if (need > t.cap) {
    if (t.d) cudaFree(t.d);
    t.d = nullptr;
    CUDA_OK(cudaMalloc(&t.d, need));
    t.cap = need;
  }
  t.B = B; t.C = C; t.L = L;
  return t.d;
}

void tap(Ctx& x, const char* name, const void* src, int dtype, int rows_per_frame, int ld, int C) {
  if (!x.m.taps_enabled) return;
  const int64_t L = (int64_t)x.g.Tmax * rows_per_frame;
  float* dst = tap_buffer(x.m, name, x.g.B, C, L);
  launch_tap_copy(src, dtype, L * ld, ld, dst, x.g.B, C, L, x.s);
  count_launch(x);
}

struct Epi {
  int64_t a_bstride = -1, out_bstride = -1;   // elements between slots of A / of every output; -1 = dense [B, Tmax*rpf, C]
  int act = ACT_NONE;
  const void* res = nullptr; const float* scale = nullptr;
  void* out_y = nullptr; int y_dtype = DT_F32;
  void* out_a = nullptr;
  const SnakeW* snake = nullptr;
  void* out_tap = nullptr;
};

// One multi-tap GEMM over the micro-batch.  A: [B, Tmax*rpf, Cin] operand; outputs have N (or N/2) columns.
void gemm(Ctx& x, const GemmW& w, const void* A, int rows_per_frame, const Epi& e, const char* op = "gemm") {
  Model& m = x.m;
  ConvGemmParams p{};
  const int64_t slot = (int64_t)x.g.Tmax * rows_per_frame;
  p.A = A; p.lda = w.Cin; p.a_bstride = e.a_bstride >= 0 ? e.a_bstride : slot * w.Cin;
  p.W = (m.op_dtype == DT_F32) ? (const void*)w.w32 : (const void*)w.w16;
  p.rows_per_frame = rows_per_frame;
  p.N = w.N; p.Cin = w.Cin; p.taps = w.taps; p.dil = w.dil;
  p.bias = w.bias;
  p.act = e.act;
  const int outN = (e.act == ACT_SWIGLU) ? w.N / 2 : w.N;
  const int64_t obs = e.out_bstride >= 0 ? e.out_bstride : slot * outN;
  p.res = e.res; p.ldres = outN; p.res_bstride = obs;
  p.scale = e.scale;
  p.out_y = e.out_y; p.ldy = outN; p.y_bstride = obs;
  p.out_a = e.out_a; p.lda_out = outN; p.ao_bstride = obs;
  if (e.snake) {
    if (e.snake->n != w.N) throw Error(Q3TTS_EINVAL, "internal: snake width does not match GEMM N");
    p.snake_ea = e.snake->ea; p.snake_ib = e.snake->ib;
  }
  p.out_tap = e.out_tap; p.ldt = outN; p.tap_bstride = obs;
  const int y_dtype = (m.op_dtype == DT_F32) ? DT_F32 : e.y_dtype;
  {
    const double rows = (double)x.valid_frames * rows_per_frame, ops = (double)dt_size(m.op_dtype);
    double bytes = rows * w.Cin * ops + (double)w.taps * w.N * w.Cin * ops;
    if (e.res) bytes += rows * outN * dt_size(y_dtype);
    if (e.out_y) bytes += rows * outN * dt_size(y_dtype);
    if (e.out_a) bytes += rows * outN * ops;
    account(x, 2.0 * rows * w.taps * w.N * w.Cin, bytes);
    launch_begin(x, op, 2.0 * rows * w.taps * w.N * w.Cin, bytes);
  }
  if (m.op_dtype != DT_F32 && tc2_supported(p, m.op_dtype)) {
    cudaError_t err = launch_conv_gemm_tc2(p, x.g, m.op_dtype, y_dtype, x.s);
    if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("tcgen05 GEMM launch: ") + cudaGetErrorString(err));
  } else {
    launch_conv_gemm_simt(p, x.g, m.op_dtype, y_dtype, x.s);
  }
  launch_end(x);
  count_launch(x);
}

}  // namespace

// Stages 4-6: upsample, main decoder, outConv + clip, from the pre-transformer output `to` [B, Tmax, latent] (operand dtype).
static void run_back(Ctx& x, Plan& P, const void* to, const int64_t* d_pcm_base, float* d_pcm) {
  Model& m = x.m;
  const q3tts_config& c = m.cfg;
  cudaStream_t s = x.s;
  const int op = m.op_dtype, B = x.g.B, Tmax = x.g.Tmax;
  const int64_t valid_frames = x.valid_frames;
  const bool taps = m.taps_enabled;
  // 4. upsample: (transposed conv k=s=r, ConvNeXt) x2 (ST.swift:766-775, 385-401)
  stage_begin(x, 3);
  const void* cur = to;
  int rate = 1;
  for (size_t i = 0; i < m.ups.size(); ++i) {
    UpsampleW& U = m.ups[i];
    { Epi e; e.out_y = P.up[i].X; e.y_dtype = DT_F32; gemm(x, U.tconv, cur, rate, e, "convT"); }   // [T*rate, r*L] == [T*rate*r, L]
    rate *= U.ratio;
    halo_in(x, P.up[i].X, rate, c.latent_dim, 4);
    launch_begin(x, "dwconv_ln", 2.0 * 7 * c.latent_dim * (double)valid_frames * rate, (double)valid_frames * rate * c.latent_dim * (4.0 + dt_size(op)));
    launch_dwconv_ln(P.up[i].X, U.dw_w, U.dw_b, U.ln_w, U.ln_b, 1e-6f, P.up[i].N, op, x.g, rate, c.latent_dim, s);
    launch_end(x);
    count_launch(x);
    { Epi e; e.act = ACT_GELU; e.out_a = P.up[i].G; gemm(x, U.pw1, P.up[i].N, rate, e, "pw1"); }
    { Epi e; e.res = P.up[i].X; e.scale = U.gamma; e.y_dtype = DT_F32; e.out_a = P.up[i].XO; gemm(x, U.pw2, P.up[i].G, rate, e, "pw2"); }
    cur = P.up[i].XO;
    tap(x, i == 0 ? "upsample0" : "upsample1", cur, op, rate, c.latent_dim, c.latent_dim);
  }
  stage_end(x);

  // 5. main decoder (ST.swift:681-690): initConv, then 4 x (snake, transposed conv, 3 residual units)
  stage_begin(x, 4);
  {
    Epi e; e.out_a = P.A0; e.snake = &m.block_in_snake[0];
    if (taps) e.out_tap = tap_buffer(m, "init_conv_cl", B, c.decoder_dim, (int64_t)Tmax * rate);
    halo_in(x, cur, rate, c.latent_dim, (int)dt_size(op));
    gemm(x, m.init_conv, cur, rate, e, "conv7");
    if (taps) tap(x, "init_conv", m.taps["init_conv_cl"].d, DT_F32, rate, c.decoder_dim, c.decoder_dim);
  }
  stage_end(x);
  const void* a_in = P.A0;
  for (int i = 0; i < 4; ++i) {
    stage_begin(x, 5 + i);
    BlockW& Bk = m.blocks[i];
    // snake(x) was applied by the producer; transposed conv -> X (stream) and A = res1.act1(X)
    const SnakeW* block_out = i < 3 ? &m.block_in_snake[i + 1] : &m.out_snake;
    ResUnitParams rp{};
    rp.C = Bk.cout; rp.rows_per_frame = rate * Bk.rate; rp.dil = 1;
    const bool fused = op != DT_F32 && m.st_dtype == op && Bk.conv7[0].taps == 7 && resunit96_supported(rp, op);
    {
      Epi e; e.out_y = P.blk[i].X; e.y_dtype = m.st_dtype;
      if (!fused) { e.out_a = P.blk[i].A; e.snake = &Bk.act_in_next[0]; }   // the fused units activate their own input
      halo_in(x, a_in, rate, Bk.tconv.Cin, (int)dt_size(op));
      gemm(x, Bk.tconv, a_in, rate, e, "convT");
    }
    rate *= Bk.rate;
    if (fused) {
      // One kernel per residual unit (kernels_res96.cu): X is read once and written once.  The stream ping-pongs between
      // the block's X and A buffers (a unit must not overwrite rows whose halo another CTA still reads); the last unit
      // writes only the next consumer's operand snake(X').
      void* bufs[4] = {P.blk[i].X, P.blk[i].A, P.blk[i].X, P.blk[i].C};
      for (int j = 0; j < 3; ++j) {
        rp.x_in = bufs[j]; rp.out = bufs[j + 1];
        rp.w7 = Bk.conv7[j].w16; rp.w1 = Bk.conv1[j].w16; rp.b7 = Bk.conv7[j].bias; rp.b1 = Bk.conv1[j].bias;
        rp.ea1 = Bk.act_in_next[j].ea; rp.ib1 = Bk.act_in_next[j].ib; rp.ea2 = Bk.act2[j].ea; rp.ib2 = Bk.act2[j].ib;
        rp.ea3 = j == 2 ? block_out->ea : nullptr; rp.ib3 = j == 2 ? block_out->ib : nullptr;
        rp.dil = Bk.conv7[j].dil;
        const bool tail = i == 3 && j == 2 && m.fused_tail;        // the decoder's last unit: outConv's products instead of activations
        rp.wout = tail ? m.tail_w16 : nullptr; rp.p_out = tail ? (float*)P.blk[i].C : nullptr;
        halo_in(x, bufs[j], rate, Bk.cout, (int)dt_size(op));
        const double rows = (double)valid_frames * rate, C = (double)Bk.cout;
        // X in, X' out (the halo stays on chip), weights once; tail mode: 16 fp32 partial products per row go out instead
        const double fl = 2.0 * rows * (8.0 * C * C + (tail ? 16.0 * C : 0.0)), by = rows * (C * 2.0 + (tail ? 64.0 : C * 2.0)) + 8.0 * C * C * 2.0;
        launch_begin(x, "resunit", fl, by);
        cudaError_t err = launch_resunit96(rp, x.g, op, s);
        if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("fused residual unit launch: ") + cudaGetErrorString(err));
        launch_end(x);
        count_launch(x);
        account(x, fl, by);
      }
      a_in = P.blk[i].C;
      if (taps) {
        // Stage tap with the fused kernels ON: the block's last unit only writes snake_next(X'), so the tap re-runs that unit
        // without the consumer's activation into the (now free) A buffer.  One extra launch, in tap mode only.
        rp.x_in = bufs[2]; rp.out = P.blk[i].A; rp.ea3 = nullptr; rp.ib3 = nullptr; rp.wout = nullptr; rp.p_out = nullptr;
        cudaError_t err = launch_resunit96(rp, x.g, op, s);
        if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("fused residual unit launch (tap): ") + cudaGetErrorString(err));
        count_launch(x);
      }
    } else {
      // conv7 + conv1 of a unit as ONE tcgen05 kernel when the 1x1 conv's operand tile and weights fit in smem (C <= 192):
      // the conv7 output never goes to HBM.  The operand ping-pongs between the A and C buffers (a unit's output operand
      // must not overwrite halo rows its later tiles still read).
      ConvGemmParams probe{};
      probe.N = Bk.conv7[0].N; probe.Cin = Bk.conv7[0].Cin; probe.lda = probe.Cin; probe.taps = Bk.conv7[0].taps; probe.dil = 9;
      probe.bias = Bk.conv7[0].bias; probe.snake_ea = Bk.act2[0].ea; probe.act = ACT_NONE;
      const bool fuse1 = op != DT_F32 && m.st_dtype == op && Bk.conv1[0].taps == 1 && tc2_fuse_supported(probe, op);
      void* abuf[2] = {P.blk[i].A, P.blk[i].C};
      for (int j = 0; j < 3; ++j) {
        const SnakeW* next = (j < 2) ? &Bk.act_in_next[j + 1] : block_out;
        // the block's last unit feeds only the next consumer's operand: its stream output is never read again
        void* x_out = (j == 2 && !taps && op != DT_F32) ? nullptr : P.blk[i].X;
        halo_in(x, fuse1 ? abuf[j & 1] : P.blk[i].A, rate, Bk.cout, (int)dt_size(op));
        if (fuse1) {
          const GemmW& w7 = Bk.conv7[j];
          const GemmW& w1 = Bk.conv1[j];
          const int64_t slot = (int64_t)x.g.Tmax * rate;
          ConvGemmParams p{};
          p.A = abuf[j & 1]; p.lda = w7.Cin; p.a_bstride = slot * w7.Cin; p.W = w7.w16; p.rows_per_frame = rate;
          p.N = w7.N; p.Cin = w7.Cin; p.taps = w7.taps; p.dil = w7.dil; p.bias = w7.bias; p.act = ACT_NONE;
          p.snake_ea = Bk.act2[j].ea; p.snake_ib = Bk.act2[j].ib;
          FusedConv1 f{w1.w16, w1.bias, P.blk[i].X, x_out, abuf[(j + 1) & 1], next->ea, next->ib};
          const double rows = (double)valid_frames * rate, C = (double)w7.N;
          const double fl = 2.0 * rows * 8.0 * C * C, by = rows * C * 2.0 * (x_out ? 4.0 : 3.0) + 8.0 * C * C * 2.0;   // A in, X in, [X' out,] A' out
          launch_begin(x, "conv7+conv1", fl, by);
          cudaError_t err = launch_conv_gemm_tc2(p, x.g, op, op, s, &f);
          if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("fused residual GEMM launch: ") + cudaGetErrorString(err));
          launch_end(x);
          count_launch(x);
          account(x, fl, by);
        } else {
          { Epi e; e.out_a = P.blk[i].C; e.snake = &Bk.act2[j]; gemm(x, Bk.conv7[j], P.blk[i].A, rate, e, "conv7"); }
          { Epi e; e.res = P.blk[i].X; e.out_y = x_out; e.y_dtype = m.st_dtype; e.out_a = P.blk[i].A; e.snake = next; gemm(x, Bk.conv1[j], P.blk[i].C, rate, e, "conv1"); }
        }
      }
      a_in = fuse1 ? abuf[1] : P.blk[i].A;        // three units: A -> C -> A -> C
    }
    static const char* kNames[4] = {"block0", "block1", "block2", "block3"};
    tap(x, kNames[i], (fused && taps) ? P.blk[i].A : P.blk[i].X, m.st_dtype, rate, Bk.cout, Bk.cout);
    stage_end(x);
  }

  // 6. outConv + clip (ST.swift:687-688, 781); outSnake was applied by block3's last epilogue
  stage_begin(x, 9);
  {
    const int C = m.blocks[3].cout;
    float* tp = taps ? tap_buffer(m, "out_conv", B, 1, (int64_t)Tmax * rate) : nullptr;
    if (m.fused_tail) {
      // a_in holds P[row][16] fp32, the outConv partial products block 3's last unit computed on chip
      const double fl = 14.0 * (double)valid_frames * rate, by = (double)valid_frames * rate * (64.0 + 4.0);
      halo_in(x, a_in, rate, 16, 4);
      launch_begin(x, "tap_sum_clip", fl, by);
      launch_tail_from_partials((const float*)a_in, (int64_t)Tmax * rate * 16, m.tail_bias, d_pcm, d_pcm_base, tp, (int64_t)Tmax * rate, x.g, rate, s);
      launch_end(x);
      count_launch(x);
      account(x, fl, by);
    } else {
      halo_in(x, a_in, rate, C, (int)dt_size(op));
      launch_begin(x, "conv7_clip", 2.0 * 7 * C * (double)valid_frames * rate, (double)valid_frames * rate * (C * (double)dt_size(op) + 4.0));
      launch_tail(a_in, op, (int64_t)Tmax * rate * C, m.tail_w, m.tail_bias, C, d_pcm, d_pcm_base, tp, (int64_t)Tmax * rate, x.g, rate, s);
      launch_end(x);
      count_launch(x);
      account(x, 2.0 * 7 * C * (double)valid_frames * rate, (double)valid_frames * rate * (C * (double)dt_size(op) + 4.0));
    }
  }
  stage_end(x);
}

void run_microbatch(Model& m, const int32_t* d_codes, const int64_t* d_code_base, int64_t sq, int64_t st,
                    const int* d_len, const int64_t* d_pcm_base, float* d_pcm, int B, int Tmax, int64_t valid_frames,
                    cudaStream_t s) {
  const q3tts_config& c = m.cfg;
  const size_t need = plan_bytes(m, B, Tmax);
  if (need > m.arena_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.arena) cudaFree(m.arena);
    m.arena = nullptr;
    m.arena_cap = 0;
    CUDA_OK(cudaMalloc(&m.arena, need));
    m.arena_cap = need;
  }
  Plan P = make_plan(m, m.arena, B, Tmax);
  Ctx x{m, BatchGeom{B, Tmax, d_len, (long long)valid_frames, nullptr, m.pcm_i16 ? 1 : 0}, s, valid_frames};
  // per-launch / per-stage event timing brackets serialised kernels: no programmatic overlap between them while profiling
  struct PdlGuard { bool on; explicit PdlGuard(bool o) : on(o) { if (on) pdl_suspend(true); } ~PdlGuard() { if (on) pdl_suspend(false); } } pdl_guard(m.profile_enabled);
  const int op = m.op_dtype;
  const int64_t R = (int64_t)B * Tmax;
  const int half = c.codebook_dim / 2;
  const bool taps = m.taps_enabled;

  // 1. RVQ dequantise (ST.swift:214-226)
  stage_begin(x, 0);
  {
    RvqParams rp{};
    rp.codes = d_codes; rp.code_base = d_code_base; rp.sq = sq; rp.st = st;
    rp.tables = m.d_tables; rp.table_sizes = m.d_table_sizes;
    rp.num_q = c.num_quantizers; rp.num_sem = c.num_semantic_quantizers; rp.half = half;
    rp.out = P.Q; rp.out_dtype = op; rp.err_flag = m.d_err;
    launch_rvq(rp, x.g, s);
    count_launch(x);
    account(x, (double)valid_frames * (c.num_quantizers - 2) * half,
            (double)valid_frames * (c.num_quantizers * (4.0 + half * 4.0) + 2.0 * half * dt_size(op)));
    if (taps) {
      const int64_t L = Tmax;
      launch_tap_copy(P.Q, op, L * 2 * half, 2 * half, tap_buffer(m, "rvq_sum_first", B, half, L), B, half, L, s);
      launch_tap_copy((const char*)P.Q + (size_t)half * dt_size(op), op, L * 2 * half, 2 * half,
                      tap_buffer(m, "rvq_sum_rest", B, half, L), B, half, L, s);
      count_launch(x, 2);
    }
    Epi e; e.out_a = P.QP;
    gemm(x, m.rvq_proj, P.Q, 1, e, "proj");
    tap(x, "quantized", P.QP, op, 1, c.codebook_dim, c.codebook_dim);
  }
  stage_end(x);

  // 2. pre_conv (ST.swift:724-728, 759)
  stage_begin(x, 1);
  { Epi e; e.out_a = P.PC; gemm(x, m.pre_conv, P.QP, 1, e, "conv3"); }
  tap(x, "pre_conv", P.PC, op, 1, c.latent_dim, c.latent_dim);
  stage_end(x);

  // 3. pre_transformer (ST.swift:629-643)
  stage_begin(x, 2);
  { Epi e; e.out_y = P.H; e.y_dtype = DT_F32; gemm(x, m.in_proj, P.PC, 1, e, "in_proj"); }
  const float scale = 1.0f / sqrtf((float)c.head_dim);   // ST.swift:502
  const int window = (m.opts.attn_mode == Q3TTS_ATTN_CAUSAL_SW) ? c.sliding_window : 0;
  for (auto& L : m.layers) {
    launch_rmsnorm(P.H, L.ln1, c.rms_norm_eps, P.NB, op, R, c.hidden_size, s); count_launch(x);
    { Epi e; e.out_a = P.QKV; gemm(x, L.qkv, P.NB, 1, e, "qkv"); }
    {
      const double tkv0 = window > 0 ? std::min<double>(window, (double)valid_frames / B) : (double)valid_frames / B;
      launch_begin(x, "attention", 4.0 * c.num_attention_heads * c.head_dim * tkv0 * valid_frames,
                   (double)valid_frames * ((double)(c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim + (double)c.num_attention_heads * c.head_dim) * dt_size(op));
    }
    launch_attention(P.QKV, op, P.AO, op, x.g, c.num_attention_heads, c.num_key_value_heads, c.head_dim, scale, window, s);
    launch_end(x);
    count_launch(x);
    {  // 4*nh*hd*T_kv FLOP per query frame; keys = own utterance (approximated by the mean valid length)
      const double tkv = window > 0 ? std::min<double>(window, (double)valid_frames / B) : (double)valid_frames / B;
      account(x, 4.0 * c.num_attention_heads * c.head_dim * tkv * valid_frames,
              (double)valid_frames * ((double)(c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim + (double)c.num_attention_heads * c.head_dim) * dt_size(op));
    }
    { Epi e; e.res = P.H; e.scale = L.ls_attn; e.out_y = P.H; e.y_dtype = DT_F32; gemm(x, L.o, P.AO, 1, e, "o_proj"); }
    launch_rmsnorm(P.H, L.ln2, c.rms_norm_eps, P.NB, op, R, c.hidden_size, s); count_launch(x);
    { Epi e; e.act = ACT_SWIGLU; e.out_a = P.GU; gemm(x, L.gate_up, P.NB, 1, e, "gate_up"); }
    { Epi e; e.res = P.H; e.scale = L.ls_mlp; e.out_y = P.H; e.y_dtype = DT_F32; gemm(x, L.down, P.GU, 1, e, "down"); }
  }
  launch_rmsnorm(P.H, m.final_norm, c.rms_norm_eps, P.NB, op, R, c.hidden_size, s); count_launch(x);
  { Epi e; e.out_a = P.TO; gemm(x, m.out_proj, P.NB, 1, e, "out_proj"); }
  tap(x, "pre_transformer", P.TO, op, 1, c.latent_dim, c.latent_dim);
  stage_end(x);

  run_back(x, P, P.TO, d_pcm_base, d_pcm);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(err));
}



// ---- chunked streaming ------------------------------------------------------------------------------------------------
std::vector<HaloStage> conv_state_layout(const Model& m) {
  // mirrors the halo_in() calls of run_back, in order (the streaming hook checks every call against this list)
  const q3tts_config& c = m.cfg;
  const int ops = (int)dt_size(m.op_dtype);
  std::vector<HaloStage> v;
  size_t off = 0;
  auto push = [&](int halo_rows, int rate, int C, int es) {
    HaloStage h{(halo_rows + rate - 1) / rate, rate, C, es, off, 0};
    h.bytes = (size_t)h.frames * rate * C * es;
    off += (h.bytes + 255) & ~(size_t)255;
    v.push_back(h);
  };
  int rate = 1;
  for (size_t i = 0; i < m.ups.size(); ++i) {
    rate *= m.ups[i].ratio;
    push(6, rate, c.latent_dim, 4);                                          // ConvNeXt depthwise k = 7 on the fp32 stream
  }
  push((m.init_conv.taps - 1) * m.init_conv.dil, rate, c.latent_dim, ops);   // initConv
  for (int i = 0; i < 4; ++i) {
    const BlockW& Bk = m.blocks[(size_t)i];
    push(1, rate, Bk.tconv.Cin, ops);                                        // transposed conv k = 2r, s = r: x[t], x[t-1]
    rate *= Bk.rate;
    for (int j = 0; j < 3; ++j) push((Bk.conv7[j].taps - 1) * Bk.conv7[j].dil, rate, Bk.cout, ops);
  }
  if (m.fused_tail) push(6, rate, 16, 4);                                    // outConv k = 7 on the partial products (fp32 [row][16])
  else push(6, rate, m.blocks[3].cout, ops);                                 // outConv k = 7
  return v;
}

void run_microbatch_graphed(Model& m, const int32_t* d_codes, const int64_t* d_code_base, int64_t sq, int64_t st,
                            const int* d_len, const int64_t* d_pcm_base, float* d_pcm, int B, int Tmax, int64_t valid_frames,
                            cudaStream_t s) {
  static const int env_mode = []() { const char* e = getenv("Q3TTS_GRAPHS"); return e ? atoi(e) : -2; }();
  const int mode = env_mode != -2 ? env_mode : m.graph_mode;
  // the legacy default stream cannot be captured
  const bool capturable = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
  const bool want = capturable && !m.profile_enabled && !m.taps_enabled && m.op_dtype != DT_F32 &&
                    (mode == 1 || (mode < 0 && (long long)B * Tmax <= m.graph_max_frames));
  if (!want) {
    run_microbatch(m, d_codes, d_code_base, sq, st, d_len, d_pcm_base, d_pcm, B, Tmax, valid_frames, s);
    return;
  }
  // the arena must not move during (or after) capture: grow it first
  const size_t need = plan_bytes(m, B, Tmax);
  if (need > m.arena_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.arena) cudaFree(m.arena);
    m.arena = nullptr; m.arena_cap = 0;
    CUDA_OK(cudaMalloc(&m.arena, need));
    m.arena_cap = need;
  }
  const std::vector<long long> key = {(long long)(uintptr_t)d_codes, (long long)(uintptr_t)d_code_base, (long long)sq, (long long)st,
                                      (long long)(uintptr_t)d_len, (long long)(uintptr_t)d_pcm_base, (long long)(uintptr_t)d_pcm,
                                      B, Tmax, (long long)valid_frames, (long long)(uintptr_t)m.arena, m.pcm_i16 ? 1 : 0};
  Model::GraphEntry* e = nullptr;
  for (auto& g : m.graphs) if (g.key == key) { e = &g; break; }
  if (e && e->exec) {
    CUDA_OK(cudaGraphLaunch(e->exec, s));
    m.launches += e->launches;
    return;
  }
  if (e && e->seen < 0) {   // this key could not be captured: stay eager
    run_microbatch(m, d_codes, d_code_base, sq, st, d_len, d_pcm_base, d_pcm, B, Tmax, valid_frames, s);
    return;
  }
  if (!e) {   // first sighting: run eagerly (one-off shapes never pay for a capture; lazy one-time setup happens outside capture)
    if (m.graphs.size() >= 16) {
      for (auto& g : m.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
      m.graphs.clear();
    }
    Model::GraphEntry ne; ne.key = key; ne.seen = 1;
    m.graphs.push_back(ne);
    run_microbatch(m, d_codes, d_code_base, sq, st, d_len, d_pcm_base, d_pcm, B, Tmax, valid_frames, s);
    return;
  }
  // second sighting: capture, instantiate, launch
  const long long before = m.launches;
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    (void)cudaGetLastError();
    e->seen = -1;
    run_microbatch(m, d_codes, d_code_base, sq, st, d_len, d_pcm_base, d_pcm, B, Tmax, valid_frames, s);
    return;
  }
  try {
    run_microbatch(m, d_codes, d_code_base, sq, st, d_len, d_pcm_base, d_pcm, B, Tmax, valid_frames, s);
  } catch (...) {
    cudaStreamEndCapture(s, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  CUDA_OK(cudaStreamEndCapture(s, &graph));
  cudaGraphExec_t exec = nullptr;
  cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
  e->exec = exec;
  e->launches = m.launches - before;
  CUDA_OK(cudaGraphLaunch(exec, s));
}

int stream_context_frames(const Model& m) {
  int h = 1;
  for (const HaloStage& st : conv_state_layout(m)) h = std::max(h, st.frames);
  return h;
}

void stream_state_alloc(Model& m, StreamState& st) {
  const q3tts_config& c = m.cfg;
  const size_t op = dt_size(m.op_dtype);
  const size_t ld = (size_t)(c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim;
  const int W1 = std::max(c.sliding_window - 1, 0);
  const std::vector<HaloStage> lay = conv_state_layout(m);
  CUDA_OK(cudaMalloc(&st.q_hist, 2 * (size_t)c.codebook_dim * op));
  CUDA_OK(cudaMalloc(&st.kv, std::max<size_t>((size_t)c.num_hidden_layers * W1 * ld * op, 16)));
  CUDA_OK(cudaMalloc(&st.conv, lay.back().off + ((lay.back().bytes + 255) & ~(size_t)255)));
  st.frames_done = 0;
}
void stream_state_free(Model& m, StreamState& st) {
  cudaSetDevice(m.device);
  if (st.q_hist) cudaFree(st.q_hist);
  if (st.kv) cudaFree(st.kv);
  if (st.conv) cudaFree(st.conv);
  st = StreamState{};
}

void run_stream_batch(Model& m, StreamState* const* streams, int S, const int32_t* d_codes, const int* n_frames,
                      float* d_pcm_out, cudaStream_t s) {
  const q3tts_config& c = m.cfg;
  const int op = m.op_dtype;
  const size_t ops = dt_size(op);
  const int Hc = stream_context_frames(m), W1 = c.sliding_window - 1, Q = c.num_quantizers;
  const std::vector<HaloStage> lay = conv_state_layout(m);
  const int cb = c.codebook_dim, lat = c.latent_dim, hid = c.hidden_size, inter = c.intermediate_size;
  const int A = c.num_attention_heads * c.head_dim, ld = (c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim;
  const int64_t up = c.total_upsample;
  int nmax = 0;
  int64_t total = 0;
  for (int i = 0; i < S; ++i) { nmax = std::max(nmax, n_frames[i]); total += n_frames[i]; }
  if (nmax == 0) return;
  const int Tq = 2 + nmax, Tkv = W1 + nmax, Tc = Hc + nmax;

  // ---- workspace: back-stage plan (arena) + streaming front buffers (stream_ws) ----
  const size_t need = plan_bytes(m, S, Tc);
  if (need > m.arena_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.arena) cudaFree(m.arena);
    m.arena = nullptr; m.arena_cap = 0;
    CUDA_OK(cudaMalloc(&m.arena, need));
    m.arena_cap = need;
  }
  Plan P = make_plan(m, m.arena, S, Tc);
  const int n_items_max = S * (8 + 2 * c.num_hidden_layers);
  Arena a{nullptr};
  auto carve = [&](Arena& ar) {
    struct { int *len_new, *len_q, *len_kv, *beg_kv, *len_c; int64_t *code_base, *pcm_base; CopyItem* items;
             void *Qsum, *QPin, *PC, *NB, *QKV, *AO, *GU, *TO, *Cin; float *H, *pcm; } w;
    w.len_new = ar.take<int>((size_t)S * 4); w.len_q = ar.take<int>((size_t)S * 4); w.len_kv = ar.take<int>((size_t)S * 4);
    w.beg_kv = ar.take<int>((size_t)S * 4); w.len_c = ar.take<int>((size_t)S * 4);
    w.code_base = ar.take<int64_t>((size_t)S * 8); w.pcm_base = ar.take<int64_t>((size_t)S * 8);
    w.items = ar.take<CopyItem>((size_t)n_items_max * sizeof(CopyItem));
    w.Qsum = ar.take((size_t)S * nmax * cb * ops); w.QPin = ar.take((size_t)S * Tq * cb * ops); w.PC = ar.take((size_t)S * Tq * lat * ops);
    w.H = ar.take<float>((size_t)S * nmax * hid * 4); w.NB = ar.take((size_t)S * nmax * hid * ops);
    w.QKV = ar.take((size_t)S * Tkv * ld * ops); w.AO = ar.take((size_t)S * Tkv * A * ops); w.GU = ar.take((size_t)S * nmax * inter * ops);
    w.TO = ar.take((size_t)S * nmax * lat * ops); w.Cin = ar.take((size_t)S * Tc * lat * ops);
    w.pcm = ar.take<float>((size_t)S * Tc * up * 4);
    return w;
  };
  carve(a);
  const size_t ws_need = a.off + 256;
  if (ws_need > m.stream_ws_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.stream_ws) cudaFree(m.stream_ws);
    m.stream_ws = nullptr; m.stream_ws_cap = 0;
    CUDA_OK(cudaMalloc(&m.stream_ws, ws_need));
    m.stream_ws_cap = ws_need;
  }
  Arena b{m.stream_ws};
  auto w = carve(b);

  // ---- metadata + copy lists (host), one pinned upload ----
  const size_t meta_bytes = (size_t)((char*)w.Qsum - (char*)w.len_new);     // everything before the activation buffers
  if (meta_bytes > m.stream_meta_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.stream_meta_h) cudaFreeHost(m.stream_meta_h);
    m.stream_meta_h = nullptr; m.stream_meta_cap = 0;
    CUDA_OK(cudaMallocHost(&m.stream_meta_h, meta_bytes));
    m.stream_meta_cap = meta_bytes;
  } else {
    CUDA_OK(cudaStreamSynchronize(s));   // the previous push may still be reading the staging block
  }
  char* hm = m.stream_meta_h;
  auto host = [&](const void* dev) { return hm + ((const char*)dev - (const char*)w.len_new); };
  int *h_len_new = (int*)host(w.len_new), *h_len_q = (int*)host(w.len_q), *h_len_kv = (int*)host(w.len_kv), *h_beg_kv = (int*)host(w.beg_kv),
      *h_len_c = (int*)host(w.len_c);
  int64_t *h_code_base = (int64_t*)host(w.code_base), *h_pcm_base = (int64_t*)host(w.pcm_base);
  CopyItem* h_items = (CopyItem*)host(w.items);
  // copy lists, grouped by when they run: [pre] histories -> slots; [layer l] cache -> qkv rows, then qkv rows -> cache;
  // [mid] TO -> Cin and history updates; [post] PCM out
  std::vector<CopyItem> pre, mid, post;
  std::vector<std::vector<CopyItem>> kv_in((size_t)c.num_hidden_layers), kv_out((size_t)c.num_hidden_layers);
  int64_t frame_off = 0;
  std::vector<int> ctxc((size_t)S);
  for (int i = 0; i < S; ++i) {
    StreamState& st = *streams[i];
    const int n = n_frames[i];
    const int cq = (int)std::min<int64_t>(2, st.frames_done), ckv = (int)std::min<int64_t>(W1, st.frames_done),
              cc = (int)std::min<int64_t>(Hc, st.frames_done);
    ctxc[(size_t)i] = cc;
    h_len_new[i] = n; h_len_q[i] = n > 0 ? 2 + n : 0; h_len_kv[i] = n > 0 ? W1 + n : 0; h_beg_kv[i] = W1 - ckv; h_len_c[i] = n > 0 ? cc + n : 0;
    h_code_base[i] = frame_off * Q; h_pcm_base[i] = (int64_t)i * Tc * up;
    if (n == 0) continue;
    char* qpin = (char*)w.QPin + (size_t)i * Tq * cb * ops;
    if (cq < 2) pre.push_back({nullptr, qpin, (long long)((2 - cq) * cb * ops)});                                   // causal zero padding
    if (cq > 0) pre.push_back({(char*)st.q_hist + (size_t)(2 - cq) * cb * ops, qpin + (size_t)(2 - cq) * cb * ops, (long long)(cq * cb * ops)});
    // new q history = last two rows of [hist | new] (n + 2 rows in the slot)
    mid.push_back({qpin + (size_t)n * cb * ops, st.q_hist, (long long)(2 * cb * ops)});
    for (int l = 0; l < c.num_hidden_layers; ++l) {
      char* cache = (char*)st.kv + (size_t)l * W1 * ld * ops;
      char* slot = (char*)w.QKV + (size_t)i * Tkv * ld * ops;
      if (ckv > 0) kv_in[(size_t)l].push_back({cache + (size_t)(W1 - ckv) * ld * ops, slot + (size_t)(W1 - ckv) * ld * ops, (long long)((size_t)ckv * ld * ops)});
      if (W1 > 0) kv_out[(size_t)l].push_back({slot + (size_t)n * ld * ops, cache, (long long)((size_t)W1 * ld * ops)});   // rows n .. n+W1-1
    }
    char* cin = (char*)w.Cin + (size_t)i * Tc * lat * ops;
    if (cc > 0) mid.push_back({nullptr, cin, (long long)((size_t)cc * lat * ops)});   // context frames: placeholders (their outputs are never read)
    mid.push_back({(char*)w.TO + (size_t)i * nmax * lat * ops, cin + (size_t)cc * lat * ops, (long long)((size_t)n * lat * ops)});
    post.push_back({(char*)w.pcm + ((size_t)i * Tc + cc) * up * 4, (char*)d_pcm_out + (size_t)frame_off * up * 4, (long long)((size_t)n * up * 4)});
    frame_off += n;
  }
  size_t cursor = 0;
  auto place = [&](const std::vector<CopyItem>& v) {
    const size_t at = cursor;
    if (cursor + v.size() > (size_t)n_items_max) throw Error(Q3TTS_EINVAL, "internal: streaming copy list overflow");
    for (const CopyItem& it : v) h_items[cursor++] = it;
    return at;
  };
  const size_t at_pre = place(pre), at_mid = place(mid), at_post = place(post);
  std::vector<size_t> at_in((size_t)c.num_hidden_layers), at_out((size_t)c.num_hidden_layers);
  for (int l = 0; l < c.num_hidden_layers; ++l) { at_in[(size_t)l] = place(kv_in[(size_t)l]); at_out[(size_t)l] = place(kv_out[(size_t)l]); }
  CUDA_OK(cudaMemcpyAsync(w.len_new, hm, meta_bytes, cudaMemcpyHostToDevice, s));
  auto run_copies = [&](size_t at, size_t n) { launch_block_copy(w.items + at, (int)n, s); if (n) m.launches += 1; };

  // ---- front: RVQ -> pre_conv -> transformer with the KV history ----
  const float scale = 1.0f / sqrtf((float)c.head_dim);
  Ctx xn{m, BatchGeom{S, nmax, w.len_new, (long long)total, nullptr}, s, total};          // the new frames
  Ctx xq{m, BatchGeom{S, Tq, w.len_q, (long long)(total + 2 * S), nullptr}, s, total};       // pre_conv slots: [2 history | new]
  run_copies(at_pre, pre.size());
  {
    RvqParams rp{};
    rp.codes = d_codes; rp.code_base = w.code_base; rp.sq = 1; rp.st = Q;
    rp.tables = m.d_tables; rp.table_sizes = m.d_table_sizes;
    rp.num_q = c.num_quantizers; rp.num_sem = c.num_semantic_quantizers; rp.half = cb / 2;
    rp.out = w.Qsum; rp.out_dtype = op; rp.err_flag = m.d_err;
    launch_rvq(rp, xn.g, s);
    m.launches += 1;
    Epi e; e.out_a = (char*)w.QPin + (size_t)2 * cb * ops; e.out_bstride = (int64_t)Tq * cb;
    gemm(xn, m.rvq_proj, w.Qsum, 1, e, "proj");
  }
  { Epi e; e.out_a = w.PC; gemm(xq, m.pre_conv, w.QPin, 1, e, "conv3"); }
  { Epi e; e.a_bstride = (int64_t)Tq * lat; e.out_y = w.H; e.y_dtype = DT_F32; gemm(xn, m.in_proj, (char*)w.PC + (size_t)2 * lat * ops, 1, e, "in_proj"); }
  BatchGeom gkv{S, Tkv, w.len_kv, (long long)(total + (int64_t)W1 * S), w.beg_kv};
  const int64_t Rn = (int64_t)S * nmax;
  for (int l = 0; l < c.num_hidden_layers; ++l) {
    LayerW& L = m.layers[(size_t)l];
    launch_rmsnorm(w.H, L.ln1, c.rms_norm_eps, w.NB, op, Rn, hid, s); m.launches += 1;
    { Epi e; e.out_a = (char*)w.QKV + (size_t)W1 * ld * ops; e.out_bstride = (int64_t)Tkv * ld; gemm(xn, L.qkv, w.NB, 1, e, "qkv"); }
    run_copies(at_in[(size_t)l], kv_in[(size_t)l].size());
    launch_attention(w.QKV, op, w.AO, op, gkv, c.num_attention_heads, c.num_key_value_heads, c.head_dim, scale, c.sliding_window, s);
    m.launches += 1;
    run_copies(at_out[(size_t)l], kv_out[(size_t)l].size());
    { Epi e; e.a_bstride = (int64_t)Tkv * A; e.res = w.H; e.scale = L.ls_attn; e.out_y = w.H; e.y_dtype = DT_F32;
      gemm(xn, L.o, (char*)w.AO + (size_t)W1 * A * ops, 1, e, "o_proj"); }
    launch_rmsnorm(w.H, L.ln2, c.rms_norm_eps, w.NB, op, Rn, hid, s); m.launches += 1;
    { Epi e; e.act = ACT_SWIGLU; e.out_a = w.GU; gemm(xn, L.gate_up, w.NB, 1, e, "gate_up"); }
    { Epi e; e.res = w.H; e.scale = L.ls_mlp; e.out_y = w.H; e.y_dtype = DT_F32; gemm(xn, L.down, w.GU, 1, e, "down"); }
  }
  launch_rmsnorm(w.H, m.final_norm, c.rms_norm_eps, w.NB, op, Rn, hid, s); m.launches += 1;
  { Epi e; e.out_a = w.TO; gemm(xn, m.out_proj, w.NB, 1, e, "out_proj"); }
  run_copies(at_mid, mid.size());

  // ---- back: the conv stack over [context | new] frames; only the new frames' PCM is kept ----
  int64_t valid_c = 0;
  for (int i = 0; i < S; ++i) valid_c += n_frames[i] > 0 ? ctxc[(size_t)i] + n_frames[i] : 0;
  Ctx xc{m, BatchGeom{S, Tc, w.len_c, (long long)valid_c, nullptr}, s, valid_c};
  // Per-consumer state carry: [restore the context rows the consumer reads] then [save the new tail], two copy lists per
  // haloed consumer, built here and uploaded through a pinned block that is only appended to during the push.
  const size_t hook_items = lay.size() * 2 * (size_t)S;
  if (hook_items * sizeof(CopyItem) > m.stream_hook_cap) {
    CUDA_OK(cudaStreamSynchronize(s));
    if (m.stream_hook_h) cudaFreeHost(m.stream_hook_h);
    if (m.stream_hook_d) cudaFree(m.stream_hook_d);
    m.stream_hook_h = nullptr; m.stream_hook_d = nullptr; m.stream_hook_cap = 0;
    CUDA_OK(cudaMallocHost(&m.stream_hook_h, hook_items * sizeof(CopyItem)));
    CUDA_OK(cudaMalloc(&m.stream_hook_d, hook_items * sizeof(CopyItem)));
    m.stream_hook_cap = hook_items * sizeof(CopyItem);
  }
  size_t hook_k = 0, hook_cursor = 0;
  std::function<void(void*, int, int, int)> hook = [&](void* buf, int rate, int C, int es) {
    if (hook_k >= lay.size()) throw Error(Q3TTS_EINVAL, "internal: more haloed consumers than conv_state_layout lists");
    const HaloStage& h = lay[hook_k++];
    if (h.rate != rate || h.C != C || h.es != es) throw Error(Q3TTS_EINVAL, "internal: conv_state_layout is out of step with run_back");
    const size_t frame_bytes = (size_t)rate * C * es, slot_bytes = (size_t)Tc * frame_bytes;
    CopyItem* items = (CopyItem*)m.stream_hook_h + hook_cursor;
    size_t n_restore = 0, n_save = 0;
    for (int i = 0; i < S; ++i) {       // restore: the last r context frames <- the newest r frames of the state (right-aligned)
      const int n = n_frames[i], cc = ctxc[(size_t)i], r = std::min(h.frames, cc);
      if (n == 0 || r == 0) continue;
      char* slot = (char*)buf + (size_t)i * slot_bytes;
      items[n_restore++] = {(char*)streams[i]->conv + h.off + (size_t)(h.frames - r) * frame_bytes, slot + (size_t)(cc - r) * frame_bytes,
                            (long long)((size_t)r * frame_bytes)};
    }
    for (int i = 0; i < S; ++i) {       // save: the last `keep` frames of [context | new] (all restored or freshly computed)
      const int n = n_frames[i], cc = ctxc[(size_t)i], keep = std::min(h.frames, cc + n);
      if (n == 0) continue;
      char* slot = (char*)buf + (size_t)i * slot_bytes;
      items[n_restore + n_save++] = {slot + (size_t)(cc + n - keep) * frame_bytes, (char*)streams[i]->conv + h.off + (size_t)(h.frames - keep) * frame_bytes,
                                     (long long)((size_t)keep * frame_bytes)};
    }
    const size_t cnt = n_restore + n_save;
    if (cnt == 0) return;
    CopyItem* d_items = (CopyItem*)m.stream_hook_d + hook_cursor;
    CUDA_OK(cudaMemcpyAsync(d_items, items, cnt * sizeof(CopyItem), cudaMemcpyHostToDevice, s));
    if (n_restore) { launch_block_copy(d_items, (int)n_restore, s); m.launches += 1; }
    if (n_save) { launch_block_copy(d_items + n_restore, (int)n_save, s); m.launches += 1; }   // a separate launch: its sources include restored rows
    hook_cursor += cnt;
  };
  xc.halo_hook = &hook;
  run_back(xc, P, w.Cin, w.pcm_base, w.pcm);
  if (hook_k != lay.size()) throw Error(Q3TTS_EINVAL, "internal: fewer haloed consumers than conv_state_layout lists");
  run_copies(at_post, post.size());
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) throw Error(Q3TTS_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(err));
  for (int i = 0; i < S; ++i) streams[i]->frames_done += n_frames[i];
}

}  // namespace q3
