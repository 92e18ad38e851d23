// extern "C" entry points of libqwen3tts_cuda (see include/qwen3tts_cuda.h for the reference
// interface each one replaces).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>

#include "api_internal.hpp"

using namespace q3;

using namespace q3api;

namespace {
#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw Error(_e == cudaErrorMemoryAllocation ? Q3TTS_ENOMEM : Q3TTS_ECUDA,                         \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                  \
  } while (0)

template <typename Tp>
void ensure_dev(Tp** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *p) return;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CUDA_OK(cudaMalloc(p, std::max<size_t>(bytes, 256)));
  *cap = bytes;
}

struct Utt { int64_t code_base, pcm_base; int frames, orig; };

int64_t frames_cap(const Model& m) {
  const size_t per = (plan_bytes(m, 1, 256) + 255) / 256;
  uint64_t ws = m.opts.workspace_bytes ? m.opts.workspace_bytes : m.default_workspace;
  int64_t cap = (int64_t)(ws / per);
  if (m.opts.max_frames_per_launch > 0) cap = std::min<int64_t>(cap, m.opts.max_frames_per_launch);
  return std::max<int64_t>(cap, 1);
}

// Enqueue the decode of `utts` (any order).  d_codes / d_pcm / d_lengths are device pointers.
// Metadata goes through a pinned staging buffer and is re-uploaded only when it changes.
// after_mb(first, count): called after the chain of each micro-batch has been enqueued; [first, first+count) index the utterances in
// decreasing-length order (stable), i.e. `utts` itself when the caller passes it sorted that way.
void decode_core(Model& m, const int32_t* d_codes, std::vector<Utt> utts, int64_t sq, int64_t st, float* d_pcm,
                 int32_t* d_lengths, cudaStream_t s, const std::function<void(int, int)>& after_mb = nullptr) {
  const int n = (int)utts.size();
  if (n == 0) return;
  // lengths first, in the caller's order (ST.swift:831-833)
  std::vector<Utt> orig = utts;
  std::stable_sort(utts.begin(), utts.end(), [](const Utt& a, const Utt& b) { return a.frames > b.frames; });
  while (!utts.empty() && utts.back().frames == 0) utts.pop_back();
  const int nz = (int)utts.size();
  const int64_t cap = frames_cap(m);
  std::vector<MicroBatch> mbs;
  for (int i = 0; i < nz;) {
    MicroBatch mb;
    mb.first = i;
    mb.Tmax = utts[(size_t)i].frames;
    if (mb.Tmax > cap)
      throw Error(Q3TTS_ENOMEM, "an utterance of " + std::to_string(mb.Tmax) + " frames does not fit the workspace (" +
                                    std::to_string(cap) + " frames); raise q3tts_options.workspace_bytes or use the streaming API");
    int b = 0;
    while (i + b < nz && (int64_t)(b + 1) * mb.Tmax <= cap) ++b;
    mb.B = b;
    mbs.push_back(mb);
    i += b;
  }
  if (m.taps_enabled && mbs.size() > 1) throw Error(Q3TTS_EINVAL, "stage taps need the whole batch in one launch chain; reduce B*T");
  // metadata block: [len_sorted int32 x nz][pad][code_base_sorted i64 x nz][pcm_base_sorted i64 x nz]
  //                 [len_orig int32 x n][pad][code_base_orig i64 x n]
  auto al8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
  const size_t o_len = 0, o_cb = al8((size_t)nz * 4), o_pb = o_cb + (size_t)nz * 8, o_lo = o_pb + (size_t)nz * 8,
               o_cbo = o_lo + al8((size_t)n * 4), total = o_cbo + (size_t)n * 8;
  std::vector<char> meta(total, 0);
  for (int i = 0; i < nz; ++i) {
    ((int32_t*)(meta.data() + o_len))[i] = utts[(size_t)i].frames;
    ((int64_t*)(meta.data() + o_cb))[i] = utts[(size_t)i].code_base;
    ((int64_t*)(meta.data() + o_pb))[i] = utts[(size_t)i].pcm_base;
  }
  for (int i = 0; i < n; ++i) {
    ((int32_t*)(meta.data() + o_lo))[i] = orig[(size_t)i].frames;
    ((int64_t*)(meta.data() + o_cbo))[i] = orig[(size_t)i].code_base;
  }
  if (meta != m.meta_key) {
    CUDA_OK(cudaStreamSynchronize(s));   // the previous chain may still be reading the old metadata
    if (total > m.h_meta_cap) {
      if (m.h_meta) cudaFreeHost(m.h_meta);
      m.h_meta = nullptr;
      CUDA_OK(cudaMallocHost(&m.h_meta, total));
      m.h_meta_cap = total;
    }
    ensure_dev(&m.d_meta, &m.d_meta_cap, total);
    std::memcpy(m.h_meta, meta.data(), total);
    CUDA_OK(cudaMemcpyAsync(m.d_meta, m.h_meta, total, cudaMemcpyHostToDevice, s));
    m.meta_key = meta;
  }
  const int* d_len = (const int*)(m.d_meta + o_len);
  const int64_t* d_cb = (const int64_t*)(m.d_meta + o_cb);
  const int64_t* d_pb = (const int64_t*)(m.d_meta + o_pb);
  if (m.profile_enabled) {
    for (auto& sp : m.prof) { sp.used = false; sp.ms_accum = 0; sp.flops = 0; sp.bytes = 0; sp.launches = 0; }
    m.kernel_totals.clear();
  }
  for (auto& mb : mbs) {
    int64_t valid = 0;
    for (int i = 0; i < mb.B; ++i) valid += utts[(size_t)(mb.first + i)].frames;
    run_microbatch_graphed(m, d_codes, d_cb + mb.first, sq, st, d_len + mb.first, d_pb + mb.first, d_pcm, mb.B, mb.Tmax, valid, s);
    if (m.profile_enabled) {
      CUDA_OK(cudaStreamSynchronize(s));
      for (auto& sp : m.prof)
        if (sp.used) {
          float ms = 0;
          if (cudaEventElapsedTime(&ms, sp.ev0, sp.ev1) == cudaSuccess) sp.ms_accum += ms;
        }
      for (auto& lp : m.launch_prof) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, lp.ev0, lp.ev1) == cudaSuccess) {
          auto& t = m.kernel_totals[lp.label];
          t[0] += ms; t[1] += 1; t[2] += lp.flops; t[3] += lp.bytes;
        }
        m.event_pool.push_back(lp.ev0);
        m.event_pool.push_back(lp.ev1);
      }
      m.launch_prof.clear();
    }
    if (after_mb) after_mb(mb.first, mb.B);
  }
  if (d_lengths) {
    launch_lengths(d_codes, (const int64_t*)(m.d_meta + o_cbo), st, (const int*)(m.d_meta + o_lo), n,
                   m.cfg.decode_upsample_rate, d_lengths, s);
    m.launches += 1;
  }
}

void check_device_errors(Model& m, cudaStream_t s) {
  CUDA_OK(cudaMemcpyAsync(m.h_err, m.d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (*m.h_err) {
    *m.h_err = 0;
    CUDA_OK(cudaMemsetAsync(m.d_err, 0, sizeof(int), s));
    throw Error(Q3TTS_EINVAL, "a code id is outside its codebook (semantic ids must be < semantic_codebook_size, acoustic < codebook_size)");
  }
}

}  // namespace

namespace q3api {
// Pinned staging of the host-buffer paths (grow-only, owned by the model).
static void ensure_pinned(char** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *p) return;
  if (*p) cudaFreeHost(*p);
  *p = nullptr; *cap = 0;
  CUDA_OK(cudaMallocHost(p, std::max<size_t>(bytes, 4096)));
  *cap = bytes;
}
static bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

void decode_host_list(Model& m, const int32_t* codes, int32_t layout, int32_t T_uniform, std::vector<HostUtt> utts, void* pcm_out,
                      bool i16, int32_t* lengths_out) {
  const int n = (int)utts.size();
  if (n == 0) return;
  std::lock_guard<std::mutex> lock(m.mu);
  CUDA_OK(cudaSetDevice(m.device));
  struct Fmt { Model& m; Fmt(Model& mm, bool v) : m(mm) { m.pcm_i16 = v; } ~Fmt() { m.pcm_i16 = false; } } fmt(m, i16);
  const int Q = m.cfg.num_quantizers;
  const int64_t up = m.cfg.total_upsample;
  const size_t ss = i16 ? 2 : 4;                       // bytes per PCM sample on the way out
  // decreasing length (stable): decode_core's own order, so device PCM of a micro-batch is one contiguous range
  std::stable_sort(utts.begin(), utts.end(), [](const HostUtt& a, const HostUtt& b) { return a.frames > b.frames; });
  int64_t total = 0;
  for (const HostUtt& u : utts) total += u.frames;
  if (total == 0) {
    if (lengths_out) for (const HostUtt& u : utts) lengths_out[u.orig] = 0;
    return;
  }
  cudaStream_t s = m.stream;
  chain_begin(m, s);
  ensure_dev(&m.d_codes, &m.d_codes_cap, (size_t)total * Q * 4);
  ensure_dev(&m.d_pcm, &m.d_pcm_cap, (size_t)total * up * 4);
  ensure_dev(&m.d_lengths, &m.d_lengths_cap, (size_t)n * 4);
  if (!m.copy_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&m.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) { CUDA_OK(cudaEventCreateWithFlags(&m.mb_done[i], cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&m.d2h_done[i], cudaEventDisableTiming)); }
  }
  // codes: packed in sorted order through pinned staging (a few MB), one H2D
  ensure_pinned(&m.h_codes, &m.h_codes_cap, (size_t)total * Q * 4);
  std::vector<Utt> dev((size_t)n);
  {
    int64_t f = 0;
    for (int i = 0; i < n; ++i) {
      const HostUtt& u = utts[(size_t)i];
      if (u.frames > 0) std::memcpy(m.h_codes + (size_t)f * Q * 4, codes + u.code_off, (size_t)u.frames * Q * 4);
      dev[(size_t)i] = Utt{f * Q, f * up, u.frames, i};
      f += u.frames;
    }
  }
  CUDA_OK(cudaMemcpyAsync(m.d_codes, m.h_codes, (size_t)total * Q * 4, cudaMemcpyHostToDevice, s));
  const int64_t sq = layout == Q3TTS_CODES_BQT ? T_uniform : 1, st = layout == Q3TTS_CODES_BQT ? 1 : Q;
  const bool direct = is_pinned(pcm_out);              // pinned / registered destination: DMA straight into it
  struct Pending { int first = 0, count = 0, buf = 0; bool live = false; } pend;
  auto scatter = [&](const Pending& pd) {               // pageable destination: pinned staging -> the caller's buffer
    CUDA_OK(cudaEventSynchronize(m.d2h_done[pd.buf]));
    if (direct) return;
    const char* src = m.h_pcm[pd.buf];
    for (int i = pd.first; i < pd.first + pd.count; ++i) {
      const HostUtt& u = utts[(size_t)i];
      const size_t bytes = (size_t)u.frames * up * ss;
      std::memcpy((char*)pcm_out + (size_t)u.pcm_off * ss, src, bytes);
      src += bytes;
    }
  };
  int k = 0;
  auto after_mb = [&](int first, int count) {
    const int buf = k & 1;
    CUDA_OK(cudaEventRecord(m.mb_done[buf], s));
    CUDA_OK(cudaStreamWaitEvent(m.copy_stream, m.mb_done[buf], 0));
    int64_t f0 = dev[(size_t)first].pcm_base, fr = 0;
    for (int i = first; i < first + count; ++i) fr += utts[(size_t)i].frames;
    const char* d_src = (const char*)m.d_pcm + (size_t)f0 * ss;   // the tail wrote `ss`-byte samples at sample offset pcm_base
    if (direct) {
      // coalesce neighbours whose destinations are contiguous too (a uniform batch is ONE copy)
      int i = first;
      while (i < first + count) {
        int j = i;
        size_t bytes = (size_t)utts[(size_t)i].frames * up * ss;
        while (j + 1 < first + count && utts[(size_t)j + 1].pcm_off == utts[(size_t)j].pcm_off + (int64_t)utts[(size_t)j].frames * up) {
          ++j;
          bytes += (size_t)utts[(size_t)j].frames * up * ss;
        }
        CUDA_OK(cudaMemcpyAsync((char*)pcm_out + (size_t)utts[(size_t)i].pcm_off * ss, d_src, bytes, cudaMemcpyDeviceToHost, m.copy_stream));
        d_src += bytes;
        i = j + 1;
      }
    } else {
      ensure_pinned(&m.h_pcm[buf], &m.h_pcm_cap[buf], (size_t)fr * up * ss);
      CUDA_OK(cudaMemcpyAsync(m.h_pcm[buf], d_src, (size_t)fr * up * ss, cudaMemcpyDeviceToHost, m.copy_stream));
    }
    CUDA_OK(cudaEventRecord(m.d2h_done[buf], m.copy_stream));
    // the previous micro-batch's copy has had a whole chain's worth of enqueueing to finish: drain it now, while this chain runs
    if (pend.live) scatter(pend);
    pend = Pending{first, count, buf, true};
    ++k;
  };
  // NOTE on d_pcm addressing: the tail kernels address samples by INDEX (pcm_base + t) in the output's own sample size, so a
  // micro-batch occupies bytes [pcm_base * ss, ...) of d_pcm in either format.
  decode_core(m, m.d_codes, dev, sq, st, m.d_pcm, lengths_out ? m.d_lengths : nullptr, s, after_mb);
  if (lengths_out) {
    ensure_pinned(&m.h_len, &m.h_len_cap, (size_t)n * 4);
    CUDA_OK(cudaMemcpyAsync(m.h_len, m.d_lengths, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  }
  // the next chain (any stream) must not overwrite d_pcm before the last copy-out has read it
  CUDA_OK(cudaStreamWaitEvent(s, m.d2h_done[(k - 1) & 1], 0));
  chain_end(m, s);
  if (pend.live) scatter(pend);
  check_device_errors(m, s);                            // synchronises s (and, through the wait above, the copy stream)
  if (lengths_out) for (int i = 0; i < n; ++i) lengths_out[utts[(size_t)i].orig] = ((const int32_t*)m.h_len)[i];
}
}  // namespace q3api

struct q3tts_stream { q3tts_model* owner; StreamState state; };
namespace { std::mutex g_lifetime_mu; }

extern "C" {


int q3tts_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p{};
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

void q3tts_options_default(q3tts_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->struct_size = sizeof(*o);
  o->device = -1;
  o->precision = Q3TTS_PREC_FP16;
  o->attn_mode = Q3TTS_ATTN_REFERENCE;
}

int q3tts_model_load(const char* dir, const q3tts_options* opts, q3tts_model** out) {
  return guarded([&]() {
    if (!dir || !out) return fail(Q3TTS_EINVAL, "NULL argument");
    *out = nullptr;
    q3tts_options o;
    q3tts_options_default(&o);
    if (opts) {
      if (opts->struct_size != sizeof(q3tts_options)) return fail(Q3TTS_EINVAL, "q3tts_options.struct_size mismatch");
      o = *opts;
    }
    if (o.precision < Q3TTS_PREC_FP32 || o.precision > Q3TTS_PREC_BF16) return fail(Q3TTS_EINVAL, "bad precision");
    if (o.attn_mode != Q3TTS_ATTN_REFERENCE && o.attn_mode != Q3TTS_ATTN_CAUSAL_SW) return fail(Q3TTS_EINVAL, "bad attn_mode");
    Checkpoint ck;
    load_checkpoint(dir, &ck);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      return fail(Q3TTS_ECUDA, "no CUDA device: libqwen3tts_cuda has no CPU fallback");
    }
    std::unique_ptr<q3tts_model> h(new q3tts_model());
    h->m = nullptr;
    h->m = model_create(ck, o);
    *out = h.release();
    return (int)Q3TTS_OK;
  });
}

void q3tts_model_free(q3tts_model* h) {
  if (!h) return;
  {
    std::lock_guard<std::mutex> life(g_lifetime_mu);
    if (h->open_streams > 0) { h->zombie = true; return; }   // deleted by the last q3tts_stream_close
  }
  delete h->m;
  delete h;
}

int q3tts_model_config(const q3tts_model* h, q3tts_config* out) {
  if (!h || !out) return fail(Q3TTS_EINVAL, "NULL argument");
  *out = h->m->cfg;
  return Q3TTS_OK;
}

int64_t q3tts_output_samples(const q3tts_model* h, int64_t frames) {
  if (!h || frames < 0) return -1;
  return frames * h->m->cfg.total_upsample;
}

// RAII: the tail of this call writes 16-bit PCM (the model mutex is held for the whole call)
struct PcmFormat {
  Model& m;
  PcmFormat(Model& mm, bool i16) : m(mm) { m.pcm_i16 = i16; }
  ~PcmFormat() { m.pcm_i16 = false; }
};

static int decode_common(q3tts_model* h, const int32_t* codes, int32_t B, int32_t T, int32_t layout, float* pcm,
                         int32_t* lengths, bool device_ptrs, cudaStream_t user_stream, bool i16 = false) {
  return guarded([&]() {
    if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
    if (B < 0 || T < 0) return fail(Q3TTS_EINVAL, "negative B or T");
    if (layout != Q3TTS_CODES_BQT && layout != Q3TTS_CODES_BTQ) return fail(Q3TTS_EINVAL, "bad codes layout");
    if (B == 0 || T == 0) return (int)Q3TTS_OK;   // empty input: nothing to decode
    if (!codes || !pcm) return fail(Q3TTS_EINVAL, "NULL buffer");
    Model& m = *h->m;
    const int Q = m.cfg.num_quantizers;
    const int64_t up = m.cfg.total_upsample;
    if (!device_ptrs) {
      std::vector<HostUtt> utts((size_t)B);
      for (int b = 0; b < B; ++b) utts[(size_t)b] = HostUtt{(int64_t)b * T * Q, (int64_t)b * T * up, T, b};
      decode_host_list(m, codes, layout, T, std::move(utts), pcm, i16, lengths);
      return (int)Q3TTS_OK;
    }
    std::lock_guard<std::mutex> lock(m.mu);
    CUDA_OK(cudaSetDevice(m.device));
    PcmFormat fmt(m, i16);
    cudaStream_t s = user_stream;
    chain_begin(m, s);   // the workspace, d_meta and d_err are shared with the previous chain, which may have run on another stream
    std::vector<Utt> utts((size_t)B);
    for (int b = 0; b < B; ++b) utts[(size_t)b] = Utt{(int64_t)b * T * Q, (int64_t)b * T * up, T, b};
    const int64_t sq = layout == Q3TTS_CODES_BQT ? T : 1, st = layout == Q3TTS_CODES_BQT ? 1 : Q;
    decode_core(m, codes, utts, sq, st, pcm, lengths, s);
    chain_end(m, s);
    return (int)Q3TTS_OK;
  });
}

int q3tts_decode(q3tts_model* h, const int32_t* codes, int32_t B, int32_t T, int32_t layout, float* pcm_out,
                 int32_t* lengths_out) {
  return decode_common(h, codes, B, T, layout, pcm_out, lengths_out, false, nullptr);
}

int q3tts_decode_int16(q3tts_model* h, const int32_t* codes, int32_t B, int32_t T, int32_t layout, int16_t* pcm_out,
                       int32_t* lengths_out) {
  return decode_common(h, codes, B, T, layout, (float*)pcm_out, lengths_out, false, nullptr, true);
}

int q3tts_decode_device(q3tts_model* h, const int32_t* d_codes, int32_t B, int32_t T, int32_t layout, float* d_pcm_out,
                        int32_t* d_lengths_out, void* stream) {
  return decode_common(h, d_codes, B, T, layout, d_pcm_out, d_lengths_out, true, (cudaStream_t)stream);
}

// Mixed-length batch with the packed codes and the PCM already in HBM; frame_offsets stays a HOST array (it sizes the launch chain).
int q3tts_decode_varlen_device(q3tts_model* h, const int32_t* d_codes_packed, const int64_t* frame_offsets, int32_t n,
                               float* d_pcm_out, int32_t* d_lengths_out, void* stream) {
  return guarded([&]() {
    if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
    if (n < 0) return fail(Q3TTS_EINVAL, "negative utterance count");
    if (n == 0) return (int)Q3TTS_OK;
    if (!frame_offsets) return fail(Q3TTS_EINVAL, "frame_offsets is NULL");
    if (frame_offsets[0] != 0) return fail(Q3TTS_EINVAL, "frame_offsets[0] must be 0");
    for (int i = 0; i < n; ++i)
      if (frame_offsets[i + 1] < frame_offsets[i] || frame_offsets[i + 1] - frame_offsets[i] > INT32_MAX)
        return fail(Q3TTS_EINVAL, "frame_offsets must be non-decreasing");
    if (frame_offsets[n] > 0 && (!d_codes_packed || !d_pcm_out)) return fail(Q3TTS_EINVAL, "NULL buffer");
    Model& m = *h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    CUDA_OK(cudaSetDevice(m.device));
    PcmFormat fmt(m, false);
    const int Q = m.cfg.num_quantizers;
    const int64_t up = m.cfg.total_upsample;
    cudaStream_t s = (cudaStream_t)stream;
    chain_begin(m, s);
    std::vector<Utt> utts((size_t)n);
    for (int i = 0; i < n; ++i)
      utts[(size_t)i] = Utt{frame_offsets[i] * Q, frame_offsets[i] * up, (int)(frame_offsets[i + 1] - frame_offsets[i]), i};
    decode_core(m, d_codes_packed, utts, 1, Q, d_pcm_out, d_lengths_out, s);
    chain_end(m, s);
    return (int)Q3TTS_OK;
  });
}

int q3tts_sync(q3tts_model* h, void* stream) {
  return guarded([&]() {
    if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
    Model& m = *h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    CUDA_OK(cudaSetDevice(m.device));
    chain_begin(m, (cudaStream_t)stream);
    check_device_errors(m, (cudaStream_t)stream);
    chain_end(m, (cudaStream_t)stream);
    return (int)Q3TTS_OK;
  });
}

static int decode_varlen_common(q3tts_model* h, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n,
                                void* pcm_out, int32_t* lengths_out, bool i16) {
  return guarded([&]() {
    if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
    if (n < 0) return fail(Q3TTS_EINVAL, "negative utterance count");
    if (n == 0) return (int)Q3TTS_OK;
    if (!frame_offsets) return fail(Q3TTS_EINVAL, "frame_offsets is NULL");
    for (int i = 0; i < n; ++i)
      if (frame_offsets[i + 1] < frame_offsets[i] || frame_offsets[i + 1] - frame_offsets[i] > INT32_MAX)
        return fail(Q3TTS_EINVAL, "frame_offsets must be non-decreasing");
    if (frame_offsets[0] != 0) return fail(Q3TTS_EINVAL, "frame_offsets[0] must be 0");
    const int64_t total = frame_offsets[n];
    Model& m = *h->m;
    if (total == 0) {
      if (lengths_out) std::memset(lengths_out, 0, (size_t)n * 4);
      return (int)Q3TTS_OK;
    }
    if (!codes_packed || !pcm_out) return fail(Q3TTS_EINVAL, "NULL buffer");
    const int Q = m.cfg.num_quantizers;
    const int64_t up = m.cfg.total_upsample;
    std::vector<HostUtt> utts((size_t)n);
    for (int i = 0; i < n; ++i)
      utts[(size_t)i] = HostUtt{frame_offsets[i] * Q, frame_offsets[i] * up, (int)(frame_offsets[i + 1] - frame_offsets[i]), i};
    decode_host_list(m, codes_packed, Q3TTS_CODES_BTQ, 0, std::move(utts), pcm_out, i16, lengths_out);
    return (int)Q3TTS_OK;
  });
}

int q3tts_decode_varlen(q3tts_model* h, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n,
                        float* pcm_out, int32_t* lengths_out) {
  return decode_varlen_common(h, codes_packed, frame_offsets, n, pcm_out, lengths_out, false);
}

int q3tts_decode_varlen_int16(q3tts_model* h, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n,
                              int16_t* pcm_out, int32_t* lengths_out) {
  return decode_varlen_common(h, codes_packed, frame_offsets, n, pcm_out, lengths_out, true);
}

// ---- taps -------------------------------------------------------------------------------------------
int q3tts_set_taps(q3tts_model* h, int32_t enable) {
  if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
  std::lock_guard<std::mutex> lock(h->m->mu);
  h->m->taps_enabled = enable != 0;
  return Q3TTS_OK;
}

int q3tts_set_graphs(q3tts_model* h, int32_t mode) {
  if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
  if (mode < -1 || mode > 1) return fail(Q3TTS_EINVAL, "mode must be -1 (automatic), 0 (off) or 1 (always)");
  std::lock_guard<std::mutex> lock(h->m->mu);
  h->m->graph_mode = mode;
  return Q3TTS_OK;
}

int q3tts_stage_tap_shape(q3tts_model* h, const char* name, int32_t* B, int32_t* C, int64_t* L) {
  if (!h || !name) return fail(Q3TTS_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lock(h->m->mu);
  auto it = h->m->taps.find(name);
  if (it == h->m->taps.end() || !it->second.d) return fail(Q3TTS_EINVAL, std::string("no tap named '") + name + "' (enable taps, then decode)");
  if (B) *B = it->second.B;
  if (C) *C = it->second.C;
  if (L) *L = it->second.L;
  return Q3TTS_OK;
}

int q3tts_stage_tap(q3tts_model* h, const char* name, float* out, int64_t out_elems) {
  return guarded([&]() {
    if (!h || !name || !out) return fail(Q3TTS_EINVAL, "NULL argument");
    Model& m = *h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    auto it = m.taps.find(name);
    if (it == m.taps.end() || !it->second.d) return fail(Q3TTS_EINVAL, std::string("no tap named '") + name + "'");
    const int64_t n = (int64_t)it->second.B * it->second.C * it->second.L;
    if (out_elems < n) return fail(Q3TTS_EINVAL, "tap output buffer too small");
    CUDA_OK(cudaSetDevice(m.device));
    CUDA_OK(cudaStreamSynchronize(m.stream));
    CUDA_OK(cudaMemcpy(out, it->second.d, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return (int)Q3TTS_OK;
  });
}

int q3tts_weight_shape(const q3tts_model* h, const char* key, int32_t* ndim, int64_t dims[4]) {
  if (!h || !key || !ndim || !dims) return fail(Q3TTS_EINVAL, "NULL argument");
  auto it = h->m->weight_shapes.find(key);
  if (it == h->m->weight_shapes.end()) return fail(Q3TTS_EINVAL, std::string("no weight named '") + key + "'");
  *ndim = (int32_t)it->second.size();
  for (size_t i = 0; i < it->second.size() && i < 4; ++i) dims[i] = it->second[i];
  return Q3TTS_OK;
}

// ---- streaming: chunked decode with causal state carry (no counterpart in the reference) -------------------
static int stream_push_common(q3tts_stream* const* streams, int32_t n_streams, const int32_t* const* codes, const int32_t* n_frames,
                              float* const* pcm_out) {
  return guarded([&]() {
    if (n_streams < 0) return fail(Q3TTS_EINVAL, "negative stream count");
    if (n_streams == 0) return (int)Q3TTS_OK;
    if (!streams || !codes || !n_frames || !pcm_out) return fail(Q3TTS_EINVAL, "NULL argument");
    q3tts_model* owner = nullptr;
    int64_t total = 0;
    for (int i = 0; i < n_streams; ++i) {
      if (!streams[i] || !streams[i]->owner) return fail(Q3TTS_ESTATE, "push on a closed or NULL stream");
      if (!owner) owner = streams[i]->owner;
      if (streams[i]->owner != owner) return fail(Q3TTS_EINVAL, "all streams of one batched push must belong to the same model (one GPU)");
      if (n_frames[i] < 0) return fail(Q3TTS_EINVAL, "negative frame count");
      if (n_frames[i] > 0 && (!codes[i] || !pcm_out[i])) return fail(Q3TTS_EINVAL, "NULL buffer");
      for (int j = 0; j < i; ++j)
        if (streams[j] == streams[i]) return fail(Q3TTS_EINVAL, "a stream appears twice in one batched push");
      total += n_frames[i];
    }
    if (total == 0) return (int)Q3TTS_OK;
    Model& m = *owner->m;
    std::lock_guard<std::mutex> lock(m.mu);
    CUDA_OK(cudaSetDevice(m.device));
    const int Q = m.cfg.num_quantizers;
    const int64_t up = m.cfg.total_upsample;
    // Code ids are checked HERE, before anything is enqueued: a push commits KV / conv state for every stream of the batch, so a bad
    // id found on the device afterwards would leave all of them advanced past frames whose PCM the caller never got.
    for (int i = 0; i < n_streams; ++i)
      for (int64_t t = 0; t < n_frames[i]; ++t)
        for (int qi = 0; qi < Q; ++qi) {
          const int32_t c = codes[i][t * Q + qi];
          const int32_t lim = qi < m.cfg.num_semantic_quantizers ? m.cfg.semantic_codebook_size : m.cfg.codebook_size;
          if (c < 0 || c >= lim)
            return fail(Q3TTS_EINVAL, "a code id is outside its codebook (stream " + std::to_string(i) + ", frame " + std::to_string(t) +
                                          ", codebook " + std::to_string(qi) + "); no stream of this push has been advanced");
        }
    cudaStream_t s = m.stream;
    chain_begin(m, s);
    ensure_dev(&m.d_codes, &m.d_codes_cap, (size_t)total * Q * 4);
    ensure_dev(&m.d_pcm, &m.d_pcm_cap, (size_t)total * up * 4);
    std::vector<StreamState*> st((size_t)n_streams);
    int64_t off = 0;
    for (int i = 0; i < n_streams; ++i) {
      st[(size_t)i] = &streams[i]->state;
      if (n_frames[i] > 0)
        CUDA_OK(cudaMemcpyAsync(m.d_codes + off * Q, codes[i], (size_t)n_frames[i] * Q * 4, cudaMemcpyHostToDevice, s));
      off += n_frames[i];
    }
    run_stream_batch(m, st.data(), n_streams, m.d_codes, n_frames, m.d_pcm, s);
    off = 0;
    for (int i = 0; i < n_streams; ++i) {
      if (n_frames[i] > 0)
        CUDA_OK(cudaMemcpyAsync(pcm_out[i], m.d_pcm + off * up, (size_t)n_frames[i] * up * 4, cudaMemcpyDeviceToHost, s));
      off += n_frames[i];
    }
    chain_end(m, s);
    check_device_errors(m, s);
    return (int)Q3TTS_OK;
  });
}

int q3tts_stream_open(q3tts_model* h, q3tts_stream** out) {
  return guarded([&]() {
    if (out) *out = nullptr;
    if (!h || !out) return fail(Q3TTS_EINVAL, "NULL argument");
    Model& m = *h->m;
    if (m.opts.attn_mode != Q3TTS_ATTN_CAUSAL_SW)
      return fail(Q3TTS_ESTATE, "chunked streaming needs Q3TTS_ATTN_CAUSAL_SW: the reference's full bidirectional attention (ST.swift:512-528, 763) makes every sample depend on the whole utterance");
    if (m.cfg.sliding_window < 1) return fail(Q3TTS_EFORMAT, "sliding_window must be >= 1 for streaming");
    std::lock_guard<std::mutex> lock(m.mu);
    CUDA_OK(cudaSetDevice(m.device));
    std::unique_ptr<q3tts_stream> st(new q3tts_stream{h, StreamState{}});
    stream_state_alloc(m, st->state);
    {
      std::lock_guard<std::mutex> life(g_lifetime_mu);
      if (h->zombie) { stream_state_free(m, st->state); return fail(Q3TTS_ESTATE, "the model has been freed"); }
      ++h->open_streams;
    }
    *out = st.release();
    return (int)Q3TTS_OK;
  });
}

int q3tts_stream_push(q3tts_stream* s, const int32_t* codes, int32_t n_frames, float* pcm_out) {
  q3tts_stream* one[1] = {s};
  const int32_t* c1[1] = {codes};
  float* p1[1] = {pcm_out};
  return stream_push_common(one, 1, c1, &n_frames, p1);
}

int q3tts_stream_push_batch(q3tts_stream* const* streams, int32_t n_streams, const int32_t* const* codes, const int32_t* n_frames,
                            float* const* pcm_out) {
  return stream_push_common(streams, n_streams, codes, n_frames, pcm_out);
}

int64_t q3tts_stream_frames(const q3tts_stream* s) { return s && s->owner ? s->state.frames_done : -1; }

void q3tts_stream_close(q3tts_stream* s) {
  if (!s) return;
  q3tts_model* owner = s->owner;
  bool last_of_zombie = false;
  if (owner) {
    {
      std::lock_guard<std::mutex> lock(owner->m->mu);
      cudaSetDevice(owner->m->device);
      cudaStreamSynchronize(owner->m->stream);
      stream_state_free(*owner->m, s->state);
    }
    std::lock_guard<std::mutex> life(g_lifetime_mu);
    last_of_zombie = --owner->open_streams == 0 && owner->zombie;
  }
  delete s;
  if (last_of_zombie) { delete owner->m; delete owner; }
}

// ---- kernel-level test / micro-benchmark hook -----------------------------------------------------------------------
// One multi-tap GEMM on seeded random data: the tcgen05 path against the CUDA-core 16-bit path of the same op, then
// `iters` timed launches (CUDA events).  mode 0: conv7-like (bias, SnakeBeta operand out); 1: conv1-like (bias,
// in-place 16-bit residual stream, SnakeBeta operand out); 2: transposed-conv-like (bias, stream out, SnakeBeta operand
// out); 3: plain (bias, operand out); 4: stream out only.  Utterance b has rows - 37*b valid rows (ragged tiles).
int q3tts_debug_conv_gemm(int32_t B, int32_t rows, int32_t Cin, int32_t N, int32_t taps, int32_t dil, int32_t mode,
                          int32_t precision, int32_t iters, float* ms_out, float* max_diff_y, float* max_diff_a) {
  return guarded([&]() {
    if (B < 1 || rows < 1 || Cin < 16 || N < 16 || taps < 1 || dil < 1 || mode < 0 || mode > 4) return fail(Q3TTS_EINVAL, "bad GEMM shape");
    const int op = precision == Q3TTS_PREC_BF16 ? DT_BF16 : DT_F16;
    cudaStream_t s = nullptr;
    CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<void*> allocs;
    auto dmalloc = [&](size_t bytes) { void* d = nullptr; CUDA_OK(cudaMalloc(&d, bytes)); allocs.push_back(d); return d; };
    uint64_t seed = 0x9E3779B97F4A7C15ull ^ ((uint64_t)Cin << 32) ^ (uint64_t)N ^ ((uint64_t)taps << 20);
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (float)((seed >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f; };
    auto upload16 = [&](size_t n, float scale) {
      std::vector<float> h(n);
      for (auto& v : h) v = rnd() * scale;
      float* d32 = (float*)dmalloc(n * 4);
      void* d16 = dmalloc(n * 2);
      CUDA_OK(cudaMemcpyAsync(d32, h.data(), n * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      launch_convert(d32, d16, op, (int64_t)n, s);
      return d16;
    };
    auto upload32 = [&](size_t n, float scale, float offset) {
      std::vector<float> h(n);
      for (auto& v : h) v = offset + rnd() * scale;
      float* d = (float*)dmalloc(n * 4);
      CUDA_OK(cudaMemcpyAsync(d, h.data(), n * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      return d;
    };
    const size_t R = (size_t)B * rows;
    void* A = upload16(R * Cin, 1.0f);
    void* W = upload16((size_t)taps * N * Cin, 1.0f / sqrtf((float)(taps * Cin)));
    void* X = upload16(R * N, 1.0f);
    float* bias = upload32((size_t)N, 0.1f, 0.f);
    float* ea = upload32((size_t)N, 0.3f, 1.0f);
    float* ib = upload32((size_t)N, 0.3f, 1.0f);
    std::vector<int> len((size_t)B);
    for (int b = 0; b < B; ++b) len[(size_t)b] = std::max(1, rows - 37 * b);
    int* d_len = (int*)dmalloc((size_t)B * 4);
    CUDA_OK(cudaMemcpyAsync(d_len, len.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    void* y[2] = {dmalloc(R * N * 2), dmalloc(R * N * 2)};
    void* a[2] = {dmalloc(R * N * 2), dmalloc(R * N * 2)};
    long long valid_rows = 0;
    for (int b = 0; b < B; ++b) valid_rows += len[(size_t)b];
    BatchGeom g{B, rows, d_len, valid_rows};
    auto params = [&](int which) {
      ConvGemmParams p{};
      p.A = A; p.lda = Cin; p.a_bstride = (int64_t)rows * Cin;
      p.W = W; p.rows_per_frame = 1; p.N = N; p.Cin = Cin; p.taps = taps; p.dil = dil;
      p.bias = bias; p.act = ACT_NONE;
      p.lda_out = N; p.ao_bstride = (int64_t)rows * N;
      if (mode != 4) p.out_a = a[which];
      if (mode != 3 && mode != 4) { p.snake_ea = ea; p.snake_ib = ib; }
      if (mode == 1 || mode == 2 || mode == 4) { p.out_y = y[which]; p.ldy = N; p.y_bstride = (int64_t)rows * N; }
      if (mode == 1) { p.res = y[which]; p.ldres = N; p.res_bstride = (int64_t)rows * N; }
      return p;
    };
    for (int w = 0; w < 2; ++w) {
      CUDA_OK(cudaMemcpyAsync(y[w], X, R * N * 2, cudaMemcpyDeviceToDevice, s));
      CUDA_OK(cudaMemsetAsync(a[w], 0, R * N * 2, s));
    }
    ConvGemmParams p0 = params(0), p1 = params(1);
    launch_conv_gemm_simt(p0, g, op, op, s);
    if (!tc2_supported(p1, op)) return fail(Q3TTS_EINVAL, "shape not supported by the tcgen05 GEMM");
    CUDA_OK(launch_conv_gemm_tc2(p1, g, op, op, s));
    CUDA_OK(cudaStreamSynchronize(s));
    auto max_diff = [&](void* d0, void* d1) {
      std::vector<uint16_t> h0(R * N), h1(R * N);
      CUDA_OK(cudaMemcpy(h0.data(), d0, R * N * 2, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(h1.data(), d1, R * N * 2, cudaMemcpyDeviceToHost));
      auto tof = [&](uint16_t u) {
        if (op == DT_BF16) { uint32_t v = (uint32_t)u << 16; float f; std::memcpy(&f, &v, 4); return f; }
        const uint32_t sgn = (u >> 15) & 1, e = (u >> 10) & 31, m = u & 1023;
        float f = e == 0 ? ldexpf((float)m, -24) : (e == 31 ? INFINITY : ldexpf((float)(m | 1024), (int)e - 25));
        return sgn ? -f : f;
      };
      float worst = 0.f;
      for (int b = 0; b < B; ++b)
        for (int t = 0; t < len[(size_t)b]; ++t)
          for (int n = 0; n < N; ++n) {
            const size_t i = ((size_t)b * rows + t) * N + n;
            const float d = fabsf(tof(h0[i]) - tof(h1[i]));
            if (!(d <= worst)) worst = d;   // NaN-propagating
          }
      return worst;
    };
    if (max_diff_a) *max_diff_a = mode != 4 ? max_diff(a[0], a[1]) : 0.f;
    if (max_diff_y) *max_diff_y = (mode == 1 || mode == 2 || mode == 4) ? max_diff(y[0], y[1]) : 0.f;
    if (iters > 0 && ms_out) {
      cudaEvent_t e0, e1;
      CUDA_OK(cudaEventCreate(&e0));
      CUDA_OK(cudaEventCreate(&e1));
      for (int i = 0; i < 2; ++i) CUDA_OK(launch_conv_gemm_tc2(p1, g, op, op, s));
      CUDA_OK(cudaEventRecord(e0, s));
      for (int i = 0; i < iters; ++i) CUDA_OK(launch_conv_gemm_tc2(p1, g, op, op, s));
      CUDA_OK(cudaEventRecord(e1, s));
      CUDA_OK(cudaStreamSynchronize(s));
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      *ms_out = ms / iters;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    for (void* d : allocs) cudaFree(d);
    cudaStreamDestroy(s);
    return (int)Q3TTS_OK;
  });
}

// Fused residual unit (kernels_res96.cu) against the same unit composed from three CUDA-core GEMM launches
// (identity 1x1 + snake1, conv7 + snake2, conv1 + residual [+ snake3]) on seeded random data; then `iters` timed launches.
int q3tts_debug_resunit(int32_t B, int32_t rows, int32_t dil, int32_t out_snake, int32_t precision, int32_t iters,
                        float* ms_out, float* max_diff) {
  return guarded([&]() {
    if (B < 1 || rows < 1 || dil < 1) return fail(Q3TTS_EINVAL, "bad shape");
    const int op = precision == Q3TTS_PREC_BF16 ? DT_BF16 : DT_F16;
    const int C = 96;
    cudaStream_t s = nullptr;
    CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<void*> allocs;
    auto dmalloc = [&](size_t bytes) { void* d = nullptr; CUDA_OK(cudaMalloc(&d, bytes)); allocs.push_back(d); return d; };
    uint64_t seed = 0xD1B54A32D192ED03ull ^ (uint64_t)dil;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (float)((seed >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f; };
    auto upload16v = [&](const std::vector<float>& h) {
      float* d32 = (float*)dmalloc(h.size() * 4);
      void* d16 = dmalloc(h.size() * 2);
      CUDA_OK(cudaMemcpyAsync(d32, h.data(), h.size() * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      launch_convert(d32, d16, op, (int64_t)h.size(), s);
      return d16;
    };
    auto rand_vec = [&](size_t n, float scale, float offset) { std::vector<float> h(n); for (auto& v : h) v = offset + rnd() * scale; return h; };
    auto upload32 = [&](const std::vector<float>& h) {
      float* d = (float*)dmalloc(h.size() * 4);
      CUDA_OK(cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      return d;
    };
    const size_t R = (size_t)B * rows;
    void* X = upload16v(rand_vec(R * C, 1.5f, 0.f));
    void* W7 = upload16v(rand_vec((size_t)7 * C * C, 1.0f / sqrtf(7.0f * C), 0.f));
    void* W1 = upload16v(rand_vec((size_t)C * C, 1.0f / sqrtf((float)C), 0.f));
    std::vector<float> eye((size_t)C * C, 0.f);
    for (int i = 0; i < C; ++i) eye[(size_t)i * C + i] = 1.f;
    void* WI = upload16v(eye);
    float* zero = upload32(std::vector<float>((size_t)C, 0.f));
    float* b7 = upload32(rand_vec((size_t)C, 0.1f, 0.f));
    float* b1 = upload32(rand_vec((size_t)C, 0.1f, 0.f));
    float *ea[3], *ib[3];
    for (int i = 0; i < 3; ++i) { ea[i] = upload32(rand_vec((size_t)C, 0.3f, 1.0f)); ib[i] = upload32(rand_vec((size_t)C, 0.3f, 1.0f)); }
    std::vector<int> len((size_t)B);
    long long valid_rows = 0;
    for (int b = 0; b < B; ++b) { len[(size_t)b] = std::max(1, rows - 301 * b); valid_rows += len[(size_t)b]; }
    int* d_len = (int*)dmalloc((size_t)B * 4);
    CUDA_OK(cudaMemcpyAsync(d_len, len.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    void* A = dmalloc(R * C * 2);
    void* Cb = dmalloc(R * C * 2);
    void* Y = dmalloc(R * C * 2);       // reference stream (in place)
    void* A3 = dmalloc(R * C * 2);      // reference snake3 output
    void* O = dmalloc(R * C * 2);       // fused output
    CUDA_OK(cudaMemsetAsync(O, 0, R * C * 2, s));
    CUDA_OK(cudaMemsetAsync(A3, 0, R * C * 2, s));
    CUDA_OK(cudaMemcpyAsync(Y, X, R * C * 2, cudaMemcpyDeviceToDevice, s));
    BatchGeom g{B, rows, d_len, valid_rows};
    auto base = [&](const void* in, const void* w, int taps, int d, const float* bias) {
      ConvGemmParams p{};
      p.A = in; p.lda = C; p.a_bstride = (int64_t)rows * C; p.W = w; p.rows_per_frame = 1; p.N = C; p.Cin = C; p.taps = taps; p.dil = d;
      p.bias = bias; p.act = ACT_NONE; p.lda_out = C; p.ao_bstride = (int64_t)rows * C; p.ldy = C; p.y_bstride = (int64_t)rows * C;
      p.ldres = C; p.res_bstride = (int64_t)rows * C;
      return p;
    };
    { ConvGemmParams p = base(X, WI, 1, 1, zero); p.out_a = A; p.snake_ea = ea[0]; p.snake_ib = ib[0]; launch_conv_gemm_simt(p, g, op, op, s); }
    { ConvGemmParams p = base(A, W7, 7, dil, b7); p.out_a = Cb; p.snake_ea = ea[1]; p.snake_ib = ib[1]; launch_conv_gemm_simt(p, g, op, op, s); }
    { ConvGemmParams p = base(Cb, W1, 1, 1, b1); p.res = Y; p.out_y = Y; p.out_a = A3; p.snake_ea = ea[2]; p.snake_ib = ib[2]; launch_conv_gemm_simt(p, g, op, op, s); }
    ResUnitParams rp{};
    rp.x_in = X; rp.out = O; rp.w7 = W7; rp.w1 = W1; rp.b7 = b7; rp.b1 = b1;
    rp.ea1 = ea[0]; rp.ib1 = ib[0]; rp.ea2 = ea[1]; rp.ib2 = ib[1];
    if (out_snake) { rp.ea3 = ea[2]; rp.ib3 = ib[2]; }
    rp.C = C; rp.dil = dil; rp.rows_per_frame = 1;
    if (!resunit96_supported(rp, op)) return fail(Q3TTS_EINVAL, "fused residual unit not supported for this shape");
    long long* d_dbg = nullptr;
    if (getenv("Q3TTS_RES_DEBUG")) { d_dbg = (long long*)dmalloc(16 * 32 * 8); CUDA_OK(cudaMemsetAsync(d_dbg, 0, 16 * 32 * 8, s)); rp.dbg = d_dbg; }
    CUDA_OK(launch_resunit96(rp, g, op, s));
    CUDA_OK(cudaStreamSynchronize(s));
    if (d_dbg) {
      std::vector<long long> h(16 * 32);
      CUDA_OK(cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost));
      static const char* kEv[16] = {"prod:slot_free", "P:before_wait", "P:t_full", "P:done", "MMA:a_ready", "MMA:c7_issued", "MMA:c_ready", "MMA:c1_issued",
                                    "E1:acc1_full", "E1:done", "E2:acc2_full", "E2:done", "P2:done", "P3:done", "P4:done", "P5:done"};
      const long long t0 = h[0 * 32 + 8];
      for (int tile = 8; tile < 20; ++tile) {
        std::printf("tile %2d:", tile);
        for (int ev = 0; ev < 16; ++ev) std::printf(" %s=%lld", kEv[ev], h[(size_t)ev * 32 + tile] - t0);
        std::printf("\n");
      }
      rp.dbg = nullptr;
    }
    {
      std::vector<uint16_t> h0(R * C), h1(R * C);
      CUDA_OK(cudaMemcpy(h0.data(), out_snake ? A3 : Y, R * C * 2, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(h1.data(), O, R * C * 2, cudaMemcpyDeviceToHost));
      auto tof = [&](uint16_t u) {
        if (op == DT_BF16) { uint32_t v = (uint32_t)u << 16; float f; std::memcpy(&f, &v, 4); return f; }
        const uint32_t sgn = (u >> 15) & 1, e = (u >> 10) & 31, m = u & 1023;
        float f = e == 0 ? ldexpf((float)m, -24) : (e == 31 ? INFINITY : ldexpf((float)(m | 1024), (int)e - 25));
        return sgn ? -f : f;
      };
      float worst = 0.f;
      for (int b = 0; b < B; ++b)
        for (int t = 0; t < len[(size_t)b]; ++t)
          for (int n = 0; n < C; ++n) {
            const size_t i = ((size_t)b * rows + t) * C + n;
            const float d = fabsf(tof(h0[i]) - tof(h1[i]));
            if (!(d <= worst)) worst = d;
          }
      if (max_diff) *max_diff = worst;
    }
    if (iters > 0 && ms_out) {
      cudaEvent_t e0, e1;
      CUDA_OK(cudaEventCreate(&e0));
      CUDA_OK(cudaEventCreate(&e1));
      for (int i = 0; i < 2; ++i) CUDA_OK(launch_resunit96(rp, g, op, s));
      CUDA_OK(cudaEventRecord(e0, s));
      for (int i = 0; i < iters; ++i) CUDA_OK(launch_resunit96(rp, g, op, s));
      CUDA_OK(cudaEventRecord(e1, s));
      CUDA_OK(cudaStreamSynchronize(s));
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      *ms_out = ms / iters;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    for (void* d : allocs) cudaFree(d);
    cudaStreamDestroy(s);
    return (int)Q3TTS_OK;
  });
}

// The residual unit fused into the tcgen05 GEMM (conv7 + snake -> smem operand tile -> conv1 + residual [+ snake], kernels_tc2.cu)
// against the same unit as two CUDA-core GEMM launches, on seeded random data; then `iters` timed launches.
int q3tts_debug_fused_unit(int32_t B, int32_t rows, int32_t C, int32_t dil, int32_t with_operand, int32_t precision, int32_t iters,
                           float* ms_out, float* max_diff_y, float* max_diff_a) {
  return guarded([&]() {
    if (B < 1 || rows < 1 || C < 64 || dil < 1) return fail(Q3TTS_EINVAL, "bad shape");
    const int op = precision == Q3TTS_PREC_BF16 ? DT_BF16 : DT_F16;
    cudaStream_t s = nullptr;
    CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<void*> allocs;
    auto dmalloc = [&](size_t bytes) { void* d = nullptr; CUDA_OK(cudaMalloc(&d, bytes)); allocs.push_back(d); return d; };
    uint64_t seed = 0xA0761D6478BD642Full ^ ((uint64_t)C << 20) ^ (uint64_t)dil;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (float)((seed >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f; };
    auto upload16 = [&](size_t n, float scale) {
      std::vector<float> h(n);
      for (auto& v : h) v = rnd() * scale;
      float* d32 = (float*)dmalloc(n * 4);
      void* d16 = dmalloc(n * 2);
      CUDA_OK(cudaMemcpyAsync(d32, h.data(), n * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      launch_convert(d32, d16, op, (int64_t)n, s);
      return d16;
    };
    auto upload32 = [&](size_t n, float scale, float offset) {
      std::vector<float> h(n);
      for (auto& v : h) v = offset + rnd() * scale;
      float* d = (float*)dmalloc(n * 4);
      CUDA_OK(cudaMemcpyAsync(d, h.data(), n * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      return d;
    };
    const size_t R = (size_t)B * rows;
    void* A = upload16(R * C, 1.5f);
    void* X = upload16(R * C, 1.5f);
    void* W7 = upload16((size_t)7 * C * C, 1.0f / sqrtf(7.0f * C));
    void* W1 = upload16((size_t)C * C, 1.0f / sqrtf((float)C));
    float* b7 = upload32((size_t)C, 0.1f, 0.f);
    float* b1 = upload32((size_t)C, 0.1f, 0.f);
    float *ea[2], *ib[2];
    for (int i = 0; i < 2; ++i) { ea[i] = upload32((size_t)C, 0.3f, 1.0f); ib[i] = upload32((size_t)C, 0.3f, 1.0f); }
    std::vector<int> len((size_t)B);
    long long valid_rows = 0;
    for (int b = 0; b < B; ++b) { len[(size_t)b] = std::max(1, rows - 211 * b); valid_rows += len[(size_t)b]; }
    int* d_len = (int*)dmalloc((size_t)B * 4);
    CUDA_OK(cudaMemcpyAsync(d_len, len.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    void* Cb = dmalloc(R * C * 2);
    void* y[2] = {dmalloc(R * C * 2), dmalloc(R * C * 2)};
    void* a[2] = {dmalloc(R * C * 2), dmalloc(R * C * 2)};
    for (int w = 0; w < 2; ++w) {
      CUDA_OK(cudaMemcpyAsync(y[w], X, R * C * 2, cudaMemcpyDeviceToDevice, s));
      CUDA_OK(cudaMemsetAsync(a[w], 0, R * C * 2, s));
    }
    BatchGeom g{B, rows, d_len, valid_rows};
    auto base = [&](const void* in, const void* w, int taps, int d, const float* bias) {
      ConvGemmParams p{};
      p.A = in; p.lda = C; p.a_bstride = (int64_t)rows * C; p.W = w; p.rows_per_frame = 1; p.N = C; p.Cin = C; p.taps = taps; p.dil = d;
      p.bias = bias; p.act = ACT_NONE; p.lda_out = C; p.ao_bstride = (int64_t)rows * C; p.ldy = C; p.y_bstride = (int64_t)rows * C;
      p.ldres = C; p.res_bstride = (int64_t)rows * C;
      return p;
    };
    { ConvGemmParams p = base(A, W7, 7, dil, b7); p.out_a = Cb; p.snake_ea = ea[0]; p.snake_ib = ib[0]; launch_conv_gemm_simt(p, g, op, op, s); }
    { ConvGemmParams p = base(Cb, W1, 1, 1, b1); p.res = y[0]; p.out_y = y[0];
      if (with_operand) { p.out_a = a[0]; p.snake_ea = ea[1]; p.snake_ib = ib[1]; }
      launch_conv_gemm_simt(p, g, op, op, s); }
    ConvGemmParams pf = base(A, W7, 7, dil, b7);
    pf.snake_ea = ea[0]; pf.snake_ib = ib[0];
    // with_operand: 0 = stream only, 1 = stream + next operand, 2 = operand only (a block's last unit: the stream is not written)
    FusedConv1 f{W1, b1, y[1], with_operand == 2 ? nullptr : y[1], with_operand ? a[1] : nullptr, with_operand ? ea[1] : nullptr, with_operand ? ib[1] : nullptr};
    if (!tc2_fuse_supported(pf, op)) return fail(Q3TTS_EINVAL, "shape not supported by the fused unit");
    CUDA_OK(launch_conv_gemm_tc2(pf, g, op, op, s, &f));
    CUDA_OK(cudaStreamSynchronize(s));
    auto max_diff = [&](void* d0, void* d1) {
      std::vector<uint16_t> h0(R * C), h1(R * C);
      CUDA_OK(cudaMemcpy(h0.data(), d0, R * C * 2, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(h1.data(), d1, R * C * 2, cudaMemcpyDeviceToHost));
      auto tof = [&](uint16_t u) {
        if (op == DT_BF16) { uint32_t v = (uint32_t)u << 16; float f; std::memcpy(&f, &v, 4); return f; }
        const uint32_t sgn = (u >> 15) & 1, e = (u >> 10) & 31, m = u & 1023;
        float f = e == 0 ? ldexpf((float)m, -24) : (e == 31 ? INFINITY : ldexpf((float)(m | 1024), (int)e - 25));
        return sgn ? -f : f;
      };
      float worst = 0.f;
      for (int b = 0; b < B; ++b)
        for (int t = 0; t < len[(size_t)b]; ++t)
          for (int n = 0; n < C; ++n) {
            const size_t i = ((size_t)b * rows + t) * C + n;
            const float d = fabsf(tof(h0[i]) - tof(h1[i]));
            if (!(d <= worst)) worst = d;
          }
      return worst;
    };
    if (max_diff_y) *max_diff_y = max_diff(with_operand == 2 ? X : y[0], y[1]);   // operand only: the stream buffer must be untouched
    if (max_diff_a) *max_diff_a = with_operand ? max_diff(a[0], a[1]) : 0.f;
    if (iters > 0 && ms_out) {
      cudaEvent_t e0, e1;
      CUDA_OK(cudaEventCreate(&e0));
      CUDA_OK(cudaEventCreate(&e1));
      for (int i = 0; i < 2; ++i) CUDA_OK(launch_conv_gemm_tc2(pf, g, op, op, s, &f));
      CUDA_OK(cudaEventRecord(e0, s));
      for (int i = 0; i < iters; ++i) CUDA_OK(launch_conv_gemm_tc2(pf, g, op, op, s, &f));
      CUDA_OK(cudaEventRecord(e1, s));
      CUDA_OK(cudaStreamSynchronize(s));
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      *ms_out = ms / iters;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    for (void* d : allocs) cudaFree(d);
    cudaStreamDestroy(s);
    return (int)Q3TTS_OK;
  });
}

// The attention kernel on its own (ST.swift:519-525: softmax(scale * Q K^T [+ mask]) V per utterance and head) on caller-provided
// q | k | v rows: the production dispatch (launch_attention: the mma.sync kernel for 16-bit head_dim 64, the fp32 kernel otherwise).
// qkv: host float [B, T, (nh + 2 nkv) * hd], rounded to the precision's operand type on the device; out: host float [B, T, nh * hd].
// len / row_begin (may be NULL): valid rows [row_begin[b], len[b]) of each slot; window 0 = full attention.
int q3tts_debug_attention(const float* qkv, int32_t B, int32_t T, int32_t nh, int32_t nkv, int32_t hd, const int32_t* len,
                          const int32_t* row_begin, int32_t window, int32_t precision, float* out) {
  return guarded([&]() {
    if (!qkv || !out || B < 1 || T < 1 || nh < 1 || nkv < 1 || nh % nkv || (hd != 32 && hd != 64 && hd != 128) || window < 0)
      return fail(Q3TTS_EINVAL, "bad attention shape");
    if (precision < Q3TTS_PREC_FP32 || precision > Q3TTS_PREC_BF16) return fail(Q3TTS_EINVAL, "bad precision");
    const int op = precision == Q3TTS_PREC_FP32 ? DT_F32 : (precision == Q3TTS_PREC_FP16 ? DT_F16 : DT_BF16);
    const size_t es = op == DT_F32 ? 4 : 2;
    const size_t ld = (size_t)(nh + 2 * nkv) * hd, n_in = (size_t)B * T * ld, n_out = (size_t)B * T * nh * hd;
    cudaStream_t s = nullptr;
    CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<void*> allocs;
    auto dmalloc = [&](size_t bytes) { void* d = nullptr; CUDA_OK(cudaMalloc(&d, bytes)); allocs.push_back(d); return d; };
    float* d32 = (float*)dmalloc(n_in * 4);
    CUDA_OK(cudaMemcpyAsync(d32, qkv, n_in * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaStreamSynchronize(s));
    void* d_in = d32;
    if (op != DT_F32) { d_in = dmalloc(n_in * 2); launch_convert(d32, d_in, op, (int64_t)n_in, s); }
    void* d_out = dmalloc(n_out * es);
    CUDA_OK(cudaMemsetAsync(d_out, 0, n_out * es, s));
    std::vector<int> hl((size_t)B, T);
    long long valid = 0;
    for (int b = 0; b < B; ++b) {
      if (len) hl[(size_t)b] = len[b];
      if (hl[(size_t)b] < 0 || hl[(size_t)b] > T) return fail(Q3TTS_EINVAL, "len out of range");
      valid += hl[(size_t)b];
    }
    int* d_len = (int*)dmalloc((size_t)B * 4);
    CUDA_OK(cudaMemcpyAsync(d_len, hl.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
    int* d_beg = nullptr;
    if (row_begin) {
      d_beg = (int*)dmalloc((size_t)B * 4);
      CUDA_OK(cudaMemcpyAsync(d_beg, row_begin, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    }
    BatchGeom g{B, T, d_len, valid, d_beg};
    launch_attention(d_in, op, d_out, op, g, nh, nkv, hd, 1.0f / sqrtf((float)hd), window, s);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s));
    if (op == DT_F32) {
      CUDA_OK(cudaMemcpy(out, d_out, n_out * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<uint16_t> h(n_out);
      CUDA_OK(cudaMemcpy(h.data(), d_out, n_out * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n_out; ++i) {
        const uint16_t u = h[i];
        if (op == DT_BF16) { const uint32_t v = (uint32_t)u << 16; std::memcpy(&out[i], &v, 4); continue; }
        const uint32_t sgn = (u >> 15) & 1, e = (u >> 10) & 31, mant = u & 1023;
        const float f = e == 0 ? ldexpf((float)mant, -24) : (e == 31 ? (mant ? NAN : INFINITY) : ldexpf((float)(mant | 1024), (int)e - 25));
        out[i] = sgn ? -f : f;
      }
    }
    for (void* d : allocs) cudaFree(d);
    cudaStreamDestroy(s);
    return (int)Q3TTS_OK;
  });
}

// ---- codec-embedding sum --------------------------------------------------------------------------
struct q3tts_codec_embedder {
  int device = 0, dtype = DT_BF16, groups = 0;
  int64_t hidden = 0;
  std::vector<int> vocab;
  std::vector<void*> d_tables;
  int* d_err = nullptr;
  int32_t* d_codes = nullptr; size_t d_codes_cap = 0;
  void* d_out = nullptr;      size_t d_out_cap = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  ~q3tts_codec_embedder() {
    cudaSetDevice(device);
    for (void* t : d_tables) cudaFree(t);
    if (d_err) cudaFree(d_err);
    if (d_codes) cudaFree(d_codes);
    if (d_out) cudaFree(d_out);
    if (stream) cudaStreamDestroy(stream);
  }
};

int q3tts_codec_embedder_load(const char* model_dir, int32_t device, q3tts_codec_embedder** out) {
  return guarded([&]() {
    if (!model_dir || !out) return fail(Q3TTS_EINVAL, "NULL argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(Q3TTS_ECUDA, "no CUDA device (this library has no CPU path)");
    CodecEmbeddingTables host;
    load_codec_embeddings(model_dir, &host);
    if ((int)host.tables.size() > kMaxCodeGroups) return fail(Q3TTS_EFORMAT, "too many code groups");
    if (host.hidden % 8) return fail(Q3TTS_EFORMAT, "hidden size must be a multiple of 8");
    std::unique_ptr<q3tts_codec_embedder> e(new q3tts_codec_embedder());
    e->device = device < 0 ? 0 : device;
    if (e->device >= ndev) return fail(Q3TTS_EINVAL, "device index out of range");
    CUDA_OK(cudaSetDevice(e->device));
    e->dtype = host.dtype == "F32" ? DT_F32 : (host.dtype == "F16" ? DT_F16 : DT_BF16);
    e->hidden = host.hidden;
    e->groups = (int)host.tables.size();
    CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&e->d_err, 4));
    CUDA_OK(cudaMemsetAsync(e->d_err, 0, 4, e->stream));
    for (auto& t : host.tables) {
      const int64_t n = t.numel();
      float* d32 = nullptr;
      CUDA_OK(cudaMalloc(&d32, (size_t)n * 4));
      CUDA_OK(cudaMemcpyAsync(d32, t.data.data(), (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
      void* d = d32;
      if (e->dtype != DT_F32) {   // the fp32 copies are exact images of the 16-bit values on disk: converting back is lossless
        CUDA_OK(cudaMalloc(&d, (size_t)n * 2));
        launch_convert(d32, d, e->dtype, n, e->stream);
        CUDA_OK(cudaStreamSynchronize(e->stream));
        cudaFree(d32);
      }
      e->d_tables.push_back(d);
      e->vocab.push_back((int)t.shape[0]);
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    *out = e.release();
    return (int)Q3TTS_OK;
  });
}

void q3tts_codec_embedder_free(q3tts_codec_embedder* e) { delete e; }

int q3tts_codec_embedder_info(const q3tts_codec_embedder* e, int32_t* hidden, int32_t* groups, int32_t* precision, int32_t* vocab) {
  if (!e) return fail(Q3TTS_EINVAL, "embedder is NULL");
  if (hidden) *hidden = (int32_t)e->hidden;
  if (groups) *groups = e->groups;
  if (precision) *precision = e->dtype == DT_F32 ? Q3TTS_PREC_FP32 : (e->dtype == DT_F16 ? Q3TTS_PREC_FP16 : Q3TTS_PREC_BF16);
  if (vocab) for (int i = 0; i < e->groups; ++i) vocab[i] = e->vocab[(size_t)i];
  return Q3TTS_OK;
}

static int codec_embed_common(q3tts_codec_embedder* e, const int32_t* codes, int64_t n, void* out, bool device_ptrs, cudaStream_t user) {
  return guarded([&]() {
    if (!e) return fail(Q3TTS_EINVAL, "embedder is NULL");
    if (n < 0) return fail(Q3TTS_EINVAL, "negative frame count");
    if (n == 0) return (int)Q3TTS_OK;
    if (!codes || !out) return fail(Q3TTS_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lock(e->mu);
    CUDA_OK(cudaSetDevice(e->device));
    cudaStream_t s = device_ptrs ? user : e->stream;
    const size_t es = e->dtype == DT_F32 ? 4 : 2, out_bytes = (size_t)n * e->hidden * es, code_bytes = (size_t)n * e->groups * 4;
    CodecEmbedParams p{};
    for (int g = 0; g < e->groups; ++g) { p.tables[g] = e->d_tables[(size_t)g]; p.vocab[g] = e->vocab[(size_t)g]; }
    p.groups = e->groups; p.H = (int)e->hidden; p.dtype = e->dtype; p.n = n; p.err_flag = e->d_err;
    if (device_ptrs) {
      p.codes = codes; p.out = out;
    } else {
      if (code_bytes > e->d_codes_cap) { if (e->d_codes) cudaFree(e->d_codes); e->d_codes = nullptr; CUDA_OK(cudaMalloc(&e->d_codes, code_bytes)); e->d_codes_cap = code_bytes; }
      if (out_bytes > e->d_out_cap) { if (e->d_out) cudaFree(e->d_out); e->d_out = nullptr; CUDA_OK(cudaMalloc(&e->d_out, out_bytes)); e->d_out_cap = out_bytes; }
      CUDA_OK(cudaMemcpyAsync(e->d_codes, codes, code_bytes, cudaMemcpyHostToDevice, s));
      p.codes = e->d_codes; p.out = e->d_out;
    }
    launch_codec_embed_sum(p, s);
    CUDA_OK(cudaGetLastError());
    if (!device_ptrs) {
      int flag = 0;
      CUDA_OK(cudaMemcpyAsync(out, e->d_out, out_bytes, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaMemcpyAsync(&flag, e->d_err, 4, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaStreamSynchronize(s));
      if (flag) {
        CUDA_OK(cudaMemsetAsync(e->d_err, 0, 4, s));
        return fail(Q3TTS_EINVAL, "a code id is outside its embedding table");
      }
    }
    return (int)Q3TTS_OK;
  });
}

int q3tts_codec_embed_sum(q3tts_codec_embedder* e, const int32_t* codes, int64_t n_frames, void* out) {
  return codec_embed_common(e, codes, n_frames, out, false, nullptr);
}

int q3tts_codec_embed_sum_device(q3tts_codec_embedder* e, const int32_t* d_codes, int64_t n_frames, void* d_out, void* stream) {
  return codec_embed_common(e, d_codes, n_frames, d_out, true, (cudaStream_t)stream);
}

// ---- measurement --------------------------------------------------------------------------------------------
int q3tts_profile_enable(q3tts_model* h, int32_t enable) {
  if (!h) return fail(Q3TTS_EINVAL, "model is NULL");
  std::lock_guard<std::mutex> lock(h->m->mu);
  h->m->profile_enabled = enable != 0;
  return Q3TTS_OK;
}

int q3tts_profile_get(q3tts_model* h, q3tts_stage_time* out, int32_t cap) {
  if (!h || (!out && cap > 0)) return 0;
  std::lock_guard<std::mutex> lock(h->m->mu);
  int n = 0;
  for (auto& sp : h->m->prof) {
    if (!sp.used) continue;
    if (n < cap) {
      std::memset(&out[n], 0, sizeof(out[n]));
      std::snprintf(out[n].name, sizeof(out[n].name), "%s", sp.name.c_str());
      out[n].ms = sp.ms_accum;
      out[n].launches = sp.launches;
      out[n].flops = sp.flops;
      out[n].bytes = sp.bytes;
    }
    ++n;
  }
  return n;
}

int q3tts_profile_kernels(q3tts_model* h, q3tts_kernel_time* out, int32_t cap) {
  if (!h || (!out && cap > 0)) return 0;
  std::lock_guard<std::mutex> lock(h->m->mu);
  int n = 0;
  for (auto& kv : h->m->kernel_totals) {
    if (n < cap) {
      std::memset(&out[n], 0, sizeof(out[n]));
      std::snprintf(out[n].name, sizeof(out[n].name), "%s", kv.first.c_str());
      out[n].ms = (float)kv.second[0];
      out[n].launches = (int32_t)kv.second[1];
      out[n].flops = kv.second[2];
      out[n].bytes = kv.second[3];
    }
    ++n;
  }
  return n;
}

int64_t q3tts_launch_count(const q3tts_model* h) { return h ? h->m->launches : 0; }

}  // extern "C"
