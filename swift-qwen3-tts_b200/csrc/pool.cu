// Multi-GPU pool (north_star (d)): one worker thread + CUDA context + model replica per GPU of one box.
// The path shards by utterance -- an utterance's PCM depends only on its own codes and the replicated read-only weights
// (ST.swift:754-784 is batch-elementwise) -- so there is no collective: the caller's host buffers are the gather.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <thread>

#include "api_internal.hpp"

using namespace q3;
using namespace q3api;

namespace {
struct Job {
  const int32_t* codes = nullptr;
  int32_t layout = Q3TTS_CODES_BTQ, T_uniform = 0;
  std::vector<HostUtt> utts;
  void* pcm = nullptr;
  bool i16 = false;
  int32_t* lengths = nullptr;
};

struct Worker {
  int device = 0;
  Model* model = nullptr;
  std::thread thread;
  std::mutex mu;
  std::condition_variable cv;
  bool has_job = false, done = false, quit = false;
  Job job;
  int status = Q3TTS_OK;
  std::string message;
  float ms = 0.f;
  int64_t frames = 0;

  void loop() {
    cudaSetDevice(device);                      // this thread's current device for the worker's life
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_job || quit; });
        if (quit) return;
      }
      const auto t0 = std::chrono::steady_clock::now();
      int64_t fr = 0;
      for (const HostUtt& u : job.utts) fr += u.frames;
      const int st = guarded([&]() {
        decode_host_list(*model, job.codes, job.layout, job.T_uniform, std::move(job.utts), job.pcm, job.i16, job.lengths);
        return (int)Q3TTS_OK;
      });
      const float elapsed = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
      {
        std::lock_guard<std::mutex> lk(mu);
        status = st;
        message = st == Q3TTS_OK ? std::string() : g_last_error;   // thread-local of THIS worker: hand it to the caller
        ms = elapsed; frames = fr;
        has_job = false; done = true;
      }
      cv.notify_all();
    }
  }
};
}  // namespace

struct q3tts_pool {
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex call_mu;                           // one pool call at a time
  q3tts_config cfg{};
  ~q3tts_pool() {
    for (auto& w : workers) {
      { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
      w->cv.notify_all();
      if (w->thread.joinable()) w->thread.join();
      delete w->model;
    }
  }
};

// Hand every worker its share and wait for all of them.
static int run_jobs(q3tts_pool* p, std::vector<Job>& jobs) {
  const int W = (int)p->workers.size();
  for (int k = 0; k < W; ++k) {
    Worker& w = *p->workers[(size_t)k];
    std::lock_guard<std::mutex> lk(w.mu);
    w.job = std::move(jobs[(size_t)k]);
    w.done = false; w.has_job = true;
  }
  for (auto& w : p->workers) w->cv.notify_all();
  int status = Q3TTS_OK;
  std::string msg;
  for (int k = 0; k < W; ++k) {
    Worker& w = *p->workers[(size_t)k];
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.status != Q3TTS_OK && status == Q3TTS_OK) { status = w.status; msg = "GPU " + std::to_string(w.device) + ": " + w.message; }
  }
  return status == Q3TTS_OK ? (int)Q3TTS_OK : fail(status, msg);
}

static int pool_decode_list(q3tts_pool* p, const int32_t* codes, int32_t layout, int32_t T_uniform, const std::vector<HostUtt>& utts,
                            void* pcm, bool i16, int32_t* lengths) {
  std::lock_guard<std::mutex> call(p->call_mu);
  const int W = (int)p->workers.size(), n = (int)utts.size();
  std::vector<int64_t> frames((size_t)n);
  for (int i = 0; i < n; ++i) frames[(size_t)i] = utts[(size_t)i].frames;
  std::vector<int32_t> part((size_t)n, 0);
  if (n > 0) {
    const int st = q3tts_partition_lpt(frames.data(), n, W, part.data());
    if (st != Q3TTS_OK) return st;
  }
  std::vector<Job> jobs((size_t)W);
  for (auto& j : jobs) { j.codes = codes; j.layout = layout; j.T_uniform = T_uniform; j.pcm = pcm; j.i16 = i16; j.lengths = lengths; }
  for (int i = 0; i < n; ++i) jobs[(size_t)part[(size_t)i]].utts.push_back(utts[(size_t)i]);
  return run_jobs(p, jobs);
}

extern "C" {

int q3tts_pool_open(const char* dir, const q3tts_options* opts, const int32_t* devices, int32_t n_devices, q3tts_pool** out) {
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
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      return fail(Q3TTS_ECUDA, "no CUDA device: libqwen3tts_cuda has no CPU fallback");
    }
    std::vector<int> devs;
    if (devices) {
      if (n_devices < 1) return fail(Q3TTS_EINVAL, "n_devices must be >= 1");
      for (int i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(Q3TTS_EINVAL, "device ordinal out of range");
        for (int j = 0; j < i; ++j) if (devices[j] == devices[i]) return fail(Q3TTS_EINVAL, "a device appears twice");
        devs.push_back(devices[i]);
      }
    } else {
      for (int i = 0; i < ndev; ++i) {
        cudaDeviceProp pr{};
        if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) devs.push_back(i);
      }
      if (n_devices > 0 && (int)devs.size() > n_devices) devs.resize((size_t)n_devices);
      if (devs.empty()) return fail(Q3TTS_ECUDA, "no sm_100 device");
    }
    Checkpoint ck;
    load_checkpoint(dir, &ck);                  // parsed once, uploaded to every device
    std::unique_ptr<q3tts_pool> p(new q3tts_pool());
    p->cfg = ck.cfg;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d : devs) {
      std::unique_ptr<Worker> w(new Worker());
      w->device = d;
      o.device = d;
      w->model = model_create(ck, o);
      p->workers.push_back(std::move(w));
    }
    cudaSetDevice(prev);
    for (auto& w : p->workers) w->thread = std::thread([wp = w.get()] { wp->loop(); });
    *out = p.release();
    return (int)Q3TTS_OK;
  });
}

void q3tts_pool_close(q3tts_pool* p) { delete p; }

int32_t q3tts_pool_size(const q3tts_pool* p) { return p ? (int32_t)p->workers.size() : 0; }

static int pool_varlen(q3tts_pool* p, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n, void* pcm_out,
                       int32_t* lengths_out, bool i16) {
  return guarded([&]() {
    if (!p) return fail(Q3TTS_EINVAL, "pool is NULL");
    if (n < 0) return fail(Q3TTS_EINVAL, "negative utterance count");
    if (n == 0) return (int)Q3TTS_OK;
    if (!frame_offsets) return fail(Q3TTS_EINVAL, "frame_offsets is NULL");
    if (frame_offsets[0] != 0) return fail(Q3TTS_EINVAL, "frame_offsets[0] must be 0");
    for (int i = 0; i < n; ++i)
      if (frame_offsets[i + 1] < frame_offsets[i] || frame_offsets[i + 1] - frame_offsets[i] > INT32_MAX)
        return fail(Q3TTS_EINVAL, "frame_offsets must be non-decreasing");
    if (frame_offsets[n] > 0 && (!codes_packed || !pcm_out)) return fail(Q3TTS_EINVAL, "NULL buffer");
    const int Q = p->cfg.num_quantizers;
    const int64_t up = p->cfg.total_upsample;
    std::vector<HostUtt> utts((size_t)n);
    for (int i = 0; i < n; ++i)
      utts[(size_t)i] = HostUtt{frame_offsets[i] * Q, frame_offsets[i] * up, (int)(frame_offsets[i + 1] - frame_offsets[i]), i};
    return pool_decode_list(p, codes_packed, Q3TTS_CODES_BTQ, 0, utts, pcm_out, i16, lengths_out);
  });
}

int q3tts_pool_decode_varlen(q3tts_pool* p, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n, float* pcm_out,
                             int32_t* lengths_out) {
  return pool_varlen(p, codes_packed, frame_offsets, n, pcm_out, lengths_out, false);
}

int q3tts_pool_decode_varlen_int16(q3tts_pool* p, const int32_t* codes_packed, const int64_t* frame_offsets, int32_t n, int16_t* pcm_out,
                                   int32_t* lengths_out) {
  return pool_varlen(p, codes_packed, frame_offsets, n, pcm_out, lengths_out, true);
}

int q3tts_pool_decode(q3tts_pool* p, const int32_t* codes, int32_t B, int32_t T, int32_t layout, float* pcm_out, int32_t* lengths_out) {
  return guarded([&]() {
    if (!p) return fail(Q3TTS_EINVAL, "pool is NULL");
    if (B < 0 || T < 0) return fail(Q3TTS_EINVAL, "negative B or T");
    if (layout != Q3TTS_CODES_BQT && layout != Q3TTS_CODES_BTQ) return fail(Q3TTS_EINVAL, "bad codes layout");
    if (B == 0 || T == 0) return (int)Q3TTS_OK;
    if (!codes || !pcm_out) return fail(Q3TTS_EINVAL, "NULL buffer");
    const int Q = p->cfg.num_quantizers;
    const int64_t up = p->cfg.total_upsample;
    std::vector<HostUtt> utts((size_t)B);
    for (int b = 0; b < B; ++b) utts[(size_t)b] = HostUtt{(int64_t)b * T * Q, (int64_t)b * T * up, T, b};
    return pool_decode_list(p, codes, layout, T, utts, pcm_out, false, lengths_out);
  });
}

int q3tts_pool_last_stats(const q3tts_pool* p, float* ms_out, int64_t* frames_out, int32_t cap) {
  if (!p) return 0;
  const int W = (int)p->workers.size();
  for (int k = 0; k < W && k < cap; ++k) {
    if (ms_out) ms_out[k] = p->workers[(size_t)k]->ms;
    if (frames_out) frames_out[k] = p->workers[(size_t)k]->frames;
  }
  return W;
}

}  // extern "C"
