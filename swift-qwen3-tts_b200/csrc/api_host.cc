// Host-only entry points of the C ABI: no CUDA headers, no device.  Built into libqwen3tts_cuda.so and, together with checkpoint.cc,
// into the AddressSanitizer / UBSan harness (make asan; tests/test_sanitizers.py).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/qwen3tts_cuda.h"
#include "checkpoint.hpp"

namespace q3api {
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
}  // namespace q3api
using namespace q3api;

extern "C" {

int q3tts_abi_version(void) { return Q3TTS_ABI_VERSION; }
const char* q3tts_last_error(void) { return g_last_error.c_str(); }

int q3tts_checkpoint_inspect(const char* dir, q3tts_config* cfg) {
  try {
    if (!dir) return fail(Q3TTS_EINVAL, "speech_tokenizer_dir is NULL");
    q3::Checkpoint ck;
    load_checkpoint(dir, &ck);
    if (cfg) *cfg = ck.cfg;
    return (int)Q3TTS_OK;
  } catch (const q3::Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(Q3TTS_ENOMEM, "host allocation failed");
  } catch (const std::exception& e) {
    return fail(Q3TTS_EINVAL, e.what());
  }
}


// ---- scheduler ------------------------------------------------------------------------------------------
int q3tts_partition_lpt(const int64_t* frames, int32_t n, int32_t parts, int32_t* part_out) {
  if (n < 0 || parts <= 0 || (n > 0 && (!frames || !part_out))) return fail(Q3TTS_EINVAL, "bad partition arguments");
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return frames[a] > frames[b]; });
  std::vector<int64_t> load((size_t)parts, 0);
  for (int32_t idx : order) {
    int best = 0;
    for (int p = 1; p < parts; ++p)
      if (load[(size_t)p] < load[(size_t)best]) best = p;
    part_out[idx] = best;
    load[(size_t)best] += std::max<int64_t>(frames[idx], 0);
  }
  return Q3TTS_OK;
}

// ---- PCM post-processing ----------------------------------------------------------------------------------
int64_t q3tts_trim_length(int64_t n, int64_t valid) { return (valid > 0 && valid < n) ? valid : n; }

int64_t q3tts_voice_clone_cut(int64_t ref_frames, int64_t total_frames, int64_t n) {
  const float cutf = (float)ref_frames / (float)std::max<int64_t>(total_frames, 1) * (float)n;   // Q3.swift:1196
  const int64_t cut = (int64_t)cutf;
  return (cut > 0 && cut < n) ? cut : 0;
}

int q3tts_pcm_to_int16(const float* pcm, int64_t n, int16_t* out) {
  if (n < 0 || (n > 0 && (!pcm || !out))) return fail(Q3TTS_EINVAL, "bad arguments");
  for (int64_t i = 0; i < n; ++i) {
    const float c = std::max(-1.0f, std::min(1.0f, pcm[i]));
    out[i] = (int16_t)(c * 32767.0f);
  }
  return Q3TTS_OK;
}

int q3tts_write_wav(const char* path, const float* pcm, int64_t n, int32_t rate) {
  if (!path || n < 0 || (n > 0 && !pcm) || rate <= 0) return fail(Q3TTS_EINVAL, "bad arguments");
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(Q3TTS_EIO, std::string("cannot open ") + path);
  const uint32_t data = (uint32_t)(n * 2), riff = 36 + data, fmt = 16, byte_rate = (uint32_t)rate * 2, sr = (uint32_t)rate;
  const uint16_t pcm_fmt = 1, ch = 1, align = 2, bits = 16;
  std::fwrite("RIFF", 1, 4, f); std::fwrite(&riff, 4, 1, f); std::fwrite("WAVE", 1, 4, f);
  std::fwrite("fmt ", 1, 4, f); std::fwrite(&fmt, 4, 1, f); std::fwrite(&pcm_fmt, 2, 1, f); std::fwrite(&ch, 2, 1, f);
  std::fwrite(&sr, 4, 1, f); std::fwrite(&byte_rate, 4, 1, f); std::fwrite(&align, 2, 1, f); std::fwrite(&bits, 2, 1, f);
  std::fwrite("data", 1, 4, f); std::fwrite(&data, 4, 1, f);
  std::vector<int16_t> buf((size_t)n);
  q3tts_pcm_to_int16(pcm, n, buf.data());
  const size_t wrote = std::fwrite(buf.data(), 2, (size_t)n, f);
  std::fclose(f);
  return wrote == (size_t)n ? Q3TTS_OK : fail(Q3TTS_EIO, "short write");
}

}  // extern "C"
