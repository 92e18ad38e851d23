// C ABI of the speech-tokenizer encoder (row N3): thin, exception-free shims over encoder.cu.
#include "api_internal.hpp"
#include "encoder.hpp"

using q3api::fail;
using q3api::guarded;

struct q3tts_encoder { q3::EncoderModel* m = nullptr; };

extern "C" {

int q3tts_encoder_load(const char* dir, const q3tts_options* opts, q3tts_encoder** out) {
  if (!dir || !out) return fail(Q3TTS_EINVAL, "q3tts_encoder_load: null argument");
  *out = nullptr;
  return guarded([&]() {
    q3tts_options o;
    q3tts_options_default(&o);
    if (opts) {
      if (opts->struct_size != sizeof(q3tts_options)) return fail(Q3TTS_EINVAL, "q3tts_options.struct_size mismatch");
      o = *opts;
    }
    std::unique_ptr<q3tts_encoder> h(new q3tts_encoder());
    h->m = q3::encoder_create(dir, o);
    *out = h.release();
    return (int)Q3TTS_OK;
  });
}

void q3tts_encoder_free(q3tts_encoder* e) {
  if (!e) return;
  q3::encoder_destroy(e->m);
  delete e;
}

int q3tts_encoder_info(const q3tts_encoder* e, int32_t* valid_quantizers, int32_t* codebook_size, int32_t* hop, int32_t* sampling_rate,
                       int64_t* num_parameters) {
  if (!e || !e->m) return fail(Q3TTS_EINVAL, "q3tts_encoder_info: null handle");
  const q3::EncoderConfig& c = e->m->cfg;
  if (valid_quantizers) *valid_quantizers = c.valid_quantizers;
  if (codebook_size) *codebook_size = c.codebook_size;
  if (hop) {
    int h = c.downsample_stride();
    for (int i = 0; i < c.n_ratios; ++i) h *= c.ratios[i];
    *hop = h;
  }
  if (sampling_rate) *sampling_rate = c.sampling_rate;
  if (num_parameters) *num_parameters = e->m->num_parameters;
  return Q3TTS_OK;
}

int64_t q3tts_encode_frames(const q3tts_encoder* e, int64_t samples) {
  if (!e || !e->m || samples < 0) return -1;
  return q3::encoder_frames(e->m->cfg, samples);
}

int q3tts_encode(q3tts_encoder* e, const float* audio, int32_t B, int64_t samples, int32_t* codes_out) {
  if (!e || !e->m) return fail(Q3TTS_EINVAL, "q3tts_encode: null handle");
  return guarded([&]() {
    q3::encoder_encode(*e->m, audio, B, samples, codes_out);
    return (int)Q3TTS_OK;
  });
}

int64_t q3tts_encoder_launch_count(const q3tts_encoder* e) { return (e && e->m) ? e->m->launches : -1; }

int q3tts_encoder_set_taps(q3tts_encoder* e, int32_t enable) {
  if (!e || !e->m) return fail(Q3TTS_EINVAL, "q3tts_encoder_set_taps: null handle");
  std::lock_guard<std::mutex> lock(e->m->mu);
  e->m->taps_enabled = enable != 0;
  return Q3TTS_OK;
}

int q3tts_encoder_tap(q3tts_encoder* e, const char* name, float* out, int64_t capacity, int64_t dims[3]) {
  if (!e || !e->m || !name || !dims) return fail(Q3TTS_EINVAL, "q3tts_encoder_tap: null argument");
  return guarded([&]() {
    q3::encoder_tap(*e->m, name, out, capacity, dims);
    return (int)Q3TTS_OK;
  });
}

}  // extern "C"
