// AddressSanitizer / UBSan harness for the host side of the library (SURVEY section 5 "race detection / sanitizers"):
// the checkpoint reader (config.json + safetensors parser + sanitize rules) and the host-only C ABI, with no CUDA in the link.
//   sanitize_harness <dir> [<dir> ...]   -- every directory is parsed as a speech_tokenizer checkpoint (valid or deliberately
//   corrupt: both must end in a status code, never in a sanitizer report), then the host-only entry points run on edge cases.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/qwen3tts_cuda.h"
#include "checkpoint.hpp"

int main(int argc, char** argv) {
  int ok = 0, rejected = 0;
  for (int i = 1; i < argc; ++i) {
    q3tts_config cfg{};
    const int st = q3tts_checkpoint_inspect(argv[i], &cfg);
    if (st == Q3TTS_OK) {
      ++ok;
      std::printf("%s: OK, %lld tensors, %lld parameters\n", argv[i], (long long)cfg.num_decoder_tensors, (long long)cfg.num_parameters);
    } else {
      ++rejected;
      std::printf("%s: status %d (%s)\n", argv[i], st, q3tts_last_error());
    }
    try {   // the codec-embedding reader walks the same safetensors parser with other keys
      q3::CodecEmbeddingTables t;
      q3::load_codec_embeddings(argv[i], &t);
    } catch (const q3::Error&) {
    }
    try {   // ... and so does the encoder reader (row N3): key remap, forced transposes, strict shape table
      q3::EncoderCheckpoint e;
      q3::load_encoder_checkpoint(argv[i], &e);
      std::printf("%s: encoder OK, %lld parameters\n", argv[i], (long long)e.num_parameters);
    } catch (const q3::Error& e) {
      std::printf("%s: encoder: %s\n", argv[i], e.what());
    }
  }
  // host-only entry points on edge cases
  {
    std::vector<int64_t> frames = {750, 25, 0, 300, 300, 1, 749, 26, -5};
    std::vector<int32_t> part(frames.size(), -1);
    if (q3tts_partition_lpt(frames.data(), (int32_t)frames.size(), 4, part.data()) != Q3TTS_OK) return 2;
    for (int32_t p : part) if (p < 0 || p >= 4) return 2;
    if (q3tts_partition_lpt(nullptr, 0, 3, nullptr) != Q3TTS_OK) return 2;
    if (q3tts_partition_lpt(frames.data(), 3, 0, part.data()) != Q3TTS_EINVAL) return 2;
  }
  {
    const float pcm[6] = {-2.0f, -1.0f, -0.5f, 0.0f, 0.99999f, 3.0f};
    int16_t out[6];
    if (q3tts_pcm_to_int16(pcm, 6, out) != Q3TTS_OK || out[0] != -32767 || out[5] != 32767 || out[3] != 0) return 3;
    if (q3tts_pcm_to_int16(nullptr, 0, nullptr) != Q3TTS_OK) return 3;
    if (q3tts_trim_length(100, 0) != 100 || q3tts_trim_length(100, 40) != 40 || q3tts_trim_length(100, 100) != 100) return 3;
    if (q3tts_voice_clone_cut(3, 10, 1000) != 300 || q3tts_voice_clone_cut(0, 0, 10) != 0) return 3;
    const char* tmp = std::getenv("Q3TTS_HARNESS_WAV");
    if (tmp && q3tts_write_wav(tmp, pcm, 6, 24000) != Q3TTS_OK) return 3;
    if (q3tts_write_wav("/nonexistent-dir/x.wav", pcm, 6, 24000) != Q3TTS_EIO) return 3;
  }
  std::printf("harness: %d parsed, %d rejected\n", ok, rejected);
  return 0;
}
