// Host-side checkpoint reader for the speech-tokenizer DECODER: config.json + *.safetensors ->
// float32 tensors keyed by the reference's Swift module path, in MLX layout.
// Replaces the decoder half of Q3.swift:1461-1494 (postLoadHook) and Q3.swift:1498-1750
// (sanitizeSpeechTokenizerWeights).  No CUDA in this file.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qwen3tts_cuda.h"

namespace q3 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

using TensorMap = std::map<std::string, HostTensor>;

struct Checkpoint {
  q3tts_config cfg{};
  TensorMap tensors;       // sanitized, MLX layout, Swift keys ("decoder.decoder.initConv.conv.weight" ...)
};

// Qwen3.swift:1246-1260 -- "already MLX layout?" for a 3-D conv weight.
bool is_mlx_conv_layout(const std::vector<int64_t>& shape);

// Parse <dir>/config.json into cfg (Cfg.swift:385-409 defaults when a key is absent).
void parse_tokenizer_config(const std::string& dir, q3tts_config* cfg);

// Load + sanitize + validate.  Throws q3::Error(Q3TTS_EIO / Q3TTS_EFORMAT).
void load_checkpoint(const std::string& dir, Checkpoint* out);

// Row N2 of SURVEY 8(f): the 1 + (num_code_groups - 1) codec-embedding tables of the MAIN checkpoint
// (<model_dir>/*.safetensors): talker.model.codec_embedding.weight [3072, H] (Talker.swift:495, 510) and
// talker.code_predictor.model.codec_embedding.{i}.weight [2048, H] (CodePredictor.swift:206, 217-219).
struct CodecEmbeddingTables {
  int64_t hidden = 0;
  std::string dtype;                 // on-disk dtype name ("BF16", "F16", "F32"): the sums are rounded to it after every add
  std::vector<HostTensor> tables;    // [0] = talker table, [1..] = code-predictor tables, values exact in float
};
void load_codec_embeddings(const std::string& model_dir, CodecEmbeddingTables* out);

// The decoder's expected tensor inventory after sanitize: key -> MLX-layout shape.
std::map<std::string, std::vector<int64_t>> expected_decoder_tensors(const q3tts_config& cfg);

}  // namespace q3
