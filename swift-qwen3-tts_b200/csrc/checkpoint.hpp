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

// Row N3 of SURVEY 8(f): the speech-tokenizer ENCODER (audio -> codes).  Hyper-parameters = Qwen3TTSTokenizerEncoderConfig
// (Config.swift:419-560, every key optional with the defaults below); tensors keyed by the Swift module path after the reference's
// remap (Qwen3.swift:1514-1527, 1544-1566, 1590-1679, 1726-1748), conv weights in MLX layout [Cout, K, Cin].
struct EncoderConfig {
  float frame_rate = 12.5f;
  int audio_channels = 1, codebook_dim = 256, codebook_size = 2048, compress = 2, dilation_growth_rate = 2, head_dim = 64;
  int hidden_size = 512, intermediate_size = 2048, kernel_size = 7, last_kernel_size = 3;
  int num_attention_heads = 8, num_filters = 64, num_hidden_layers = 8, num_key_value_heads = 8, num_quantizers = 32;
  int num_residual_layers = 1, residual_kernel_size = 3, sampling_rate = 24000;
  float rope_theta = 10000.0f;
  int n_ratios = 4, ratios[8] = {8, 6, 5, 4, 0, 0, 0, 0};     // upsampling_ratios; the Seanet walks them REVERSED (STE.swift:420)
  int use_causal_conv = 1, use_conv_shortcut = 0;
  int valid_quantizers = 16;                                   // encoder_valid_num_quantizers (Config.swift:586); STE.swift:957
  int downsample_stride() const {                              // STE.swift:1005-1006
    int s = 1;
    for (int i = 0; i < n_ratios; ++i) s *= ratios[i];
    return (int)(((float)sampling_rate / (float)s) / frame_rate);
  }
};
struct EncoderCheckpoint {
  EncoderConfig cfg;
  TensorMap tensors;
  int64_t num_parameters = 0;
};
// Throws q3::Error(Q3TTS_EFORMAT) when config.json has no encoder_config (the "lite" checkpoints) or a tensor is missing / misshapen.
void load_encoder_checkpoint(const std::string& dir, EncoderCheckpoint* out);

// The decoder's expected tensor inventory after sanitize: key -> MLX-layout shape.
std::map<std::string, std::vector<int64_t>> expected_decoder_tensors(const q3tts_config& cfg);

}  // namespace q3
