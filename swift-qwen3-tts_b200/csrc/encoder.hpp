// Speech-tokenizer ENCODER (audio -> codes), SURVEY 8(f) row N3.  Reference: SpeechTokenizerEncoder.swift:955-1056.
//
// Data layout (device, float32, channels-last): every level of the Seanet pyramid is [B, P, C] with P = frames of the NEXT level x
// that level's stride, i.e. the valid rows L rounded up to whole strides; rows L..P-1 stay zero, which is exactly the reference's
// right-hand "extra padding" (getExtraPaddingForConv1d, STE.swift:115-119), and lets a strided conv with k = 2 * stride run as a
// 2-tap GEMM over the same memory viewed as [B, P / stride, stride * C].  Codes leave as int32 [B, valid_quantizers, T].
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "checkpoint.hpp"
#include "kernels.cuh"

namespace q3 {

struct EncGemm {                     // [taps][N][Cin] fp32; tensor-core engine: the same weights as [taps][N][3 Cin] fp16 triples (kernels.cuh)
  float* w = nullptr; float* bias = nullptr; int taps = 1, Cin = 0, N = 0;
  __half* w3 = nullptr; float* bias_s = nullptr;   // bias x kSplitScale (zeros when there is none)
};
struct EncStage { EncGemm res3, res1, down; int dim = 0, ratio = 1; };                        // STE.swift:353-391
struct EncLayer { float *n1w, *n1b, *n2w, *n2b, *ls1, *ls2; EncGemm qkv, o, fc1, fc2; float *ls1_s = nullptr, *ls2_s = nullptr; };   // STE.swift:545-591; *_s = layer scale / kSplitScale
struct EncBook { int part = 0; EncGemm score; };                                               // score.w = E [K][D], score.bias = -|E|^2 / 2
struct EncTap { size_t offset; int64_t slot_rows, valid_rows; int C; };

struct EncoderModel {
  EncoderConfig cfg;
  int device = 0;
  int64_t num_parameters = 0;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  std::mutex mu;                       // calls on one handle are serialised
  float *init_w = nullptr, *init_b = nullptr, *inv_freq = nullptr;
  bool tc = false;                     // tensor-core engine (opts.precision == Q3TTS_PREC_FP16): GEMMs as three tcgen05 products of split fp16 operands
  float *zeros = nullptr, *inv_split = nullptr;   // [max N]: zero bias, 1 / kSplitScale
  std::vector<EncStage> stages;
  EncGemm final_conv, downsample, proj[2];
  std::vector<EncLayer> layers;
  std::vector<EncBook> books;
  void* arena = nullptr;               // grow-only workspace
  size_t arena_cap = 0;
  int zeroed_B = -1; int64_t zeroed_samples = -1; bool zeroed_taps = false;   // the shape the workspace was last cleared for
  bool taps_enabled = false;
  std::map<std::string, EncTap> tap_index;   // stage outputs of the last encode (inside the arena)
  int last_B = 0;
  int64_t launches = 0;                // kernels launched by the last encode
};

EncoderModel* encoder_create(const std::string& speech_tokenizer_dir, const q3tts_options& opts);
void encoder_destroy(EncoderModel* m);
int64_t encoder_frames(const EncoderConfig& c, int64_t samples);
// audio [B, samples] float32 (host) -> codes [B, valid_quantizers, frames] int32 (host).  Throws q3::Error.
void encoder_encode(EncoderModel& m, const float* audio, int B, int64_t samples, int32_t* codes_out);
// dims = {B, valid rows, C}; out == nullptr only queries the shape.  Layout [B, rows, C].
void encoder_tap(EncoderModel& m, const std::string& name, float* out, int64_t cap, int64_t dims[3]);

}  // namespace q3
