// Decoder engine: packed device weights + workspace + the launch chain for one micro-batch.
// Stage order follows Qwen3TTSSpeechTokenizerDecoder.callAsFunction (ST.swift:754-784).
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "checkpoint.hpp"
#include "kernels.cuh"

namespace q3 {

struct GemmW {            // one packed multi-tap GEMM weight (see ConvGemmParams)
  float* w32 = nullptr;   // [taps][N][Cin] fp32 (always kept: parity mode + source for 16-bit copies)
  void* w16 = nullptr;    // same, operand dtype (fast mode only)
  float* bias = nullptr;  // [N] or null
  int taps = 1, N = 0, Cin = 0, dil = 1;
};

struct SnakeW { float* ea = nullptr; float* ib = nullptr; int n = 0; };

struct LayerW {
  float *ln1 = nullptr, *ln2 = nullptr, *ls_attn = nullptr, *ls_mlp = nullptr;
  GemmW qkv, o, gate_up, down;
};

struct UpsampleW {
  GemmW tconv, pw1, pw2;
  float *dw_w = nullptr, *dw_b = nullptr, *ln_w = nullptr, *ln_b = nullptr, *gamma = nullptr;
  int ratio = 2;
};

struct BlockW {
  GemmW tconv;            // taps=2, N = r*cout
  SnakeW act_in_next[3];  // snake applied to the block's running output before res1/res2/res3's conv7
  GemmW conv7[3], conv1[3];
  SnakeW act2[3];
  int rate = 1, cin = 0, cout = 0;
};

struct StageProfile {
  std::string name;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool used = false;
  int launches = 0;
  double flops = 0, bytes = 0;
  float ms_accum = 0;
};

// One launch of the most recent profiled decode (profile mode only): label = stage.op, events on the launch stream.
struct LaunchProfile { std::string label; cudaEvent_t ev0 = nullptr, ev1 = nullptr; double flops = 0, bytes = 0; };

struct TapBuf { float* d = nullptr; int B = 0, C = 0; int64_t L = 0; size_t cap = 0; };

struct Model {
  q3tts_config cfg{};
  q3tts_options opts{};
  int device = 0;
  int op_dtype = DT_F32;      // operand dtype (GEMM inputs)
  int st_dtype = DT_F32;      // stream dtype inside the decoder blocks
  cudaStream_t stream = nullptr;
  std::mutex mu;
  // Every launch chain shares the workspace arena, d_meta and d_err.  The mutex serialises host-side enqueueing only, so chains on
  // DIFFERENT streams (q3tts_decode_device takes the caller's) are ordered on the device through this event: recorded at the end of
  // every chain, waited for by the next chain when its stream differs.
  cudaEvent_t chain_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_chain = false;

  // weights
  std::vector<void*> allocs;              // every device allocation made for weights
  std::vector<float*> codebooks;          // [num_q] device fp32 tables
  const float** d_tables = nullptr;       // device array of table pointers
  int32_t* d_table_sizes = nullptr;
  GemmW rvq_proj, pre_conv, in_proj, out_proj, init_conv;
  float* final_norm = nullptr;
  std::vector<LayerW> layers;
  std::vector<UpsampleW> ups;
  SnakeW block_in_snake[4];               // blockN.snake (applied by the producer of the block's input)
  BlockW blocks[4];
  SnakeW out_snake;
  float* tail_w = nullptr;                // [7][C]
  float tail_bias = 0.f;
  void* tail_w16 = nullptr;               // [16][C] hi / lo tap tile in the operand type (launch_tail_tile)
  bool fused_tail = false;                // block 3's last residual unit multiplies its output with tail_w16 and stores 16 partial products per row
  std::map<std::string, std::vector<int64_t>> weight_shapes;  // Swift key -> MLX shape

  // workspace
  char* arena = nullptr;
  size_t arena_cap = 0;
  uint64_t default_workspace = 24ull << 30;   // activation budget when q3tts_options.workspace_bytes == 0 (set at load)
  int32_t* d_codes = nullptr; size_t d_codes_cap = 0;
  // CUDA graphs for launch-bound chains (small micro-batches: ~95 launches of a few microseconds each): a micro-batch whose
  // key (shape, every pointer the launches bake in, output format) has been seen before is captured once and replayed.
  struct GraphEntry { std::vector<long long> key; cudaGraphExec_t exec = nullptr; long long launches = 0; int seen = 0; };
  std::vector<GraphEntry> graphs;
  int graph_mode = 0;            // 0: off (default: measured SLOWER than eager launches), 1: always, -1: chains of <= graph_max_frames frames
  long long graph_max_frames = 2048;
  char* stream_hook_h = nullptr; char* stream_hook_d = nullptr; size_t stream_hook_cap = 0;   // streaming: per-consumer copy lists
  float* d_pcm = nullptr;     size_t d_pcm_cap = 0;
  // host-buffer decodes: copy-out stream + pinned staging (codes in; PCM out, double-buffered per micro-batch; lengths)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t mb_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
  char* h_codes = nullptr;    size_t h_codes_cap = 0;
  char* h_pcm[2] = {nullptr, nullptr}; size_t h_pcm_cap[2] = {0, 0};
  char* h_len = nullptr;      size_t h_len_cap = 0;
  bool pcm_i16 = false;       // this call's tail writes int16 PCM into the (float-sized) output buffer
  int32_t* d_lengths = nullptr; size_t d_lengths_cap = 0;
  char* d_meta = nullptr;     size_t d_meta_cap = 0;     // len_frames / code_base / pcm_base per micro-batch
  char* h_meta = nullptr;     size_t h_meta_cap = 0;     // pinned staging
  std::vector<char> meta_key;                            // last uploaded metadata (skip re-upload when equal)
  int* d_err = nullptr;
  int* h_err = nullptr;                                  // pinned

  // streaming workspace (grow-only) + pinned staging for its metadata
  char* stream_ws = nullptr;  size_t stream_ws_cap = 0;
  char* stream_meta_h = nullptr; size_t stream_meta_cap = 0;

  // taps / profiling
  bool taps_enabled = false;
  std::map<std::string, TapBuf> taps;
  bool profile_enabled = false;
  std::vector<StageProfile> prof;
  std::vector<LaunchProfile> launch_prof;   // per-launch records of the current profiled decode
  std::vector<cudaEvent_t> event_pool;      // recycled events
  std::map<std::string, std::array<double, 4>> kernel_totals;   // label -> {ms, launches, flops, bytes}
  int64_t launches = 0;

  ~Model();
};

// Chunked streaming (Q3TTS_ATTN_CAUSAL_SW only).  Per stream, on the device, right-aligned histories:
//   q_hist  [2][codebook_dim]            last inputs of pre_conv (k = 3)
//   kv      [layers][W-1][qkv width]     K and V of the last W-1 frames of every transformer layer (W = sliding_window)
//   conv   per haloed consumer of the conv stack (ConvNeXt depthwise convs, initConv, the blocks' transposed convs,
//          every residual unit's dilated conv7, outConv): the last ceil(halo / rate) frames of ITS INPUT.  A push lays
//          the slots out as [Hs context frames | new frames] (Hs = the largest of those frame counts, 3 for the 12 Hz
//          decoder), computes every stage over the whole slot and, right before each haloed consumer, overwrites the
//          context rows it reads with the saved state (and saves the new tail).  Outputs computed for context frames
//          are garbage by construction and are never read: exact, at (Hs + n) / n times the conv work of a chunk of n
//          frames (the receptive-field overlap-save this replaces cost (10 + n) / n).
struct StreamState {
  int64_t frames_done = 0;
  void* q_hist = nullptr;
  void* kv = nullptr;
  void* conv = nullptr;          // one block, stages at the offsets of conv_state_layout()
};
struct HaloStage { int frames, rate, C, es; size_t off, bytes; };   // state of one haloed consumer: frames x rate rows x C x es bytes
std::vector<HaloStage> conv_state_layout(const Model& m);           // in the order run_back reaches the consumers
int stream_context_frames(const Model& m);             // Hs
void stream_state_alloc(Model& m, StreamState& st);
void stream_state_free(Model& m, StreamState& st);
// One chunk for each of S streams in one launch chain.  d_codes: packed [sum n, Q] frame-major; n_frames: host [S];
// d_pcm_out: packed [sum n * total_upsample].  Advances frames_done.
// run_microbatch, or the replay of its captured graph (see Model::graphs).
void run_microbatch_graphed(Model& m, const int32_t* d_codes, const int64_t* d_code_base, int64_t sq, int64_t st,
                            const int* d_len, const int64_t* d_pcm_base, float* d_pcm, int B, int Tmax, int64_t valid_frames,
                            cudaStream_t s);
void run_stream_batch(Model& m, StreamState* const* streams, int S, const int32_t* d_codes, const int* n_frames,
                      float* d_pcm_out, cudaStream_t s);

struct MicroBatch { int first = 0, B = 0, Tmax = 0; };   // utterances [first, first+B) of the sorted order

// Order a new launch chain on `s` behind the previous chain of this model (no-op when it ran on the same stream), and mark its end.
void chain_begin(Model& m, cudaStream_t s);
void chain_end(Model& m, cudaStream_t s);

// Build a model from a parsed checkpoint (uploads weights).  Throws q3::Error.
Model* model_create(const Checkpoint& ck, const q3tts_options& opts);

// Bytes of arena needed for a micro-batch of B utterances x Tmax frames.
size_t plan_bytes(const Model& m, int B, int Tmax);

// Enqueue the whole launch chain for one micro-batch on `s`.  len/code_base/pcm_base are device arrays.
void run_microbatch(Model& m, const int32_t* d_codes, const int64_t* d_code_base, int64_t sq, int64_t st,
                    const int* d_len, const int64_t* d_pcm_base, float* d_pcm, int B, int Tmax, int64_t valid_frames,
                    cudaStream_t s);

}  // namespace q3
