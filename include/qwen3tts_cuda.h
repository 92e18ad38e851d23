/*
 * libqwen3tts_cuda -- C ABI of the B200-native (sm_100a) 12 Hz speech-tokenizer DECODER
 * (codec token grids [B,16,T] int32 -> 24 kHz float PCM).
 *
 * Drop-in boundary for ONE path of AtomGradient/swift-qwen3-tts.  Every entry point names
 * the reference interface it replaces (paths relative to the reference repo;
 * ST.swift = Sources/Qwen3TTS/Models/SpeechTokenizer.swift,
 * Q3.swift = Sources/Qwen3TTS/Models/Qwen3.swift, Cfg.swift = Sources/Qwen3TTS/Models/Config.swift).
 *
 * Plain C: pointers and sizes only.  No PyTorch / MLX types, no CPU fallback: every compute
 * entry point fails with Q3TTS_ECUDA when no sm_100 device is usable.
 * Buffers are caller-owned; the library owns device weights and a grow-only workspace per model.
 * Calls on the same handle are serialised internally; use one handle per GPU.
 */
#ifndef QWEN3TTS_CUDA_H
#define QWEN3TTS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q3TTS_ABI_VERSION 1

/* ---- status codes (returned by every function that returns int) ------------------------
 * Plain integer macros, not enums: Swift imports `#define NAME <int literal>` as an Int32 constant that compares directly
 * with the `int` these functions return (an anonymous C enum would import as `Int`, a named one as a struct with rawValue). */
#define Q3TTS_OK 0
#define Q3TTS_EINVAL 1   /* bad argument / shape / code id >= codebook size (checked on device)  */
#define Q3TTS_EIO 2      /* file missing / unreadable                                            */
#define Q3TTS_EFORMAT 3  /* malformed config.json / safetensors, missing tensor, wrong shape     */
#define Q3TTS_ECUDA 4    /* CUDA error, or no sm_100 device                                      */
#define Q3TTS_ENOMEM 5   /* host or device allocation failed                                     */
#define Q3TTS_ESTATE 6   /* call not valid in this state (e.g. push on a closed stream)          */

/* ---- options ----------------------------------------------------------------------------- */
#define Q3TTS_PREC_FP32 0   /* CUDA-core fp32 everywhere: the parity anchor (PCM max-abs <= 1e-4) */
#define Q3TTS_PREC_FP16 1   /* tcgen05 fp16 operands, fp32 accumulate (SNR >= 40 dB)               */
#define Q3TTS_PREC_BF16 2   /* tcgen05 bf16 operands, fp32 accumulate                              */

#define Q3TTS_ATTN_REFERENCE 0 /* full, unmasked, no RoPE: what ST.swift:512-528,763 does          */
#define Q3TTS_ATTN_CAUSAL_SW 1 /* causal, window = sliding_window (Cfg.swift:401): streaming       */

#define Q3TTS_CODES_BQT 0   /* [B,16,T]  -- Qwen3TTSSpeechTokenizerDecoder.callAsFunction, ST.swift:754 */
#define Q3TTS_CODES_BTQ 1   /* [B,T,16]  -- Qwen3TTSSpeechTokenizer.decode, ST.swift:823               */

typedef struct q3tts_options {
  uint32_t struct_size;       /* = sizeof(q3tts_options)                                       */
  int32_t device;             /* CUDA device ordinal; -1 = current device                      */
  int32_t precision;          /* Q3TTS_PREC_*                                                  */
  int32_t attn_mode;          /* Q3TTS_ATTN_*                                                  */
  uint64_t workspace_bytes;   /* activation workspace cap; 0 = min(64 GiB, 45 % of free HBM)   */
  int32_t max_frames_per_launch; /* micro-batch cap in codec frames; 0 = derive from workspace */
  int32_t reserved;
} q3tts_options;

/* Decoder hyper-parameters as parsed from config.json (Cfg.swift:361-408 keys and defaults). */
typedef struct q3tts_config {
  int32_t latent_dim, codebook_dim, codebook_size, decoder_dim, hidden_size, intermediate_size;
  int32_t num_hidden_layers, num_attention_heads, num_key_value_heads, head_dim;
  int32_t sliding_window, num_quantizers, num_semantic_quantizers, semantic_codebook_size;
  int32_t num_upsample_rates, upsample_rates[8];
  int32_t num_upsampling_ratios, upsampling_ratios[8];
  int32_t total_upsample;     /* Cfg.swift:411-414 (1920)                                      */
  int32_t decode_upsample_rate; /* Cfg.swift:590 (1920): factor used for audioLengths         */
  int32_t output_sample_rate; /* Cfg.swift:589 (24000)                                         */
  int32_t has_encoder_config; /* ST.swift:816 hasEncoder (the encoder itself is out of scope)  */
  float rms_norm_eps, rope_theta, layer_scale_initial_scale;
  int64_t num_decoder_tensors;/* on-disk decoder.* tensors consumed (271 for the full model)   */
  int64_t num_parameters;     /* decoder parameters after codebook folding                     */
} q3tts_config;

typedef struct q3tts_model q3tts_model;
typedef struct q3tts_stream q3tts_stream;

/* ---- library ------------------------------------------------------------------------------ */
int q3tts_abi_version(void);
/* Thread-local message for the last non-OK status on this thread. */
const char* q3tts_last_error(void);
/* Number of usable sm_100 devices (0 when there is none: every compute call then fails).     */
int q3tts_device_count(void);
void q3tts_options_default(q3tts_options* opts);

/* ---- load / free ---------------------------------------------------------------------------
 * Replaces the speech-tokenizer half of Qwen3TTSModel.postLoadHook (Q3.swift:1461-1494):
 * reads <dir>/config.json (Qwen3TTSTokenizerConfig, Cfg.swift:565-595) and every
 * <dir>/ *.safetensors, applies sanitizeSpeechTokenizerWeights' decoder rules
 * (Q3.swift:1498-1512, 1530-1543, 1581-1588, 1687-1724 incl. the layout heuristic 1246-1260),
 * folds codebooks (embedding_sum / clip(cluster_usage, 1e-5)), packs and uploads the weights.
 * encoder.* tensors are ignored.  Unlike `update(verify: [])` (Q3.swift:1486) a missing or
 * mis-shaped decoder tensor is an error (Q3TTS_EFORMAT).  A config without decoder_config is
 * Q3TTS_EFORMAT (the reference calls fatalError, ST.swift:801-805).  F32 / F16 / BF16 files.   */
int q3tts_model_load(const char* speech_tokenizer_dir, const q3tts_options* opts, q3tts_model** out);
void q3tts_model_free(q3tts_model* m);
int q3tts_model_config(const q3tts_model* m, q3tts_config* out);
/* Host-only: parse + validate a checkpoint without touching CUDA (config, tensor inventory,
 * shapes, layout rules).  Fills *cfg when non-NULL.                                           */
int q3tts_checkpoint_inspect(const char* speech_tokenizer_dir, q3tts_config* cfg);

/* samples produced for T frames = T * total_upsample (ST.swift:752-753; Cfg.swift:411-414).   */
int64_t q3tts_output_samples(const q3tts_model* m, int64_t frames);

/* ---- decode: host buffers -------------------------------------------------------------------
 * Replaces Qwen3TTSSpeechTokenizer.decode (ST.swift:823-836) when layout = Q3TTS_CODES_BTQ and
 * Qwen3TTSSpeechTokenizerDecoder.callAsFunction (ST.swift:754-784) when Q3TTS_CODES_BQT.
 * codes: int32 [B,T,16] or [B,16,T].  pcm_out: float [B, T*total_upsample], clipped to [-1,1].
 * lengths_out (may be NULL): int32 [B] = count(code[b,t,0] > 0) * decode_upsample_rate
 * (ST.swift:831-833).  Synchronous: returns when PCM is in the caller's buffer.
 * A code id outside its codebook is Q3TTS_EINVAL (the reference leaves it unchecked).          */
int q3tts_decode(q3tts_model* m, const int32_t* codes, int32_t B, int32_t T, int32_t layout,
                 float* pcm_out, int32_t* lengths_out);

/* Mixed-length batch: utterance i has frame_offsets[i+1]-frame_offsets[i] frames.
 * codes_packed: int32 [sum_T, 16] (frame-major, = the [T,16] rows `generate` stacks, Q3.swift:736-741).
 * pcm_out: float [sum_T * total_upsample], utterance i at sample offset frame_offsets[i]*total_upsample.
 * Every utterance equals its own B=1 q3tts_decode (no cross-utterance leakage).                */
int q3tts_decode_varlen(q3tts_model* m, const int32_t* codes_packed, const int64_t* frame_offsets,
                        int32_t n_utterances, float* pcm_out, int32_t* lengths_out);

/* ---- decode straight to 16-bit PCM ------------------------------------------------------------
 * Same as q3tts_decode / q3tts_decode_varlen, but the last kernel of the chain writes
 * Int16(clamp(x,-1,1) * 32767) -- the conversion the reference's WAV writer applies to every
 * sample on the CPU (Sources/Qwen3TTSDemo/main.swift:158-160) -- so the device-to-host copy and the
 * PCM write traffic halve.  Bit-identical to q3tts_pcm_to_int16 of the float output.            */
int q3tts_decode_int16(q3tts_model* m, const int32_t* codes, int32_t B, int32_t T, int32_t layout,
                       int16_t* pcm_out, int32_t* lengths_out);
int q3tts_decode_varlen_int16(q3tts_model* m, const int32_t* codes_packed, const int64_t* frame_offsets,
                              int32_t n_utterances, int16_t* pcm_out, int32_t* lengths_out);

/* ---- CUDA graphs for small, launch-bound decodes (experimental, off by default) -----------------
 * A decode of a few hundred frames is ~95 kernel launches of a few microseconds each.  mode 1: a launch chain whose
 * shape and buffers have been seen before is captured once and replayed; -1: only chains of at most 2048 frames;
 * 0 (default): never -- on the B = 1, T = 125 decode the replay measured 2.83 ms against 2.28 ms for eager launches
 * (95 cluster-launch nodes with ~700 bytes of tensor-map parameters each).  Results are bit-identical either way.
 * Streams that cannot be captured (the legacy default stream) always run eagerly.                                */
int q3tts_set_graphs(q3tts_model* m, int32_t mode);

/* ---- codec-embedding sum for the Talker's next-step input -----------------
 * Replaces  codecEmbed = talker.getInputEmbeddings()(code0); for i in 1..<G { codecEmbed = codecEmbed +
 * codePredictor.codecEmbedding[i-1](code_i) }  (Qwen3.swift:720-728, 927-935, 1157-1162 per generated frame;
 * 485-491 for the whole reference-audio prefix of voice cloning), batched over frames.
 * q3tts_codec_embedder_load reads talker.model.codec_embedding.weight [3072,H] (Talker.swift:495, 510) and
 * talker.code_predictor.model.codec_embedding.{i}.weight [2048,H] (CodePredictor.swift:206, 217-219) from
 * the .safetensors files of <model_dir> and keeps them in their on-disk dtype.  codes: int32 [n_frames, G] frame-major
 * (the rows `generate` stacks, Q3.swift:736-741).  out: [n_frames, H] in the tables' dtype; every add is rounded
 * to that dtype, left to right, exactly as the reference's sequence of MLX adds.
 * A code id outside its table is Q3TTS_EINVAL (host variant; the reference leaves it unchecked).  */
typedef struct q3tts_codec_embedder q3tts_codec_embedder;
int q3tts_codec_embedder_load(const char* model_dir, int32_t device, q3tts_codec_embedder** out);
void q3tts_codec_embedder_free(q3tts_codec_embedder* e);
/* precision: Q3TTS_PREC_* of the tables; vocab (may be NULL): int32 [groups] table sizes */
int q3tts_codec_embedder_info(const q3tts_codec_embedder* e, int32_t* hidden, int32_t* groups, int32_t* precision,
                              int32_t* vocab);
int q3tts_codec_embed_sum(q3tts_codec_embedder* e, const int32_t* codes, int64_t n_frames, void* out);
int q3tts_codec_embed_sum_device(q3tts_codec_embedder* e, const int32_t* d_codes, int64_t n_frames, void* d_out,
                                 void* stream);

/* ---- speech-tokenizer ENCODER: audio -> codes ----------------------------
 * Replaces Qwen3TTSSpeechTokenizer.encode / Qwen3TTSSpeechTokenizerEncoder.encode
 * (SpeechTokenizer.swift:838-852, SpeechTokenizerEncoder.swift:1031-1056; callers: voice cloning, Qwen3.swift:430-440):
 * Seanet encoder (causal convs, strides 4,5,6,8) -> 8-layer causal transformer with RoPE -> stride-2 conv -> split residual
 * vector quantizer (nearest codebook entry per layer, float32), first `valid_quantizers` (16) codebooks returned.
 * q3tts_encoder_load reads <dir>/config.json (`encoder_config`, Config.swift:419-560) and the `encoder.*` tensors of the
 * directory's .safetensors files with the reference's key remap (Qwen3.swift:1514-1748); a checkpoint without an
 * encoder (the "lite" variants) is Q3TTS_EFORMAT.  opts->device selects the GPU; opts->precision selects the engine, both of
 * float32 accuracy (a code is an argmin: lower precision would be a different tokenizer): Q3TTS_PREC_FP16 (the options' default)
 * = tensor cores, every float32 operand carried as a pair of fp16 numbers and every GEMM as one tcgen05 product of three times
 * the depth; Q3TTS_PREC_FP32 = CUDA cores.  Q3TTS_PREC_BF16 is rejected (two bf16 halves hold 16 mantissa bits).
 * audio: float32 [B, samples] (the reference's [B, 1, samples]); codes_out: int32 [B, valid_quantizers, T] with
 * T = q3tts_encode_frames(samples) = the ceil-division chain of the strides (12.5 frames per second).
 * Every utterance of a call has the same length (the reference API); encode different lengths in separate calls.     */
typedef struct q3tts_encoder q3tts_encoder;
int q3tts_encoder_load(const char* speech_tokenizer_dir, const q3tts_options* opts, q3tts_encoder** out);
void q3tts_encoder_free(q3tts_encoder* e);
/* valid_quantizers, codebook_size, hop (input samples per code frame when the length divides evenly), sampling_rate; any may be NULL */
int q3tts_encoder_info(const q3tts_encoder* e, int32_t* valid_quantizers, int32_t* codebook_size, int32_t* hop,
                       int32_t* sampling_rate, int64_t* num_parameters);
int64_t q3tts_encode_frames(const q3tts_encoder* e, int64_t samples);
int q3tts_encode(q3tts_encoder* e, const float* audio, int32_t B, int64_t samples, int32_t* codes_out);
/* kernels launched by the last q3tts_encode on this handle (benchmark bookkeeping) */
int64_t q3tts_encoder_launch_count(const q3tts_encoder* e);
/* debug / parity: keep stage outputs of the next encodes; q3tts_encoder_tap copies stage `name` as float32 [B, rows, C]
 * (channels last) and writes {B, rows, C} to dims; out == NULL only queries dims.  Names: "hid0".."hid3" / "res0".."res3" (the Seanet stages' hidden
 * activation and stream after the residual block), "layer3" (last strided conv), "seanet", "transformer", "downsample".                                                                   */
int q3tts_encoder_set_taps(q3tts_encoder* e, int32_t enable);
int q3tts_encoder_tap(q3tts_encoder* e, const char* name, float* out, int64_t capacity, int64_t dims[3]);

/* ---- decode: device buffers (codes and PCM already in HBM), asynchronous on `stream` ---------
 * Same semantics as q3tts_decode; `stream` is a cudaStream_t (NULL = legacy default stream).
 * Returns after enqueueing; errors detected on device (bad code ids) surface at
 * q3tts_sync().                                                                                */
int q3tts_decode_device(q3tts_model* m, const int32_t* d_codes, int32_t B, int32_t T, int32_t layout,
                        float* d_pcm_out, int32_t* d_lengths_out, void* stream);
/* q3tts_decode_varlen with the packed codes / PCM / lengths in HBM; frame_offsets is a HOST array [n+1].           */
int q3tts_decode_varlen_device(q3tts_model* m, const int32_t* d_codes_packed, const int64_t* frame_offsets,
                               int32_t n_utterances, float* d_pcm_out, int32_t* d_lengths_out, void* stream);
int q3tts_sync(q3tts_model* m, void* stream);

/* ---- stage taps (debug / parity) --------------------------------------------------------------
 * The reference's test walks the decoder stage by stage (Tests.swift:57-257).  After a decode
 * with taps enabled, q3tts_stage_tap copies stage `name` as float32 NCT [B, C, L] (the
 * reference's inter-module layout) into `out`.  Names: "rvq_sum_first", "rvq_sum_rest",
 * "quantized", "pre_conv", "pre_transformer", "upsample0", "upsample1", "init_conv",
 * "block0".."block3", "out_conv".  Taps force single-launch decode (B*T bounded by workspace). */
int q3tts_set_taps(q3tts_model* m, int32_t enable);
int q3tts_stage_tap_shape(q3tts_model* m, const char* name, int32_t* B, int32_t* C, int64_t* L);
int q3tts_stage_tap(q3tts_model* m, const char* name, float* out, int64_t out_elems);
/* Packed weight probe, MLX layout of the reference module tree, e.g.
 * "decoder.decoder.initConv.conv.weight" -> [1536,7,1024] (Tests.swift:131-132).               */
int q3tts_weight_shape(const q3tts_model* m, const char* swift_key, int32_t* ndim, int64_t dims[4]);

/* ---- streaming (chunked decode with causal state carry; Q3TTS_ATTN_CAUSAL_SW only) ------------
 * The reference has no chunked PCM streaming (Qwen3+Streaming.swift:19-120 emits one final
 * .audio); correctness here is chunk-invariance: concatenated chunk PCM == one-shot decode in
 * the same mode.  A stream is pinned to the model's GPU for life.  State per stream (device): the last 2
 * pre_conv inputs, K/V of the last sliding_window-1 frames per transformer layer, and, for each of the 20 consumers
 * of the conv stack that read rows in front of their tile (ConvNeXt depthwise convs, initConv, the transposed convs, every
 * dilated conv7, outConv), the last 1-3 frames of ITS input (about 4 MB per stream in fp16).  A push computes the conv stack
 * over [3 context frames | new frames] and restores each consumer's context rows from that state.
 * Code ids are validated on the host before anything is enqueued: a rejected push (Q3TTS_EINVAL) advances no stream.
 * q3tts_model_free on a model with open streams only marks it; the last q3tts_stream_close deletes it.              */
int q3tts_stream_open(q3tts_model* m, q3tts_stream** out);
/* codes: int32 [n_frames,16]; pcm_out: float [n_frames*total_upsample].                        */
int q3tts_stream_push(q3tts_stream* s, const int32_t* codes, int32_t n_frames, float* pcm_out);
/* Push one chunk for each of n streams in ONE batched launch chain (config 5).                 */
int q3tts_stream_push_batch(q3tts_stream* const* streams, int32_t n_streams,
                            const int32_t* const* codes, const int32_t* n_frames, float* const* pcm_out);
/* Frames pushed so far (-1 for a NULL / closed stream). */
int64_t q3tts_stream_frames(const q3tts_stream* s);
void q3tts_stream_close(q3tts_stream* s);

/* ---- batch scheduler (host-only, no CUDA) ------------------------------------------------------
 * Longest-processing-time-first partition of utterances over `n_parts` GPUs by frame count;
 * part_out[i] in [0,n_parts).  Deterministic (ties by index) so every rank computes the same map. */
int q3tts_partition_lpt(const int64_t* frames, int32_t n_utterances, int32_t n_parts, int32_t* part_out);

/* ---- multi-GPU pool: one worker thread + CUDA context + model replica per GPU of ONE box (north_star (d)) --------
 * The reference decodes one utterance at a time on one device (Q3.swift:744, 951, 1186); a batch of independent utterances
 * shards by utterance with no collective.  q3tts_pool_open loads the checkpoint once per device
 * (`devices` = NULL: every sm_100 device, else n_devices ordinals; opts->device is ignored).  A decode call partitions the
 * utterances with q3tts_partition_lpt, hands each worker its share, and returns when every worker has written its PCM into
 * the caller's buffers (pageable or pinned; each worker pipelines its own copies).  Results are bit-identical to
 * q3tts_decode_varlen on one GPU: an utterance's PCM does not depend on which utterances share its launch chain.
 * Calls on one pool are serialised.  Errors: the first failing worker's status and message.                       */
typedef struct q3tts_pool q3tts_pool;
int q3tts_pool_open(const char* speech_tokenizer_dir, const q3tts_options* opts, const int32_t* devices,
                    int32_t n_devices, q3tts_pool** out);
void q3tts_pool_close(q3tts_pool* p);
int32_t q3tts_pool_size(const q3tts_pool* p);
int q3tts_pool_decode_varlen(q3tts_pool* p, const int32_t* codes_packed, const int64_t* frame_offsets,
                             int32_t n_utterances, float* pcm_out, int32_t* lengths_out);
int q3tts_pool_decode_varlen_int16(q3tts_pool* p, const int32_t* codes_packed, const int64_t* frame_offsets,
                                   int32_t n_utterances, int16_t* pcm_out, int32_t* lengths_out);
/* Uniform batch [B,T,16] / [B,16,T] split by utterance over the pool (q3tts_decode semantics). */
int q3tts_pool_decode(q3tts_pool* p, const int32_t* codes, int32_t B, int32_t T, int32_t layout, float* pcm_out,
                      int32_t* lengths_out);
/* Per worker, for the most recent pool call: wall milliseconds spent in its decode and frames it was given. */
int q3tts_pool_last_stats(const q3tts_pool* p, float* ms_out, int64_t* frames_out, int32_t cap);

/* ---- PCM post-processing (caller-side semantics of the reference) ------------------------------
 * q3tts_trim_length: Q3.swift:746-752 (valid_len in (0,n) trims, else keeps n).
 * q3tts_voice_clone_cut: Q3.swift:1196-1199, Float arithmetic: first sample to keep.
 * q3tts_pcm_to_int16: Sources/Qwen3TTSDemo/main.swift:134-165 (Int16(clamp(x,-1,1) * 32767)).   */
int64_t q3tts_trim_length(int64_t n_samples, int64_t valid_len);
int64_t q3tts_voice_clone_cut(int64_t ref_frames, int64_t total_frames, int64_t n_samples);
int q3tts_pcm_to_int16(const float* pcm, int64_t n, int16_t* out);
int q3tts_write_wav(const char* path, const float* pcm, int64_t n, int32_t sample_rate);

/* ---- measurement -------------------------------------------------------------------------------
 * Per-stage CUDA-event timing of the most recent decode (enable before decoding).
 * q3tts_profile_get: fills up to `cap` entries; returns the number of stages recorded.          */
typedef struct q3tts_stage_time {
  char name[32];
  float ms;            /* device time of the stage (sum over its kernels, events on the launch stream) */
  int32_t launches;    /* kernels launched by this stage                                               */
  double flops;        /* algorithmic FLOPs of the stage for this decode                               */
  double bytes;        /* algorithmic HBM bytes (one read of inputs + one write of outputs + weights)   */
} q3tts_stage_time;
int q3tts_profile_enable(q3tts_model* m, int32_t enable);
int q3tts_profile_get(q3tts_model* m, q3tts_stage_time* out, int32_t cap);
/* Per-kernel CUDA-event timing of the most recent profiled decode, aggregated by "stage.op" label
 * (events bracket each launch on the launch stream); flops / bytes are the algorithmic work.     */
typedef struct q3tts_kernel_time {
  char name[48];
  float ms;            /* sum over the launches of this label                                    */
  int32_t launches;
  double flops, bytes; /* summed algorithmic FLOPs / HBM bytes of those launches                 */
} q3tts_kernel_time;
int q3tts_profile_kernels(q3tts_model* m, q3tts_kernel_time* out, int32_t cap);
/* Kernels launched by this model since load (all decodes).                                      */
int64_t q3tts_launch_count(const q3tts_model* m);

/* ---- kernel-level test hook (no reference counterpart) ------------------------------------------
 * One multi-tap GEMM (the op behind every conv / linear / transposed conv, ST.swift:293-305, 339-353)
 * on seeded random data: the tcgen05 kernel against the CUDA-core kernel of the same op, then
 * `iters` timed launches.  mode 0 conv7-like, 1 conv1-like (in-place residual), 2 transposed-conv-like,
 * 3 plain.  Outputs: mean ms per launch, max |tc - simt| of the stream and operand outputs.       */
int q3tts_debug_conv_gemm(int32_t B, int32_t rows, int32_t Cin, int32_t N, int32_t taps, int32_t dil,
                          int32_t mode, int32_t precision, int32_t iters, float* ms_out,
                          float* max_diff_y, float* max_diff_a);
/* The fused residual unit of the 96-channel block (ST.swift:430-437: x + conv1(snake(conv7(snake(x))))) against
 * the same unit composed from three CUDA-core GEMM launches; out_snake = apply the consumer's SnakeBeta. */
int q3tts_debug_resunit(int32_t B, int32_t rows, int32_t dil, int32_t out_snake, int32_t precision,
                        int32_t iters, float* ms_out, float* max_diff);
/* The residual unit fused into the tcgen05 GEMM (C = 128 or 192: conv7 + snake -> smem operand tile -> conv1 +
 * residual [+ snake]) against two CUDA-core GEMM launches.                                                          */
int q3tts_debug_fused_unit(int32_t B, int32_t rows, int32_t C, int32_t dil, int32_t with_operand, int32_t precision,
                           int32_t iters, float* ms_out, float* max_diff_y, float* max_diff_a);

/* The attention op on its own (ST.swift:519-525: MLXFast.scaledDotProductAttention(q, k, v, scale: head_dim^-0.5, mask)) through the
 * production dispatch, on caller-provided rows.  qkv: host float [B, T, (nh + 2 nkv) * hd] (q | k | v per row, rounded to the
 * precision's operand type on the device); out: host float [B, T, nh * hd].  len / row_begin (may be NULL): int32 [B], valid rows
 * [row_begin[b], len[b]) of each slot; window 0 = full (reference) attention, > 0 = causal sliding window.                     */
int q3tts_debug_attention(const float* qkv, int32_t B, int32_t T, int32_t nh, int32_t nkv, int32_t hd, const int32_t* len,
                          const int32_t* row_begin, int32_t window, int32_t precision, float* out);

#ifdef __cplusplus
}
#endif
#endif /* QWEN3TTS_CUDA_H */
