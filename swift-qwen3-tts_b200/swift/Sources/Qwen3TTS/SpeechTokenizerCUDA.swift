//
//  SpeechTokenizerCUDA.swift
//
//  Drop-in replacement for the decode side of Sources/Qwen3TTS/Models/SpeechTokenizer.swift
//  (Qwen3TTSSpeechTokenizer / Qwen3TTSSpeechTokenizerDecoder, lines 696-852 of the reference)
//  whose MLXArray ops are replaced by calls into libqwen3tts_cuda through the CQwen3TTSCUDA
//  system-library module.  Same type names, method names, argument labels and shapes; MLXArray is
//  replaced by the minimal host tensors below because MLX is not a dependency of this path.
//
//  NOTE: authored against the C header, NOT compiled in the build image (no Swift toolchain there);
//  the same C ABI is exercised by the C++/ctypes tests.  See INTEGRATION.md.
//

import Foundation
import CQwen3TTSCUDA

/// Minimal row-major host tensors (what the decode path needs from MLXArray).
public struct Int32Tensor {
    public var shape: [Int]
    public var data: [Int32]
    public init(shape: [Int], data: [Int32]) {
        precondition(shape.reduce(1, *) == data.count, "shape does not match element count")
        self.shape = shape
        self.data = data
    }
}

public struct FloatTensor {
    public var shape: [Int]
    public var data: [Float]
    public subscript(row: Int) -> ArraySlice<Float> {          // audio[0]  (Qwen3.swift:748)
        let n = shape.dropFirst().reduce(1, *)
        return data[(row * n)..<((row + 1) * n)]
    }
}

/// Mirrors `AudioGenerationError.audioDecodingFailed(String)` (Core/GenerationTypes.swift:67).
public enum Qwen3TTSCUDAError: Error, LocalizedError {
    case audioDecodingFailed(String)
    public var errorDescription: String? {
        if case .audioDecodingFailed(let m) = self { return "Audio decoding failed: \(m)" }
        return nil
    }
}

// The header's status / option constants are integer-literal macros (`#define Q3TTS_OK 0`), which Swift imports as Int32
// constants: they compare directly with the `Int32` every entry point returns.
@inline(__always)
private func check(_ status: Int32) throws {
    if status != Q3TTS_OK {
        throw Qwen3TTSCUDAError.audioDecodingFailed(String(cString: q3tts_last_error()))
    }
}

public enum Qwen3TTSPrecision: Int32 { case fp32 = 0, fp16 = 1, bf16 = 2 }

/// `.reference`: full, unmasked attention without RoPE -- what SpeechTokenizer.swift:512-528, 763 computes.
/// `.causalSlidingWindow`: causal, `sliding_window` keys (Config.swift:401) -- required by `Qwen3TTSDecodeStream`.
public enum Qwen3TTSAttentionMode: Int32 { case reference = 0, causalSlidingWindow = 1 }

/// Main decoder: codes -> audio waveform (SpeechTokenizer.swift:696-785).
public final class Qwen3TTSSpeechTokenizerDecoder {
    let handle: OpaquePointer
    let totalUpsample: Int
    let numQuantizers: Int

    init(handle: OpaquePointer, config: q3tts_config) {
        self.handle = handle
        self.totalUpsample = Int(config.total_upsample)
        self.numQuantizers = Int(config.num_quantizers)
    }

    /// Decode codes to audio.
    /// - Parameter codes: [batch, num_quantizers, time]
    /// - Returns: [batch, 1, samples], clipped to [-1, 1]      (SpeechTokenizer.swift:754-784)
    public func callAsFunction(_ codes: Int32Tensor) throws -> FloatTensor {
        precondition(codes.shape.count == 3 && codes.shape[1] == numQuantizers)
        let (b, t) = (codes.shape[0], codes.shape[2])
        var pcm = [Float](repeating: 0, count: b * t * totalUpsample)
        try codes.data.withUnsafeBufferPointer { c in
            try pcm.withUnsafeMutableBufferPointer { p in
                try check(q3tts_decode(handle, c.baseAddress, Int32(b), Int32(t), Q3TTS_CODES_BQT,
                                       p.baseAddress, nil))
            }
        }
        return FloatTensor(shape: [b, 1, t * totalUpsample], data: pcm)
    }
}

/// Main speech tokenizer class (SpeechTokenizer.swift:790-852), decode side.
public final class Qwen3TTSSpeechTokenizer {
    public let decoder: Qwen3TTSSpeechTokenizerDecoder
    let decodeUpsampleRate: Int
    let handle: OpaquePointer

    /// Replaces the speech-tokenizer half of `postLoadHook(modelDir:)` (Qwen3.swift:1461-1494):
    /// `speechTokenizerDir` = `<modelDir>/speech_tokenizer` (config.json + *.safetensors).
    public init(speechTokenizerDir: URL, precision: Qwen3TTSPrecision = .fp16,
                attentionMode: Qwen3TTSAttentionMode = .reference, device: Int32 = -1) throws {
        var opts = q3tts_options()
        q3tts_options_default(&opts)
        opts.precision = precision.rawValue
        opts.attn_mode = attentionMode.rawValue
        opts.device = device
        var h: OpaquePointer?
        try check(q3tts_model_load(speechTokenizerDir.path, &opts, &h))
        guard let model = h else { throw Qwen3TTSCUDAError.audioDecodingFailed("q3tts_model_load returned NULL") }
        var cfg = q3tts_config()
        try check(q3tts_model_config(model, &cfg))
        self.handle = model
        self.decodeUpsampleRate = Int(cfg.decode_upsample_rate)
        self.decoder = Qwen3TTSSpeechTokenizerDecoder(handle: model, config: cfg)
    }

    deinit { q3tts_model_free(handle) }

    /// The encoder is a separate handle (`Qwen3TTSSpeechTokenizerEncoder`): a decode-only deployment never loads its weights.
    public var hasEncoder: Bool { false }                      // SpeechTokenizer.swift:816

    /// Decode codes to audio
    /// - Parameter audioCodes: [batch, seq_len, num_quantizers]
    /// - Returns: (audio [batch, samples], audio_lengths [batch])   (SpeechTokenizer.swift:823-836)
    public func decode(_ audioCodes: Int32Tensor) throws -> (audio: FloatTensor, audioLengths: [Int32]) {
        precondition(audioCodes.shape.count == 3 && audioCodes.shape[2] == decoder.numQuantizers)
        let (b, t) = (audioCodes.shape[0], audioCodes.shape[1])
        var pcm = [Float](repeating: 0, count: b * t * decoder.totalUpsample)
        var lengths = [Int32](repeating: 0, count: b)
        try audioCodes.data.withUnsafeBufferPointer { c in
            try pcm.withUnsafeMutableBufferPointer { p in
                try lengths.withUnsafeMutableBufferPointer { l in
                    try check(q3tts_decode(handle, c.baseAddress, Int32(b), Int32(t), Q3TTS_CODES_BTQ,
                                           p.baseAddress, l.baseAddress))
                }
            }
        }
        return (FloatTensor(shape: [b, t * decoder.totalUpsample], data: pcm), lengths)
    }

    /// Decode straight to the 16-bit samples the demo's WAV writer produces (`Int16(clamped * 32767.0)`,
    /// Sources/Qwen3TTSDemo/main.swift:158-160): the conversion runs in the last CUDA kernel.
    public func decodeInt16(_ audioCodes: Int32Tensor) throws -> (audio: [Int16], audioLengths: [Int32]) {
        precondition(audioCodes.shape.count == 3 && audioCodes.shape[2] == decoder.numQuantizers)
        let (b, t) = (audioCodes.shape[0], audioCodes.shape[1])
        var pcm = [Int16](repeating: 0, count: b * t * decoder.totalUpsample)
        var lengths = [Int32](repeating: 0, count: b)
        try audioCodes.data.withUnsafeBufferPointer { c in
            try pcm.withUnsafeMutableBufferPointer { p in
                try lengths.withUnsafeMutableBufferPointer { l in
                    try check(q3tts_decode_int16(handle, c.baseAddress, Int32(b), Int32(t), Q3TTS_CODES_BTQ,
                                                 p.baseAddress, l.baseAddress))
                }
            }
        }
        return (pcm, lengths)
    }

    /// Batch of utterances of different lengths, each [T_i, 16]; every result equals its own B=1 decode.
    public func decodeBatch(_ utterances: [Int32Tensor]) throws -> (audio: [[Float]], audioLengths: [Int32]) {
        var offsets = [Int64](repeating: 0, count: utterances.count + 1)
        var packed = [Int32]()
        for (i, u) in utterances.enumerated() {
            precondition(u.shape.count == 2 && u.shape[1] == decoder.numQuantizers)
            offsets[i + 1] = offsets[i] + Int64(u.shape[0])
            packed.append(contentsOf: u.data)
        }
        let up = decoder.totalUpsample
        var pcm = [Float](repeating: 0, count: Int(offsets.last!) * up)
        var lengths = [Int32](repeating: 0, count: utterances.count)
        try packed.withUnsafeBufferPointer { c in
            try offsets.withUnsafeBufferPointer { o in
                try pcm.withUnsafeMutableBufferPointer { p in
                    try lengths.withUnsafeMutableBufferPointer { l in
                        try check(q3tts_decode_varlen(handle, c.baseAddress, o.baseAddress, Int32(utterances.count),
                                                      p.baseAddress, l.baseAddress))
                    }
                }
            }
        }
        let audio = (0..<utterances.count).map { i in Array(pcm[(Int(offsets[i]) * up)..<(Int(offsets[i + 1]) * up)]) }
        return (audio, lengths)
    }
}

/// Codec-embedding sum for the Talker's next-step input (Qwen3.swift:720-728; 485-491 for the voice-cloning prefix):
/// `talker.getInputEmbeddings()(code0) + codePredictor.codecEmbedding[0](code1) + ...`, left to right in the checkpoint's
/// dtype, for `n` frames in one launch.  `modelDir` holds the main checkpoint's safetensors.
/// Speech-tokenizer encoder: audio -> codes (SpeechTokenizerEncoder.swift:955-1056), for voice cloning (Qwen3.swift:430-440).
/// Same method name and shapes as the reference: `encode([batch, 1, samples]) -> [batch, 16, time]`.
public final class Qwen3TTSSpeechTokenizerEncoder {
    let handle: OpaquePointer
    public let validNumQuantizers: Int                          // SpeechTokenizerEncoder.swift:957

    /// `speechTokenizerDir` must hold an `encoder_config` and the `encoder.*` tensors; a "lite" checkpoint throws
    /// (the reference: `Qwen3TTSSpeechTokenizerError.encoderNotAvailable`, SpeechTokenizer.swift:842-844).
    public init(speechTokenizerDir: URL, device: Int32 = 0) throws {
        var opts = q3tts_options()
        q3tts_options_default(&opts)
        opts.device = device
        var h: OpaquePointer?
        try check(q3tts_encoder_load(speechTokenizerDir.path, &opts, &h))
        guard let enc = h else { throw Qwen3TTSCUDAError.audioDecodingFailed("q3tts_encoder_load returned NULL") }
        var nq: Int32 = 0
        try check(q3tts_encoder_info(enc, &nq, nil, nil, nil, nil))
        self.handle = enc
        self.validNumQuantizers = Int(nq)
    }

    deinit { q3tts_encoder_free(handle) }

    /// - Parameter audio: [batch, 1, samples] (or [batch, samples]) 24 kHz waveform
    /// - Returns: codes [batch, num_quantizers, time]
    public func encode(_ audio: FloatTensor) throws -> Int32Tensor {
        precondition(audio.shape.count == 2 || (audio.shape.count == 3 && audio.shape[1] == 1))
        let (b, samples) = (audio.shape[0], audio.shape[audio.shape.count - 1])
        let frames = Int(q3tts_encode_frames(handle, Int64(samples)))
        var codes = [Int32](repeating: 0, count: b * validNumQuantizers * frames)
        try audio.data.withUnsafeBufferPointer { a in
            try codes.withUnsafeMutableBufferPointer { c in
                try check(q3tts_encode(handle, a.baseAddress, Int32(b), Int64(samples), c.baseAddress))
            }
        }
        return Int32Tensor(shape: [b, validNumQuantizers, frames], data: codes)
    }
}

public final class Qwen3TTSCodecEmbedder {
    private let handle: OpaquePointer
    public let hiddenSize: Int
    public let numCodeGroups: Int

    public init(modelDir: URL, device: Int32 = 0) throws {
        var e: OpaquePointer?
        try check(q3tts_codec_embedder_load(modelDir.path, device, &e))
        guard let opened = e else { throw Qwen3TTSCUDAError.audioDecodingFailed("q3tts_codec_embedder_load returned NULL") }
        handle = opened
        var h: Int32 = 0, g: Int32 = 0, p: Int32 = 0
        try check(q3tts_codec_embedder_info(opened, &h, &g, &p, nil))
        hiddenSize = Int(h)
        numCodeGroups = Int(g)
    }

    deinit { q3tts_codec_embedder_free(handle) }

    /// - Parameter codes: [n, numCodeGroups]
    /// - Returns: n * hiddenSize 16-bit patterns (bf16 / fp16, the checkpoint's dtype), row-major
    public func callAsFunction(_ codes: Int32Tensor) throws -> [UInt16] {
        precondition(codes.shape.count == 2 && codes.shape[1] == numCodeGroups)
        var out = [UInt16](repeating: 0, count: codes.shape[0] * hiddenSize)
        try codes.data.withUnsafeBufferPointer { c in
            try out.withUnsafeMutableBufferPointer { o in
                try check(q3tts_codec_embed_sum(handle, c.baseAddress, Int64(codes.shape[0]), o.baseAddress))
            }
        }
        return out
    }
}

/// Chunked streaming decode: feed codec frames as the Talker emits them (Qwen3.swift:640-729 produces one
/// 16-code vector per step), get the matching 1920·n samples back.  The tokenizer must have been created with
/// `attentionMode: .causalSlidingWindow`; per-stream state is the transformer KV window plus 1-3 frames of input
/// per haloed convolution.  No counterpart in the reference, whose `generateStream` decodes once at the end
/// (Qwen3+Streaming.swift:19-120).
public final class Qwen3TTSDecodeStream {
    private var stream: OpaquePointer
    private let tokenizer: Qwen3TTSSpeechTokenizer      // keeps the model alive for the stream's life
    private let samplesPerFrame: Int                    // what the library writes per frame: total_upsample
    private let numQuantizers: Int

    public init(_ tokenizer: Qwen3TTSSpeechTokenizer) throws {
        var s: OpaquePointer?
        try check(q3tts_stream_open(tokenizer.handle, &s))
        guard let opened = s else { throw Qwen3TTSCUDAError.audioDecodingFailed("q3tts_stream_open returned NULL") }
        stream = opened
        self.tokenizer = tokenizer
        samplesPerFrame = Int(q3tts_output_samples(tokenizer.handle, 1))
        numQuantizers = tokenizer.decoder.numQuantizers
    }

    deinit { q3tts_stream_close(stream) }

    /// Frames decoded so far on this stream.
    public var framesDone: Int { Int(q3tts_stream_frames(stream)) }

    /// - Parameter codes: [n_frames, num_quantizers]
    /// - Returns: n_frames · 1920 samples continuing the stream's waveform
    public func push(_ codes: Int32Tensor) throws -> [Float] {
        precondition(codes.shape.count == 2 && codes.shape[1] == numQuantizers)
        let n = codes.shape[0]
        var pcm = [Float](repeating: 0, count: n * samplesPerFrame)
        try codes.data.withUnsafeBufferPointer { c in
            try pcm.withUnsafeMutableBufferPointer { p in
                try check(q3tts_stream_push(stream, c.baseAddress, Int32(n), p.baseAddress))
            }
        }
        return pcm
    }
}

/// One decoder replica + worker thread per GPU of the box; utterances are sharded by length
/// (longest-processing-time-first) with no collective.  The reference decodes one utterance at a time on one device
/// (Qwen3.swift:744, 951, 1186); this is the batch scheduler of the CUDA path.
public final class Qwen3TTSDecoderPool {
    private let pool: OpaquePointer
    private let totalUpsample: Int
    private let numQuantizers: Int
    public let gpuCount: Int

    public init(speechTokenizerDir: URL, precision: Qwen3TTSPrecision = .fp16, devices: [Int32]? = nil) throws {
        var opts = q3tts_options()
        q3tts_options_default(&opts)
        opts.precision = precision.rawValue
        var cfg = q3tts_config()
        try check(q3tts_checkpoint_inspect(speechTokenizerDir.path, &cfg))
        totalUpsample = Int(cfg.total_upsample)
        numQuantizers = Int(cfg.num_quantizers)
        var p: OpaquePointer?
        if let devs = devices {
            try devs.withUnsafeBufferPointer { d in
                try check(q3tts_pool_open(speechTokenizerDir.path, &opts, d.baseAddress, Int32(devs.count), &p))
            }
        } else {
            try check(q3tts_pool_open(speechTokenizerDir.path, &opts, nil, 0, &p))
        }
        guard let opened = p else { throw Qwen3TTSCUDAError.audioDecodingFailed("q3tts_pool_open returned NULL") }
        pool = opened
        gpuCount = Int(q3tts_pool_size(opened))
    }

    deinit { q3tts_pool_close(pool) }

    /// Utterances of different lengths, each [T_i, 16]; every result equals its own single-utterance decode.
    public func decodeBatch(_ utterances: [Int32Tensor]) throws -> (audio: [[Float]], audioLengths: [Int32]) {
        var offsets = [Int64](repeating: 0, count: utterances.count + 1)
        var packed = [Int32]()
        for (i, u) in utterances.enumerated() {
            precondition(u.shape.count == 2 && u.shape[1] == numQuantizers)
            offsets[i + 1] = offsets[i] + Int64(u.shape[0])
            packed.append(contentsOf: u.data)
        }
        var pcm = [Float](repeating: 0, count: Int(offsets.last!) * totalUpsample)
        var lengths = [Int32](repeating: 0, count: utterances.count)
        try packed.withUnsafeBufferPointer { c in
            try offsets.withUnsafeBufferPointer { o in
                try pcm.withUnsafeMutableBufferPointer { p in
                    try lengths.withUnsafeMutableBufferPointer { l in
                        try check(q3tts_pool_decode_varlen(pool, c.baseAddress, o.baseAddress, Int32(utterances.count),
                                                           p.baseAddress, l.baseAddress))
                    }
                }
            }
        }
        let up = totalUpsample
        let audio = (0..<utterances.count).map { i in Array(pcm[(Int(offsets[i]) * up)..<(Int(offsets[i + 1]) * up)]) }
        return (audio, lengths)
    }
}
